"""GPU parity: every hot-path entry point of the C ABI against the reference's own SEAL
(oracle/_ref/libseal_ref.so) on identical raw limbs.  Integer work is compared BIT-EXACTLY."""
import numpy as np
import pytest

import refseal
from util import (CNN_BITS, GPT2_BITS, SMALL_BITS, assert_ct_equal, rand_slots, ref_fresh_ct, to_engine)

pytestmark = pytest.mark.gpu

if not refseal.available():  # pragma: no cover
    pytest.skip("oracle/_ref/libseal_ref.so not built", allow_module_level=True)


@pytest.fixture(scope="module")
def small():
    import b200ckks as bk

    ref = refseal.RefSeal(13, SMALL_BITS, hamming_weight=64, seed=11)
    eng = bk.Context(13, ref.primes)
    rk = eng.upload_kskey(ref.relin_key())
    steps = [1, -3, 0, 64]
    ref.make_galois_keys(steps)
    gk = eng.galois_keys()
    for st in steps:
        elt = ref.galois_elt(st)
        gk.set(elt, eng.upload_kskey(ref.galois_key(elt)))
    yield ref, eng, rk, gk
    eng.close()
    ref.close()


@pytest.mark.parametrize("log_n,bits", [(12, [40, 40, 41]), (13, SMALL_BITS), (14, [50, 45, 45, 50]),
                                        (15, [55, 50, 50, 55])])
def test_ntt_matches_reference_small(log_n, bits):
    import b200ckks as bk

    ref = refseal.RefSeal(log_n, bits, hamming_weight=0, seed=3)
    eng = bk.Context(log_n, ref.primes)
    rng = np.random.default_rng(log_n)
    n = 1 << log_n
    idx = list(range(len(bits)))
    data = np.stack([rng.integers(0, int(ref.primes[i]), n, dtype=np.uint64) for i in idx])
    want = np.stack([ref.ntt(i, data[i]) for i in idx])
    got = eng.ntt_limbs_host(data, idx)
    assert np.array_equal(got, want)
    back = eng.ntt_limbs_host(got, idx, inverse=True)
    assert np.array_equal(back, data)
    want_inv = np.stack([ref.ntt(i, data[i], inverse=True) for i in idx])
    assert np.array_equal(eng.ntt_limbs_host(data, idx, inverse=True), want_inv)
    # edge values: 0, q-1
    edge = np.zeros_like(data)
    edge[:, ::2] = (ref.primes[idx] - np.uint64(1))[:, None]
    assert np.array_equal(eng.ntt_limbs_host(edge, idx), np.stack([ref.ntt(i, edge[i]) for i in idx]))
    eng.close()
    ref.close()


def test_ntt_matches_reference_n65536_all_cnn_primes():
    import b200ckks as bk

    ref = refseal.RefSeal(16, CNN_BITS, hamming_weight=192, seed=5)
    eng = bk.Context(16, ref.primes)
    rng = np.random.default_rng(0x5EA1)
    idx = list(range(32))
    data = np.stack([rng.integers(0, int(ref.primes[i]), 65536, dtype=np.uint64) for i in idx])
    got = eng.ntt_limbs_host(data, idx)
    for i in (0, 1, 16, 17, 30, 31):
        assert np.array_equal(got[i], ref.ntt(i, data[i])), f"prime {i}"
    assert np.array_equal(eng.ntt_limbs_host(got, idx, inverse=True), data)
    for i in (0, 5, 31):
        assert np.array_equal(eng.ntt_limbs_host(data[i:i + 1], [i], inverse=True)[0], ref.ntt(i, data[i], inverse=True))
    eng.close()
    ref.close()


def _pair(small, limbs, seed, scale=2.0 ** 40):
    ref, eng, rk, gk = small
    rng = np.random.default_rng(seed)
    a = ref_fresh_ct(ref, rand_slots(rng, ref.n // 2), limbs, scale)
    b = ref_fresh_ct(ref, rand_slots(rng, ref.n // 2, complex_=True), limbs, scale)
    return a, b


@pytest.mark.parametrize("limbs", [5, 3, 1])
def test_add_sub_negate(small, limbs):
    ref, eng, rk, gk = small
    a, b = _pair(small, limbs, 1)
    ea, eb = to_engine(eng, ref, a), to_engine(eng, ref, b)
    ref.op("add", a, b)
    eng.add_inplace(ea, eb)
    assert_ct_equal(ea, ref, a, "add")
    # (the seeded reference PRNG hands every encryption the same randomness, so a - b would be a
    # transparent ciphertext, which the reference refuses: go through 2a + b first)
    ref.op("add", a, a)
    eng.add_inplace(ea, ea)
    ref.op("sub", a, b)
    eng.sub_inplace(ea, eb)
    assert_ct_equal(ea, ref, a, "sub")
    ref.op("negate", a)
    eng.negate_inplace(ea)
    assert_ct_equal(ea, ref, a, "negate")


@pytest.mark.parametrize("limbs", [5, 2])
def test_multiply_relinearize_rescale(small, limbs):
    ref, eng, rk, gk = small
    a, b = _pair(small, limbs, 2)
    ea, eb = to_engine(eng, ref, a), to_engine(eng, ref, b)
    ref.op("multiply", a, b)
    eng.multiply_inplace(ea, eb)
    assert_ct_equal(ea, ref, a, "multiply")
    # size-3 add/sub against size-2
    c = ref.ct_new()
    ref.ct_copy(c, b)
    ec = eb.copy()
    ref.op("mod_switch_to", c, iarg=limbs)
    ref.ct_set_scale(c, ref.ct_info(a)[2])
    ec.scale = ea.scale
    ref.op("sub", c, a)
    eng.sub_inplace(ec, ea)
    assert_ct_equal(ec, ref, c, "sub size2 - size3")
    ref.op("relinearize", a)
    eng.relinearize_inplace(ea, rk)
    assert_ct_equal(ea, ref, a, "relinearize")
    ref.op("rescale", a)
    eng.rescale_to_next_inplace(ea)
    assert_ct_equal(ea, ref, a, "rescale")


@pytest.mark.parametrize("limbs", [5, 2])
def test_relinearize_rescale_as_one_call_is_the_two_calls_on_the_exact_path(small, limbs):
    """bk_relinearize_rescale_inplace outside hybrid mode: relinearize_inplace then rescale_to_next_inplace, limb for limb
    what the reference's two calls give (evaluator.cpp:1061-1116, 1378-1414)."""
    ref, eng, rk, gk = small
    a, b = _pair(small, limbs, 7)
    ea, eb = to_engine(eng, ref, a), to_engine(eng, ref, b)
    ref.op("multiply", a, b)
    eng.multiply_inplace(ea, eb)
    ref.op("relinearize", a)
    ref.op("rescale", a)
    eng.relinearize_rescale_inplace(ea, rk)
    assert_ct_equal(ea, ref, a, "relinearize + rescale in one call")


def test_square_and_rescale_size3(small):
    ref, eng, rk, gk = small
    a, _ = _pair(small, 4, 3)
    ea = to_engine(eng, ref, a)
    ref.op("square", a)
    eng.square_inplace(ea)
    assert_ct_equal(ea, ref, a, "square")
    ref.op("rescale", a)
    eng.rescale_to_next_inplace(ea)
    assert_ct_equal(ea, ref, a, "rescale size 3")
    ref.op("relinearize", a)
    eng.relinearize_inplace(ea, rk)
    assert_ct_equal(ea, ref, a, "relinearize after rescale")


@pytest.mark.parametrize("limbs", [5, 3, 1])
@pytest.mark.parametrize("step", [1, -3, 64])
def test_rotate(small, limbs, step):
    ref, eng, rk, gk = small
    a, _ = _pair(small, limbs, 4 + limbs)
    ea = to_engine(eng, ref, a)
    ref.op("rotate", a, iarg=step)
    eng.rotate_vector_inplace(ea, step, gk)
    assert_ct_equal(ea, ref, a, f"rotate {step}")


def test_rotate_naf_fallback_and_conjugate(small):
    ref, eng, rk, gk = small
    _, b = _pair(small, 3, 9)
    eb = to_engine(eng, ref, b)
    ref.op("rotate", b, iarg=65)  # 65 = 64 + 1 -> NAF path (evaluator.cpp:2256-2278)
    eng.rotate_vector_inplace(eb, 65, gk)
    assert_ct_equal(eb, ref, b, "rotate 65 via NAF")
    ref.op("conjugate", b)
    eng.complex_conjugate_inplace(eb, gk)
    assert_ct_equal(eb, ref, b, "conjugate")
    import b200ckks as bk
    with pytest.raises(bk.InvalidArgument, match="Galois key not present"):
        eng.rotate_vector_inplace(eb, 2, gk)


def test_mod_switch(small):
    ref, eng, rk, gk = small
    a, _ = _pair(small, 5, 12)
    ea = to_engine(eng, ref, a)
    ref.op("mod_switch_next", a)
    eng.mod_switch_to_next_inplace(ea)
    assert_ct_equal(ea, ref, a, "mod_switch_to_next")
    ref.op("mod_switch_to", a, iarg=2)
    eng.mod_switch_to_inplace(ea, 2)
    assert_ct_equal(ea, ref, a, "mod_switch_to")


def test_ntt_transforms_of_ciphertext(small):
    ref, eng, rk, gk = small
    a, _ = _pair(small, 4, 13)
    ea = to_engine(eng, ref, a)
    ref.op("ntt_inv", a)
    eng.transform_from_ntt_inplace(ea)
    assert_ct_equal(ea, ref, a, "transform_from_ntt")
    ref.op("ntt_fwd", a)
    eng.transform_to_ntt_inplace(ea)
    assert_ct_equal(ea, ref, a, "transform_to_ntt")


@pytest.mark.parametrize("limbs", [5, 2])
def test_plain_ops_and_encode(small, limbs):
    ref, eng, rk, gk = small
    rng = np.random.default_rng(21)
    a, _ = _pair(small, limbs, 14)
    ea = to_engine(eng, ref, a)
    vals = rand_slots(rng, ref.n // 2, complex_=True)
    pt = ref.pt_new()
    scale = ref.ct_info(a)[2]
    ref.encode(pt, vals, limbs, scale)
    ept = eng.encode(vals, limbs, scale)
    assert np.array_equal(ept.download(), ref.pt_get(pt)), "encode complex"
    ref.op("add_plain", a, pt)
    eng.add_plain_inplace(ea, ept)
    assert_ct_equal(ea, ref, a, "add_plain")
    ref.op("multiply_plain", a, pt)
    eng.multiply_plain_inplace(ea, ept)
    assert_ct_equal(ea, ref, a, "multiply_plain")
    # real input, shorter than slot count, encoded at the new scale
    vals2 = rand_slots(rng, 100)
    scale2 = ref.ct_info(a)[2]
    ref.encode(pt, vals2, limbs, scale2)
    ept2 = eng.encode(vals2, limbs, scale2)
    assert np.array_equal(ept2.download(), ref.pt_get(pt)), "encode real short"
    ref.op("sub_plain", a, pt)
    eng.sub_plain_inplace(ea, ept2)
    assert_ct_equal(ea, ref, a, "sub_plain")


def test_const_ops(small):
    ref, eng, rk, gk = small
    a, _ = _pair(small, 4, 15)
    ea = to_engine(eng, ref, a)
    for v in (0.5, -1.25e-3, 3.0):
        ref.op("add_const", a, darg=v)
        eng.add_const_inplace(ea, v)
        assert_ct_equal(ea, ref, a, f"add_const {v}")
    ref.op("multiply_const", a, darg=-0.37)
    eng.multiply_const_inplace(ea, -0.37)
    assert_ct_equal(ea, ref, a, "multiply_const")
    ref.op("rescale", a)
    eng.rescale_to_next_inplace(ea)
    assert_ct_equal(ea, ref, a, "rescale after multiply_const")


def test_scalar_linear_combination(small):
    """bk_scalar_linear_combination (the one-pass leaves of the polynomial evaluation trees).  With all terms at one
    level and scale it must equal, limb by limb, the reference's sequence multiply_const per term + add + add_const
    (same integer scalars: round(v * scale) mod q); with terms at different levels and scales every limb must equal
    sum_j round(v_j * target / scale_j) * ct_j + round(c * target) computed in exact integer arithmetic."""
    ref, eng, rk, gk = small
    vals = (0.37, -1.5e-3, 2.25)
    cts = [_pair(small, 4, 50 + j)[0] for j in range(3)]
    ecs = [to_engine(eng, ref, c) for c in cts]
    scale = ecs[0].scale
    got = eng.scalar_linear_combination(ecs, vals, -0.625, scale * scale)
    for c, v in zip(cts, vals):
        ref.op("multiply_const", c, darg=v)
    ref.op("add", cts[0], cts[1])
    ref.op("add", cts[0], cts[2])
    ref.op("add_const", cts[0], darg=-0.625)
    assert_ct_equal(got, ref, cts[0], "linear combination vs multiply_const / add / add_const")

    # mismatched levels and scales: exact integers
    primes = [int(q) for q in eng.primes]
    a = to_engine(eng, ref, _pair(small, 5, 60)[0])
    b = to_engine(eng, ref, _pair(small, 3, 61, scale=2.0 ** 38)[0])
    target = 2.0 ** 79
    out = eng.scalar_linear_combination([a, b], (0.75, -0.2), 0.125, target)
    assert out.info()[:3] == (2, 3, target)
    da, db, do = a.download(), b.download(), out.download()
    for i in range(3):
        q = primes[i]
        ra, rb = round(0.75 * target / a.scale) % q, round(-0.2 * target / b.scale) % q
        rc = round(0.125 * target) % q
        for p in range(2):
            want = (da[p, i].astype(object) * ra + db[p, i].astype(object) * rb + (rc if p == 0 else 0)) % q
            assert np.array_equal(do[p, i].astype(object), want), (p, i)


def test_multiply_vector_reduced_error_path(small):
    """evaluator.h:1270-1278: encode at the top level, drop to the ciphertext level, multiply."""
    ref, eng, rk, gk = small
    rng = np.random.default_rng(33)
    a, _ = _pair(small, 3, 16)
    ea = to_engine(eng, ref, a)
    vals = rand_slots(rng, ref.n // 2)
    ref.multiply_vector_reduced_error(a, vals)
    pt = eng.encode(vals, ea.limbs, ea.scale, top_dropped=True)
    eng.multiply_plain_inplace(ea, pt)
    assert_ct_equal(ea, ref, a, "multiply_vector_reduced_error")


def test_decode_and_decrypt_match_reference(small):
    ref, eng, rk, gk = small
    rng = np.random.default_rng(44)
    vals = rand_slots(rng, ref.n // 2, complex_=True)
    a = ref_fresh_ct(ref, vals, 3, 2.0 ** 40)
    ea = to_engine(eng, ref, a)
    sk = eng.upload_secret_key(ref.secret_key())
    pt = ref.pt_new()
    ref.decrypt(a, pt)
    ept = eng.decrypt(sk, ea)
    assert np.array_equal(ept.download(), ref.pt_get(pt)), "decrypt"
    want = ref.decode(pt)
    got = eng.decode(ept)
    assert np.max(np.abs(got - want)) < 1e-12
    assert np.max(np.abs(got - vals)) < 1e-6


def test_native_keygen_encrypt_roundtrip(small):
    """Keys/ciphertexts generated on the device use a different PRNG than the reference, so they
    are checked functionally: decrypt(op(encrypt(x))) ~= op(x)."""
    ref, eng, _, _ = small
    rng = np.random.default_rng(55)
    sk = eng.generate_secret_key(hamming_weight=64, seed=77)
    s = sk.download()
    pk = eng.create_public_key(sk)
    rk = eng.create_relin_key(sk)
    gk = eng.create_galois_keys(sk, [1, 0])
    x = rand_slots(rng, eng.slots, complex_=True)
    y = rand_slots(rng, eng.slots)
    scale = 2.0 ** 40
    cx = eng.encrypt(pk, eng.encode(x, 5, scale))
    cy = eng.encrypt_symmetric(sk, eng.encode(y, 5, scale))
    assert cx.limbs == 5 and cy.limbs == 5
    assert np.max(np.abs(eng.decode(eng.decrypt(sk, cx)) - x)) < 1e-6
    eng.multiply_inplace(cx, cy)
    eng.relinearize_inplace(cx, rk)
    eng.rescale_to_next_inplace(cx)
    assert np.max(np.abs(eng.decode(eng.decrypt(sk, cx)) - x * y)) < 1e-5
    eng.rotate_vector_inplace(cx, 1, gk)
    assert np.max(np.abs(eng.decode(eng.decrypt(sk, cx)) - np.roll(x * y, -1))) < 1e-5
    eng.complex_conjugate_inplace(cx, gk)
    assert np.max(np.abs(eng.decode(eng.decrypt(sk, cx)) - np.conj(np.roll(x * y, -1)))) < 1e-5
    # the secret really has Hamming weight 64 (coefficient form = INTT of limb 0)
    coeffs = eng.ntt_limbs_host(s[:1], [0], inverse=True)[0]
    q0 = int(eng.primes[0])
    nz = [int(v) for v in coeffs if v]
    assert len(nz) == 64 and all(v in (1, q0 - 1) for v in nz)


def test_modraise_matches_bootstrapper_loop(small):
    """Bootstrapper::modraise_inplace (Bootstrapper.cpp:2894-2948) restated in numpy on the
    reference's own INTT/NTT."""
    ref, eng, rk, gk = small
    a, _ = _pair(small, 1, 66)
    ea = to_engine(eng, ref, a)
    raw = ref.ct_get(a)  # [2][1][N] NTT form
    q = [int(v) for v in ref.primes]
    L = ref.n_primes - 1
    want = np.zeros((2, L, ref.n), dtype=np.uint64)
    for p in range(2):
        c = ref.ntt(0, raw[p, 0], inverse=True)
        big = c > np.uint64(q[0] >> 1)
        for j in range(L):
            r = c % np.uint64(q[j])
            minus = np.uint64(0 if j == 0 else q[j] - q[0] % q[j])
            r2 = np.where(big, r + minus, r)
            r2 = np.where(r2 >= np.uint64(q[j]), r2 - np.uint64(q[j]), r2)
            want[p, j] = ref.ntt(j, r2)
    eng.modraise_inplace(ea)
    assert ea.limbs == L and ea.is_ntt_form
    assert np.array_equal(ea.download(), want)


def test_error_behaviour(small):
    import b200ckks as bk

    ref, eng, rk, gk = small
    a, b = _pair(small, 3, 70)
    ea, eb = to_engine(eng, ref, a), to_engine(eng, ref, b)
    eb.scale = eb.scale * 2
    with pytest.raises(bk.InvalidArgument, match="scale mismatch"):
        eng.add_inplace(ea, eb)
    eng.mod_switch_to_next_inplace(eb)
    with pytest.raises(bk.InvalidArgument, match="parameter mismatch"):
        eng.multiply_inplace(ea, eb)
    ec = ea.copy()
    eng.mod_switch_to_inplace(ec, 1)
    with pytest.raises(bk.InvalidArgument, match="end of modulus switching chain"):
        eng.rescale_to_next_inplace(ec)
    ea.scale = 2.0 ** 100
    with pytest.raises(bk.InvalidArgument, match="scale out of bounds"):
        eng.square_inplace(ea)


# ---- N = 2^16, the CNN's parameter set -------------------------------------------------------------
@pytest.fixture(scope="module")
def cnn():
    import b200ckks as bk

    ref = refseal.RefSeal(16, CNN_BITS, hamming_weight=192, seed=7)
    eng = bk.Context(16, ref.primes)
    rk = eng.upload_kskey(ref.relin_key())
    ref.make_galois_keys([1])
    gk = eng.galois_keys()
    elt = ref.galois_elt(1)
    gk.set(elt, eng.upload_kskey(ref.galois_key(elt)))
    yield ref, eng, rk, gk
    eng.close()
    ref.close()


@pytest.mark.parametrize("limbs", [31, 17, 3])
def test_cnn_parameters_keyswitch_rescale(cnn, limbs):
    ref, eng, rk, gk = cnn
    rng = np.random.default_rng(limbs)
    scale = 2.0 ** 46
    a = ref_fresh_ct(ref, rand_slots(rng, 32768), limbs, scale)
    b = ref_fresh_ct(ref, rand_slots(rng, 32768), limbs, scale)
    ea, eb = to_engine(eng, ref, a), to_engine(eng, ref, b)
    ref.op("rotate", a, iarg=1)
    eng.rotate_vector_inplace(ea, 1, gk)
    assert_ct_equal(ea, ref, a, "rotate")
    ref.op("multiply", a, b)
    eng.multiply_inplace(ea, eb)
    assert_ct_equal(ea, ref, a, "multiply")
    ref.op("relinearize", a)
    eng.relinearize_inplace(ea, rk)
    assert_ct_equal(ea, ref, a, "relinearize")
    ref.op("rescale", a)
    eng.rescale_to_next_inplace(ea)
    assert_ct_equal(ea, ref, a, "rescale")
    vals = rand_slots(rng, 32768)
    ref.multiply_vector_reduced_error(a, vals)
    pt = eng.encode(vals, ea.limbs, ea.scale, top_dropped=True)
    eng.multiply_plain_inplace(ea, pt)
    assert_ct_equal(ea, ref, a, "multiply_vector_reduced_error")


def test_cnn_parameters_ks_chunk_invariance(cnn):
    ref, eng, rk, gk = cnn
    rng = np.random.default_rng(99)
    a = ref_fresh_ct(ref, rand_slots(rng, 32768), 20, 2.0 ** 46)
    ea = to_engine(eng, ref, a)
    ref.op("rotate", a, iarg=1)
    for chunk in (1, 3, 8, 32):
        eng.set_ks_chunk(chunk)
        e2 = ea.copy()
        eng.rotate_vector_inplace(e2, 1, gk)
        assert_ct_equal(e2, ref, a, f"rotate chunk={chunk}")
    eng.set_ks_chunk(4)


# ---- N = 2^16, the GPT-2 parameter set (gpt2 util.h:22-27: 49 | 46x21 | 49x14 | 60, 37 primes) ----------------
# the 60-bit special prime is above 2^57, so this chain runs the classic (reduced) forward butterflies
@pytest.mark.parametrize("limbs", [36, 22, 2])
def test_gpt2_parameters_keyswitch_rescale(limbs):
    import b200ckks as bk

    ref = refseal.RefSeal(16, GPT2_BITS, hamming_weight=192, seed=9)
    eng = bk.Context(16, ref.primes)
    rk = eng.upload_kskey(ref.relin_key(), max_limbs=limbs)
    ref.make_galois_keys([3])
    gk = eng.galois_keys()
    elt = ref.galois_elt(3)
    gk.set(elt, eng.upload_kskey(ref.galois_key(elt), max_limbs=limbs))
    rng = np.random.default_rng(limbs)
    scale = 2.0 ** 46
    a = ref_fresh_ct(ref, rand_slots(rng, 32768, complex_=True), limbs, scale)
    b = ref_fresh_ct(ref, rand_slots(rng, 32768), limbs, scale)
    ea, eb = to_engine(eng, ref, a), to_engine(eng, ref, b)
    ref.op("rotate", a, iarg=3)
    eng.rotate_vector_inplace(ea, 3, gk)
    assert_ct_equal(ea, ref, a, "rotate")
    ref.op("multiply", a, b)
    eng.multiply_inplace(ea, eb)
    ref.op("relinearize", a)
    eng.relinearize_inplace(ea, rk)
    assert_ct_equal(ea, ref, a, "multiply + relinearize")
    ref.op("rescale", a)
    eng.rescale_to_next_inplace(ea)
    assert_ct_equal(ea, ref, a, "rescale")
    eng.close()
    ref.close()


def test_hoisted_rotations_decrypt_like_plain_rotations(small):
    """bk_apply_galois_hoisted (engine extension): one decomposition for a set of automorphisms.  Same slots as
    rotate_vector up to key-switching noise; NOT the reference's limbs (the decomposition precedes the automorphism),
    which is why only the opt-in BSGS fast path uses it."""
    import b200ckks as bk

    ref, eng, rk, gk = small
    rng = np.random.default_rng(77)
    x = rand_slots(rng, 4096, complex_=True)
    sk = eng.upload_secret_key(ref.secret_key())
    for limbs in (5, 2):
        a = ref_fresh_ct(ref, x, limbs, 2.0 ** 40)
        ea = to_engine(eng, ref, a)
        steps = [1, -3, 64]
        elts = [ref.galois_elt(s) for s in steps] + [ref.galois_elt(0)]
        outs = eng.apply_galois_hoisted(ea, elts, gk)
        for st, o in zip(steps, outs):
            got = eng.decode(eng.decrypt(sk, o))
            assert np.abs(got - np.roll(x, -st)).max() < 1e-6
            plain = ea.copy()
            eng.rotate_vector_inplace(plain, st, gk)
            assert o.info() == plain.info()
            # same message, different (equally valid) key-switching noise: the limbs are not the reference's
            assert not np.array_equal(o.download(), plain.download())
        assert np.abs(eng.decode(eng.decrypt(sk, outs[-1])) - np.conj(x)).max() < 1e-6
    with pytest.raises(bk.InvalidArgument):
        eng.apply_galois_hoisted(ea, [ref.galois_elt(7)], gk)        # no such key
