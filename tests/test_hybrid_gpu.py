"""Level-aware hybrid key switching (tolerance mode, include/b200ckks.h: bk_context_set_hybrid): at level l the idle
primes above the level join the special prime as temporary special moduli and the l limbs are switched in
ceil(l / (alpha - 1)) digits.  Ciphertext limbs differ from SEAL's one-digit-per-prime path by construction, so the parity
statement is on decrypted values: rotation, conjugation, hoisted rotations and relinearization at every level must
decrypt to what the plain operation gives, with the noise of a key switch (the bit-exact path is ~1e-9 here too)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "fhe-gpt-2_b200", "python"))

pytestmark = pytest.mark.gpu

CNN_BITS = [51] + [46] * 16 + [51] * 14 + [51]
GPT2_BITS = [49] + [46] * 21 + [49] * 14 + [60]


def run_levels(bk, log_n, bits, levels):
    primes = bk.coeff_modulus_create(log_n, bits)
    eng = bk.Context(log_n, primes)
    eng.set_hybrid(True)
    top = len(bits) - 1
    sk = eng.generate_secret_key(192 if log_n == 16 else 64, 11)
    pk = eng.create_public_key(sk)
    rk = eng.create_relin_key(sk)
    gk = eng.create_galois_keys(sk, [1, 5, 0])
    rng = np.random.default_rng(log_n)
    x = rng.uniform(-1, 1, eng.slots) + 1j * rng.uniform(-1, 1, eng.slots)
    ct = eng.encrypt(pk, eng.encode(x, top, 2.0 ** 46))
    dec = lambda c: eng.decode(eng.decrypt(sk, c))
    worst = {}
    for l in levels:
        c = ct.copy()
        eng.mod_switch_to_inplace(c, l)
        r = c.copy()
        eng.rotate_vector_inplace(r, 5, gk)
        worst.setdefault("rotate", []).append(np.abs(dec(r) - np.roll(x, -5)).max())
        r = c.copy()
        eng.complex_conjugate_inplace(r, gk)
        worst.setdefault("conjugate", []).append(np.abs(dec(r) - np.conj(x)).max())
        hs = eng.apply_galois_hoisted(c, [bk.galois_elt_from_step(log_n, 1), bk.galois_elt_from_step(log_n, 5)], gk)
        worst.setdefault("hoisted", []).append(max(np.abs(dec(hs[0]) - np.roll(x, -1)).max(), np.abs(dec(hs[1]) - np.roll(x, -5)).max()))
        if l >= 3:
            m = c.copy()
            eng.multiply_inplace(m, c)
            eng.relinearize_inplace(m, rk)
            eng.rescale_to_next_inplace(m)
            worst.setdefault("relinearize", []).append(np.abs(dec(m) - x * x).max())
            # the same pair as one call: ModDown and rescale merged into one division by q_{l-1} P_S where the level
            # has idle primes (elsewhere the two calls) - same level, same scale, same value
            m2 = c.copy()
            eng.multiply_inplace(m2, c)
            eng.relinearize_rescale_inplace(m2, rk)
            assert m2.info()[:2] == m.info()[:2] and abs(m2.info()[2] / m.info()[2] - 1) < 1e-12, (l, m2.info(), m.info())
            worst.setdefault("relinearize+rescale merged", []).append(np.abs(dec(m2) - x * x).max())
            worst.setdefault("merged vs two calls", []).append(np.abs(dec(m2) - dec(m)).max())
            # linear follow-ups move in front of the relinearization: 2 x^2 - x on the size-3 product
            p3 = c.copy()
            eng.multiply_inplace(p3, c)
            lc = eng.scalar_linear_combination([p3, c], [2.0, -1.0], 0.25, p3.info()[2])
            assert lc.info()[0] == 3
            eng.relinearize_rescale_inplace(lc, rk)
            assert lc.info()[:2] == m.info()[:2]
            worst.setdefault("size-3 combination", []).append(np.abs(dec(lc) - (2 * x * x - x + 0.25)).max())
    on, nbytes, keys = eng.hybrid_info()
    assert on and keys >= len(levels) and nbytes > 0
    eng.close()
    return worst


@pytest.mark.parametrize("log_n,bits,levels", [
    (13, CNN_BITS, [31, 30, 29, 28, 25, 20, 17, 12, 9, 5, 3, 2, 1]),
    (16, CNN_BITS, [31, 28, 17, 3]),
    (16, GPT2_BITS, [36, 22, 2]),
])
def test_hybrid_key_switch_decrypts_like_the_reference_path(log_n, bits, levels):
    import b200ckks as bk

    worst = run_levels(bk, log_n, bits, levels)
    # tolerance: the reference's own scheme shows 5e-7 .. 1.3e-6 on a rotation at the top levels (51-bit digits over a
    # 51-bit special prime; tools/hybrid_noise.py prints both paths side by side) - that is also what the top level and
    # l <= 5 use in hybrid mode; in between, digits are one prime smaller than P_S and the error is ~2e-8
    for op, errs in worst.items():
        assert max(errs) < 3e-6, (op, errs)


@pytest.mark.parametrize("log_n,logn", [(12, 9), (16, 14)])
def test_double_hoisted_bootstrap_matches_single_hoisting(log_n, logn):
    """Double hoisting (one division by the special modulus per giant step: bk_bsgs_inner_sums) against the hoisted and
    the rotation-by-rotation bootstraps on the same input: same level, scale and values within the bootstrapping error."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
    import app_cases as cases
    from b200ckks.app import App

    old = os.environ.get("B200CKKS_HYBRID_KS")
    os.environ["B200CKKS_HYBRID_KS"] = "1"
    try:
        s = App().session(log_n, cases.BOOT_BITS, hamming_weight=64 if log_n < 16 else 192)
    finally:
        if old is None:
            del os.environ["B200CKKS_HYBRID_KS"]
        else:
            os.environ["B200CKKS_HYBRID_KS"] = old
    n = 1 << logn
    xs = np.tile(np.random.default_rng(4).uniform(-1, 1, n), s.slots // n)
    boot = s.bootstrapper(logn)
    out = {}
    for name, (on, double) in {"one by one": (False, False), "hoisted": (True, False), "double hoisted": (True, True)}.items():
        boot.set_hoisting(on, double=double)
        before = s.double_hoisted_groups()
        ct = boot.bootstrap(s.encrypt(xs, 2.0 ** 46, limbs=1), real_message=True)
        assert ct.info() == (2, 17, 2.0 ** 46)
        used = s.double_hoisted_groups() - before
        assert (used > 0) == (name == "double hoisted"), (name, used)
        out[name] = s.decrypt(ct)
        assert np.abs(out[name] - xs).max() < 5e-5, name
    assert np.abs(out["double hoisted"] - out["hoisted"]).max() < 5e-5
    s.close()
