"""CPU tests (no GPU): pin the plain-C oracle (oracle/ckks_port.c) against
 (a) the known-answer vectors in the reference's own unit tests
     (seal-modified-3.6.6/native/tests/seal/util/*.cpp, cited per test), and
 (b) the reference itself (oracle/_ref/libseal_ref.so, compiled in place from /root/reference)
     on seeded inputs - incl. a whole key switch, rescale and rotation."""
import numpy as np
import pytest

import ckks_port as port
import refseal

if not port.available():  # pragma: no cover
    pytest.skip("oracle/_ref/libckks_port.so not built (python -c 'import __graft_entry__ as g; g.build()')",
                allow_module_level=True)

Q60 = 0xFFFFFFFFFFC0001


def test_kat_primitive_roots():
    """tests/seal/util/ntt.cpp:55-75 NTTPrimitiveRootsTest"""
    t = port.Tables(1, [Q60])
    rp = t.root_powers(0)
    assert rp[0] == 1 and rp[1] == 288794978602139552
    inv = pow(288794978602139552, Q60 - 2, Q60)
    assert t.root_powers(0, inverse=True)[1] == inv
    t = port.Tables(2, [Q60])
    assert list(t.root_powers(0)) == [1, 288794978602139552, 178930308976060547, 748001537669050592]


def test_kat_negacyclic_ntt():
    """tests/seal/util/ntt.cpp:77-103 NegacyclicNTTTest"""
    t = port.Tables(1, [Q60])
    assert list(t.ntt(0, [0, 0])) == [0, 0]
    assert list(t.ntt(0, [1, 0])) == [1, 1]
    assert list(t.ntt(0, [1, 1])) == [288794978602139553, 864126526004445282]


def test_kat_inverse_ntt_roundtrip():
    """tests/seal/util/ntt.cpp:105-133 InverseNegacyclicNTTTest (n = 8)"""
    t = port.Tables(3, [Q60])
    assert not t.ntt(0, np.zeros(8, np.uint64), inverse=True).any()
    rng = np.random.default_rng(0)
    x = rng.integers(0, Q60, 8, dtype=np.uint64)
    assert np.array_equal(t.ntt(0, t.ntt(0, x), inverse=True), x)


def test_kat_barrett_and_operand():
    """tests/seal/util/uintarithsmallmod.cpp:142-212 BarrettReduce128, :375-404 MultiplyUIntModOperand"""
    L = port.lib()
    M = 0xFFFFFFFFFFFFFFFF
    assert L.port_barrett_reduce_128(0, 0, 2) == 0 and L.port_barrett_reduce_128(1, 0, 2) == 1
    assert L.port_barrett_reduce_128(M, M, 2) == 1
    assert L.port_barrett_reduce_128(123, 456, 3) == 0 and L.port_barrett_reduce_128(M, M, 3) == 0
    q = 13131313131313
    for lo, hi in ((0, 0), (1, 0), (123, 456), (M, M), (24242424242424, 79797979797979)):
        assert L.port_barrett_reduce_128(lo, hi, q) == ((hi << 64) | lo) % q
    assert L.port_shoup_quotient(1, 3) == 6148914691236517205
    assert L.port_shoup_quotient(2, 3) == 12297829382473034410
    assert L.port_shoup_quotient(1, 2147483647) == 8589934596
    assert L.port_shoup_quotient(2147483646, 2147483647) == 18446744065119617019
    assert L.port_shoup_quotient(1, 2305843009211596801) == 8
    assert L.port_shoup_quotient(2305843009211596800, 2305843009211596801) == 18446744073709551607


def test_kat_dyadic_product():
    """tests/seal/util/polyarithsmallmod.cpp:545-590 DyadicProductCoeffMod"""
    assert list(port.dyadic_product([1, 1, 1], [2, 3, 4], 13)) == [2, 3, 4]
    assert list(port.dyadic_product([1, 2, 1], [2, 3, 4], 13)) == [2, 6, 4]
    assert list(port.dyadic_product([2, 1, 2], [2, 3, 4], 7)) == [4, 3, 1]


def test_kat_apply_galois():
    """tests/seal/util/galois.cpp:81-115 ApplyGalois / ApplyGaloisNTT (n = 8, elt = 3, q = 17)"""
    x = np.arange(8, dtype=np.uint64)
    assert list(port.apply_galois(x, 3, 3, 17)) == [0, 14, 6, 1, 13, 7, 2, 12]
    assert list(port.apply_galois_ntt(x, 3, 3)) == [4, 5, 7, 6, 1, 0, 2, 3]


def test_kat_divide_and_round_q_last_ntt():
    """tests/seal/util/rns.cpp:1010-1072 DivideAndRoundQLastNTTInplace (n = 2, base {53, 13})"""
    t = port.Tables(1, [53, 13])

    def run(c53, c13):
        poly = np.stack([t.ntt(0, c53), t.ntt(1, c13)])
        out = t.divide_and_round_q_last_ntt(poly)
        return [int(v) for v in t.ntt(0, out[0], inverse=True)]

    assert run([0, 0], [0, 0]) == [0, 0]
    assert run([1, 2], [1, 2]) == [0, 0]
    a = run([4, 12], [4, 12])
    assert (53 + 1 - a[0]) % 53 <= 1 and (53 + 2 - a[1]) % 53 <= 1
    a = run([25, 35], [12, 9])
    assert (53 + 2 - a[0]) % 53 <= 1 and (53 + 3 - a[1]) % 53 <= 1


def test_galois_elt_from_step_uses_generator_5():
    """galois.h:169 (this fork: generator 5) + galois.cpp:53-95; stock SEAL's KAT (generator 3,
    tests/seal/util/galois.cpp:28-41) therefore does NOT apply - check against the built reference."""
    assert port.galois_elt_from_step(3, 0) == 15
    assert port.galois_elt_from_step(3, 1) == 5
    assert port.galois_elt_from_step(16, 1) == 5 and port.galois_elt_from_step(16, 2) == 25


needs_ref = pytest.mark.skipif(not refseal.available(), reason="oracle/_ref/libseal_ref.so not built")


@pytest.fixture(scope="module")
def ref12():
    r = refseal.RefSeal(12, [40, 36, 36, 40], hamming_weight=32, seed=21)
    yield r
    r.close()


@needs_ref
def test_port_tables_and_ntt_match_reference(ref12):
    t = port.Tables(12, ref12.primes)
    rng = np.random.default_rng(1)
    for i in range(4):
        assert np.array_equal(t.root_powers(i), ref12.root_powers(i))
        assert np.array_equal(t.root_powers(i, True), ref12.root_powers(i, True))
        x = rng.integers(0, int(ref12.primes[i]), 4096, dtype=np.uint64)
        assert np.array_equal(t.ntt(i, x), ref12.ntt(i, x))
        assert np.array_equal(t.ntt(i, x, inverse=True), ref12.ntt(i, x, inverse=True))
    for step in (0, 1, -1, 7, -100, 2047):
        assert port.galois_elt_from_step(12, step) == ref12.galois_elt(step)
    x = rng.integers(0, int(ref12.primes[0]), 4096, dtype=np.uint64)
    for elt in (5, 25, 8191, 3):
        assert np.array_equal(port.apply_galois_ntt(x, 12, elt), ref12.apply_galois_ntt(elt, x))


def _fresh(ref, rng, limbs, scale=2.0 ** 15):
    pt, ct = ref.pt_new(), ref.ct_new()
    ref.encode(pt, rng.uniform(-1, 1, ref.n // 2), ref.n_primes - 1, scale)
    ref.encrypt(pt, ct)
    if limbs < ref.n_primes - 1:
        ref.op("mod_switch_to", ct, iarg=limbs)
    return ct


@needs_ref
@pytest.mark.parametrize("limbs", [3, 2, 1])
def test_port_multiply_relinearize_rescale_rotate_match_reference(ref12, limbs):
    ref = ref12
    t = port.Tables(12, ref.primes)
    rng = np.random.default_rng(limbs)
    a, b = _fresh(ref, rng, limbs), _fresh(ref, rng, limbs)
    ra, rb = ref.ct_get(a), ref.ct_get(b)
    rk = ref.relin_key()
    ref.op("multiply", a, b)
    prod = t.multiply(ra, rb)
    assert np.array_equal(prod, ref.ct_get(a))
    ref.op("relinearize", a)
    relin = t.relinearize(prod, rk)
    assert np.array_equal(relin, ref.ct_get(a))
    if limbs > 1:
        ref.op("rescale", a)
        # only the first `limbs` tables take part at this level
        t_level = port.Tables(12, ref.primes[:limbs])
        assert np.array_equal(t_level.rescale(relin), ref.ct_get(a))
    ref.make_galois_keys([3])
    elt = ref.galois_elt(3)
    gk = ref.galois_key(elt)
    ref.op("rotate", b, iarg=3)
    assert np.array_equal(t.apply_galois_ct(rb, elt, gk), ref.ct_get(b))


@needs_ref
def test_port_modraise_matches_reference_arithmetic(ref12):
    """Bootstrapper.cpp:2928-2944 uses plain % on the reference's primes; compare with a
    big-integer centred lift."""
    rng = np.random.default_rng(5)
    q = [int(v) for v in ref12.primes[:3]]
    src = rng.integers(0, q[0], 4096, dtype=np.uint64)
    out = port.modraise_coeffs(src, q)
    for j in range(3):
        want = [(int(v) - q[0] if int(v) > q[0] >> 1 else int(v)) % q[j] for v in src]
        assert [int(v) for v in out[j]] == want
