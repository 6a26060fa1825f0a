"""Row a13: Evaluator::{add,sub,multiply}_inplace_reduced_error at MISMATCHED levels
(evaluator.cpp:322-344, 380-402, 439-486: the operand with more limbs is multiplied by the constant
s_low * q_last / s_high^2, its scale forced, rescaled, mod-switched down, then the operation).  The engine's facade
composes the same sequence from its own kernels; here its result is compared LIMB BY LIMB with the reference's SEAL
on the same input ciphertexts and the same relinearization key."""
import numpy as np
import pytest

import refseal
from util import SMALL_BITS, rand_slots, ref_fresh_ct

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    from b200ckks.app import App

    ref = refseal.RefSeal(13, SMALL_BITS, hamming_weight=64, seed=29)
    sess = App().session(13, SMALL_BITS, hamming_weight=64)
    assert [int(p) for p in sess.primes()] == [int(p) for p in ref.primes]
    sess.import_relin_key(ref.relin_key())
    ref.make_galois_keys([1])
    yield ref, sess
    sess.close()
    ref.close()


def _check(ref, ct_id, got, what):
    size, limbs, scale, ntt = ref.ct_info(ct_id)
    assert got.info() == (size, limbs, scale), f"{what}: {got.info()} vs {(size, limbs, scale)}"
    want = ref.ct_get(ct_id)
    have = got.download()
    assert np.array_equal(have, want), f"{what}: {(have != want).sum()} of {want.size} words differ"


@pytest.mark.parametrize("op", ["add", "sub", "multiply"])
@pytest.mark.parametrize("la,lb", [(5, 3), (3, 5), (5, 4), (2, 5), (4, 4)])
def test_reduced_error_ops_across_levels(pair, op, la, lb):
    ref, sess = pair
    rng = np.random.default_rng(100 * la + lb)
    # different scales on the two operands, as after different numbers of rescales in the network
    a = ref_fresh_ct(ref, rand_slots(rng, ref.n // 2), la, 2.0 ** 40 * 1.0009)
    b = ref_fresh_ct(ref, rand_slots(rng, ref.n // 2), lb, 2.0 ** 40 * 0.9993)
    ref.op("rotate", b, iarg=1)    # the seeded reference draws the same randomness for every fresh ciphertext: without
    #                                this a - b at equal levels has c1 = 0 and SEAL refuses the transparent result
    ea = sess.upload(ref.ct_get(a), ref.ct_info(a)[2])
    eb = sess.upload(ref.ct_get(b), ref.ct_info(b)[2])
    ref.op(op + "_reduced_error", a, b)
    sess.reduced_error_op(op, ea, eb)
    _check(ref, a, ea, f"{op}_inplace_reduced_error at {la} and {lb} limbs")
    # the second operand is left as it was (evaluator.cpp:325: works on a copy)
    _check(ref, b, eb, "second operand")
