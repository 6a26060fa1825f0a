"""Evaluation without the secret key (ADVICE r1 / VERDICT r1 weak #9).  The reference materialises every evaluation key
up front (infer_seal.cpp:379) and its Evaluator never sees the secret key; this engine prunes keys to the levels they
are used at.  A KeyPlan (host/seal/seal.h) closes the gap: the (Galois element, level) pairs a workload touches are
learnt from a dry run under a THROW-AWAY key, then exactly those keys are generated up front for the real key and the
evaluation keys drop every reference to the secret."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BITS = [50] + [40] * 8 + [50]


def _session(hybrid):
    from b200ckks.app import App

    old = os.environ.get("B200CKKS_HYBRID_KS")
    os.environ["B200CKKS_HYBRID_KS"] = "1" if hybrid else "0"      # read when the engine context is created
    try:
        return App().session(13, BITS, hamming_weight=64, rotation_steps=[1, 3, 0])
    finally:
        if old is None:
            del os.environ["B200CKKS_HYBRID_KS"]
        else:
            os.environ["B200CKKS_HYBRID_KS"] = old


def _workload(s, x, y):
    a = s.encrypt(x, 2.0 ** 40, limbs=7)
    s.rotate(a, 1)                       # rotation at 7 limbs
    b = s.encrypt(y, 2.0 ** 40, limbs=7)
    s.multiply_relin_rescale(a, b)       # relinearization at 7 limbs -> 6 limbs
    s.rotate(a, 3)                       # rotation at 6 limbs
    s.mod_switch_to(a, 3)
    s.rotate(a, 1)                       # the same element again, at 3 limbs
    return a


@pytest.mark.parametrize("hybrid", [False, True])
def test_keys_from_a_plan_need_no_secret_key(hybrid):
    rng = np.random.default_rng(5)
    dry = _session(hybrid)
    x, y = rng.uniform(-1, 1, dry.slots), rng.uniform(-1, 1, dry.slots)
    _workload(dry, x, y)
    plan = dry.key_plan()
    dry.close()
    lines = plan.split("\n")
    assert lines[0] == f"hybrid {int(hybrid)}"
    galois = sorted((int(l.split()[1]), int(l.split()[2])) for l in lines if l.startswith("g "))
    e1, e3 = 5, pow(5, 3, 2 << 13)
    if hybrid:      # one key per (element, level)
        assert galois == sorted([(e1, 7), (e1, 3), (e3, 6)])
        assert [int(l.split()[1]) for l in lines if l.startswith("r ")] == [7]
    else:           # SEAL's layout: one key per element, pruned to the highest level it is used at
        assert galois == sorted([(e1, 7), (e3, 6)])

    s = _session(hybrid)                   # the real key
    s.apply_key_plan(plan, detach_secret=True)
    generated = s.key_residency()[1]
    assert generated >= len(galois)
    out = _workload(s, x, y)
    assert s.key_residency()[1] == generated          # nothing was generated during evaluation
    want = np.roll(np.roll(np.roll(x, -1) * y, -3), -1)
    assert np.abs(s.decrypt(out).real - want).max() < 1e-4
    # outside the plan: like a missing key in the reference (evaluator.cpp:2151-2154), not a silent key generation
    c = s.encrypt(x, 2.0 ** 40, limbs=9)
    with pytest.raises(ValueError, match="not present"):
        s.rotate(c, 3)                     # element 5^3 is only planned for 6 limbs
    if hybrid:
        d = s.encrypt(x, 2.0 ** 40, limbs=5)
        with pytest.raises(ValueError, match="not present"):
            s.multiply_relin_rescale(d, s.encrypt(y, 2.0 ** 40, limbs=5))      # relinearization planned at 7 limbs only
    s.close()
