"""Application layers on the B200 engine, through include/b200ckks_app.h: the same cases as tests/test_app_cpu.py
(small parameters), the bootstrapping entry points at the reference's N = 2^16 parameter set, and the op counts the
reference's call graph implies."""
import numpy as np
import pytest

import app_cases as cases
import plain_model as pm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def app():
    from b200ckks.app import App

    a = App()
    assert a.backend == "engine"
    return a


@pytest.fixture(scope="module")
def cnn_session(app):
    s = app.session(cases.SMALL_LOG_N, cases.CNN_SMALL_BITS, hamming_weight=64, rotation_steps=list(range(1, 2048)))
    yield s
    s.close()


@pytest.mark.parametrize("k,h,w,c,co,st", [(1, 8, 8, 4, 4, 1), (1, 8, 8, 4, 8, 2), (2, 4, 4, 8, 8, 1), (2, 8, 8, 8, 16, 2),
                                            (1, 8, 8, 3, 4, 1)])
def test_conv(cnn_session, k, h, w, c, co, st):
    cases.case_conv(cnn_session, k, h, w, c, co, st)


def test_bn_add_downsample_pool_fc(cnn_session):
    cases.case_bn_add_downsample_pool_fc(cnn_session)


@pytest.mark.parametrize("k,h,w,c,co", [(1, 8, 8, 4, 8), (2, 4, 4, 8, 16)])
def test_conv1x1_stride2_shortcut(cnn_session, k, h, w, c, co):
    cases.case_conv1x1_shortcut(cnn_session, k, h, w, c, co)


def test_relu_small(app):
    s = app.session(cases.SMALL_LOG_N, cases.RELU_BITS, hamming_weight=64)
    cases.case_relu(s)
    st = s.stats()
    # 8 + 8 + 10 + 1 non-scalar multiplications; the one chain of two products in each of the three
    # polynomials is relinearized once (common/func.h: merged_rescale; 27 with $B200CKKS_MERGED_RESCALE=0, asserted in
    # test_switches_gpu.py, and for the reference's own object code in test_dropin_gpu.py)
    assert st["multiply"] == 27 and st["key_switch_relin"] == 24
    s.close()


@pytest.mark.parametrize("logn,real,hoisting", [(9, True, False), (10, False, False), (9, True, True), (10, False, True)])
def test_bootstrap_small(app, logn, real, hoisting):
    """hoisting off = the reference's one-by-one baby rotations; on = one shared decomposition per BSGS stage"""
    s = app.session(cases.SMALL_LOG_N, cases.BOOT_BITS, hamming_weight=64)
    cases.case_bootstrap(s, logn=logn, real=real, hoisting=hoisting)
    s.close()


@pytest.fixture(scope="module")
def big_session(app):
    s = app.session(16, cases.BOOT_BITS, hamming_weight=192)
    yield s
    s.close()


@pytest.mark.parametrize("logn,rot,mulplain", [(14, 76, 299), (13, 69, 235), (12, 62, 171)])
def test_bootstrap_n65536_counts_and_precision(big_session, logn, rot, mulplain):
    s = big_session
    s.stats(reset=True)
    cases.case_bootstrap(s, logn=logn, real=True, tol=1e-4)
    st = s.stats()
    # key switches (LT rotations + SubSum + conjugations + rot(n)) and run-time plaintext multiplies
    assert st["key_switch_rotate"] == rot
    assert st["encode_vector"] == mulplain
    assert st["key_switch_relin"] == 18          # 16 in the degree-59 cosine + 2 double-angle steps




def test_resnet20_end_to_end_matches_model_and_reference_trajectory(big_session):
    """config 1 of BASELINE.json (./cnn 20 10 0 0) on a synthetic image with random-init weights: decrypted logits
    against the float64 model of the same network, and the stage-by-stage (operation, remaining level, scale) trace
    against the reference's committed run log."""
    import json
    import os

    from b200ckks import synthetic

    s = big_session       # same parameter set; its bootstrapping keys are reused
    w = synthetic.random_weights(20, seed=0)
    net = s.resnet(20, w)
    img = synthetic.synthetic_image(0)
    logits, trace = net.infer(img)
    want = pm.resnet_forward(20, w, img)
    # tolerance: each bootstrap returns its input with a ~1e-5 error whose mean over the slots is not zero (a
    # key-dependent offset of up to ~2.5e-5: the fork's *_reduced_error adds overwrite scales that differ by ~1e-6
    # inside EvalMod, evaluator.cpp:316-321, which shifts the evaluated sine slightly); values are carried divided by
    # B = 40, so 18 bootstraps can move a logit by up to ~2e-2.  Measured over several keys: 7e-4 .. 1.2e-2.
    assert np.abs(logits - want).max() < 3e-2
    assert int(np.argmax(logits)) == int(np.argmax(want))
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resnet20_trajectory.json")))
    assert [r["op"] for r in trace] == [r["op"] for r in gold["rows"]]
    for got, ref in zip(trace, gold["rows"]):
        if ref["level"] is not None:
            assert got["level"] == ref["level"], (got, ref)
            assert abs(got["scale"] / ref["scale"] - 1) < 2e-5, (got, ref)       # the log prints 6 significant digits
    # a second image through the split path (encrypted image resident in HBM) agrees with the one-call path
    ct = net.encrypt_image(img)
    out, _ = net.infer_encrypted(ct)
    assert np.abs(net.decrypt_logits(out) - logits).max() < 5e-3      # same keys: only fresh-encryption noise differs
    kb, _ = s.key_residency()
    assert kb < 100 * 2 ** 30            # level-pruned keys fit one B200 (the reference's layout needs 275 GiB)
    # the reference's image loop (one image per OpenMP thread over shared keys, infer_seal.cpp:404): three images, two
    # in flight on their own host threads / CUDA streams, agree with the images run one after the other
    imgs = np.stack([synthetic.synthetic_image(i) for i in range(3)])
    alone = [logits, net.infer(imgs[1], trace=False)[0], net.infer(imgs[2], trace=False)[0]]
    together = net.infer_batch(imgs, 2)
    assert np.abs(together - np.stack(alone)).max() < 5e-3
    cts = [net.encrypt_image(imgs[i]) for i in range(3)]
    outs = net.infer_encrypted_batch(cts, 3)
    assert np.abs(np.stack([net.decrypt_logits(o) for o in outs]) - np.stack(alone)).max() < 5e-3


GPT2_BITS = [49] + [46] * 21 + [49] * 14 + [60]     # gpt2 util.h:22-27 (INIT macro): 37 primes, 60-bit special prime


def test_gpt2_chain_full_slot_bootstrap_and_rotations(app):
    """config 5's fixture (gpt2_ckks/run/run_approx_test.cpp:704-737 SingleBootstrap): INIT parameters, full-slot
    Bootstrapper (logn = 15), bootstrap_3 on a ciphertext switched down to one limb, plus the packed-matmul primitives
    the GPT-2 operators are made of (rotate / multiply_vector / multiply+relin+rescale at that chain)."""
    s = app.session(16, GPT2_BITS, hamming_weight=192)
    rng = np.random.default_rng(5)
    # (a) rotate-and-sum with a mask, the inner loop of row_matrix_multiplication_seal / quickSum (MatrixMul.cpp:118, Fold.cpp:20)
    x, m = rng.uniform(-1, 1, s.slots), rng.uniform(-1, 1, s.slots)
    s.add_rotation_steps([1, 2, 4, 8, 2048])
    ct = s.encrypt(x, 2.0 ** 46, limbs=8)
    acc = ct.clone()
    want = x.copy()
    for st in (1, 2, 4, 8):
        r = acc.clone()
        s.rotate(r, st)
        s.add_reduced_error(acc, r)
        want = want + np.roll(want, -st)
    s.multiply_vector_rescale(acc, m)
    assert np.abs(s.decrypt(acc).real - want * m).max() < 1e-5
    sq = acc.clone()
    s.multiply_relin_rescale(sq, acc)
    assert np.abs(s.decrypt(sq).real - (want * m) ** 2).max() < 1e-4
    # (b) full-slot bootstrapping of a complex message (bootstrap_3 -> bootstrap_full_3)
    z = rng.uniform(-1, 1, s.slots) + 1j * rng.uniform(-1, 1, s.slots)
    boot = s.bootstrapper(15, total_level=35)
    out = boot.bootstrap(s.encrypt(z, 2.0 ** 46, limbs=1), real_message=False)
    size, limbs, scale = out.info()
    assert limbs == 22 and scale == 2.0 ** 46          # 35 levels - 14 consumed: remaining_level 21
    assert np.abs(s.decrypt(out) - z).max() < 2e-4
    s.close()
