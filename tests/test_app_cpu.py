"""Application layers on the CPU: the host code of fhe-gpt-2_b200/host compiled against the REFERENCE's own SEAL
(oracle/_ref/libapp_ref.so) must reproduce the float64 plaintext model.  This pins the restated Bootstrapper / ReLU /
CNN operators on the reference library itself; tests/test_app_gpu.py then runs the identical cases on the engine."""
import numpy as np
import pytest

import app_cases as cases
import appref
import plain_model as pm

if not appref.available():  # pragma: no cover
    pytest.skip("oracle/_ref/libapp_ref.so not built", allow_module_level=True)


@pytest.fixture(scope="module")
def app():
    return appref.app()


@pytest.fixture(scope="module")
def cnn_session(app):
    s = app.session(cases.SMALL_LOG_N, cases.CNN_SMALL_BITS, hamming_weight=64, rotation_steps=list(range(1, 2048)))
    yield s
    s.close()


def test_evaluation_trees_match_the_reference_dp(app):
    # (m, l) of the reference's trees for degrees {15, 15, 27}: SURVEY.md 8 a19
    heap, depth, m, l = app.oddbaby_tree(15)
    assert (m, l) == (4, 2) and heap == pm.oddbaby_tree(15)[0]
    heap, depth, m, l = app.oddbaby_tree(27)
    assert (m, l) == (5, 3) and heap == pm.oddbaby_tree(27)[0]
    assert [pm.coeff_number(d, pm.oddbaby_tree(d)) for d in (15, 15, 27)] == [16, 16, 28]   # 60 values of d13.txt


def test_minimax_relu_model_precision():
    x = np.linspace(-1, 1, 40001)
    assert np.abs(pm.minimax_relu(x) - np.maximum(x, 0)).max() < 2.0 ** -13


@pytest.mark.parametrize("k,h,w,c,co,st", [(1, 8, 8, 4, 4, 1), (1, 8, 8, 4, 8, 2), (2, 4, 4, 8, 8, 1), (2, 8, 8, 8, 16, 2),
                                            (1, 8, 8, 3, 4, 1)])
def test_conv(cnn_session, k, h, w, c, co, st):
    cases.case_conv(cnn_session, k, h, w, c, co, st)


def test_bn_add_downsample_pool_fc(cnn_session):
    cases.case_bn_add_downsample_pool_fc(cnn_session)


def test_reference_constant_encryptions_give_the_same_layers():
    """$B200CKKS_ENCRYPT_CONSTANTS=1: the reference's own sequence (fresh encryptions of zero / shift / one-half / T0 walked
    down by the reduced-error adds) instead of the default plaintext adds - same values, levels and scales.  The switch
    is read once per process, so the cases run in a child process."""
    import os
    import subprocess
    import sys

    code = ("import sys; sys.path[:0] = [%r, %r]; import appref, app_cases as c; a = appref.app(); "
            "s = a.session(c.SMALL_LOG_N, c.CNN_SMALL_BITS, hamming_weight=64, rotation_steps=list(range(1, 2048))); "
            "c.case_conv(s, 1, 8, 8, 4, 8, 2); c.case_conv(s, 1, 8, 8, 3, 4, 1); c.case_bn_add_downsample_pool_fc(s); s.close(); "
            "s = a.session(c.SMALL_LOG_N, c.RELU_BITS, hamming_weight=64); c.case_relu(s); s.close()"
            % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.abspath(appref.__file__))))
    def run():
        return subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200CKKS_ENCRYPT_CONSTANTS="1"),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)

    r = run()
    # the child competes with this process's Galois keys for host memory: rerun once ONLY if it was killed for memory
    if r.returncode in (-9, 137) or "bad_alloc" in r.stdout or "MemoryError" in r.stdout:
        r = run()
    assert r.returncode == 0, r.stdout[-2000:]


@pytest.mark.parametrize("k,h,w,c,co", [(1, 8, 8, 4, 8), (2, 4, 4, 8, 16)])
def test_conv1x1_stride2_shortcut(cnn_session, k, h, w, c, co):
    cases.case_conv1x1_shortcut(cnn_session, k, h, w, c, co)


def test_relu(app):
    s = app.session(cases.SMALL_LOG_N, cases.RELU_BITS, hamming_weight=64)
    cases.case_relu(s)
    s.close()


def test_bootstrap_sparse_real(app):
    s = app.session(cases.SMALL_LOG_N, cases.BOOT_BITS, hamming_weight=64)
    cases.case_bootstrap(s, logn=9, real=True)
    s.close()
