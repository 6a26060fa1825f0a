"""The REFERENCE's OWN application code as the oracle for rows a15-a21: its cnn_ckks sources compiled
unmodified (oracle/Makefile target cnn_ref -> oracle/_ref/libcnn_ref.so; oracle/ntl_shim stands in for NTL) and run
on the reference's SEAL on the CPU.

* the NTL stand-in is checked against mpmath;
* the reference's layers reproduce the float64 model (so the model used everywhere else is pinned to the reference);
* the restated layers of fhe-gpt-2_b200/host - in their reference-sequence mode - produce ciphertexts that equal the
  reference's LIMB BY LIMB for the same seeded keys and randomness: convolutions, batch norm, residual add across
  levels, down-sampling, pooling, FC, the minimax ReLU and a whole sparse-slot bootstrap (EvalMod fed with the
  polynomial of the reference's own Remez run)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import app_cases as cases
import cnnref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, ".."))

if not cnnref.available():  # pragma: no cover
    pytest.skip("oracle/_ref/libcnn_ref.so not built (needs /root/reference at build time)", allow_module_level=True)


def test_ntl_stand_in_against_mpmath(tmp_path):
    mp = pytest.importorskip("mpmath")
    exe = str(tmp_path / "selftest")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "oracle", "ntl_shim"),
                    os.path.join(ROOT, "oracle", "ntl_shim", "selftest.cpp"), "-o", exe], check=True)
    out = dict(l.split(" ", 1) for l in subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout.splitlines())
    mp.mp.dps = 70
    want = {"pi": mp.pi, "cos1": mp.cos(1), "sin10.5": mp.sin(mp.mpf("10.5")), "cos-37.25": mp.cos(mp.mpf("-37.25")),
            "sqrt2": mp.sqrt(2), "third": mp.mpf(1) / 3, "e": mp.e, "log10": mp.log(10), "pow": mp.sqrt(2), "3^200": mp.mpf(3) ** 200}
    for k, v in want.items():
        got = mp.mpf(out[k])
        assert abs(got - v) <= abs(v) * mp.mpf(10) ** -58, k          # 60 printed digits
    assert out["rounding"].split() == ["-3", "-2", "2", "4", "0", "-7"]   # floor, ceil, round (ties to even), trunc
    assert out["parsed"] == "-0.00125"
    assert out["det"].split() == ["18", "inv00", "0.61111111111111111", "inv12", "-0.11111111111111111"]
    assert out["zz"] == "0 0" and out["identity"] == "ok" and out["hilbert"] == "ok"


@pytest.fixture(scope="module")
def app():
    return cnnref.app()


@pytest.fixture(scope="module")
def cnn_session(app):
    s = app.session(cases.SMALL_LOG_N, cases.CNN_SMALL_BITS, hamming_weight=64, rotation_steps=list(range(1, 2048)))
    yield s
    s.close()


@pytest.mark.parametrize("k,h,w,c,co,st", [(1, 8, 8, 4, 8, 2), (2, 4, 4, 8, 8, 1)])
def test_reference_convolution_matches_the_float_model(cnn_session, k, h, w, c, co, st):
    """multiplexed_parallel_convolution_seal (cnn_seal.cpp:284-530), the reference's object code"""
    cases.case_conv(cnn_session, k, h, w, c, co, st)


def test_reference_bn_add_downsample_pool_fc_match_the_float_model(cnn_session):
    """cnn_seal.cpp:531-787"""
    cases.case_bn_add_downsample_pool_fc(cnn_session)


def test_reference_relu_matches_the_float_model(app):
    """minimax_ReLU_seal (SEALcomp.cpp:3-60) reading ../result/d13.txt as the reference does"""
    with cnnref.workdir():
        s = app.session(cases.SMALL_LOG_N, cases.RELU_BITS, hamming_weight=64)
        cases.case_relu(s)
        s.close()


def test_evaluation_trees_of_the_reference(app):
    import plain_model as pm

    for deg, ml in ((15, (4, 2)), (27, (5, 3))):
        heap, depth, m, l = app.oddbaby_tree(deg)
        assert (m, l) == ml and heap == pm.oddbaby_tree(deg)[0]


def test_restated_layers_equal_the_references_limb_by_limb():
    r = subprocess.run([sys.executable, os.path.join(HERE, "refapp_differential.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]
    report = json.loads(r.stdout.split("REPORT ", 1)[1])
    expected = ["conv_s1", "conv_s2", "conv_k2", "conv_3ch", "bn", "add_mismatched_levels", "downsample", "avgpool", "fc", "relu",
                "boot_rotation_steps", "bootstrap_sparse_real"] + [f"boot_lt_coefficients_{i}" for i in range(6)]
    for k in expected:
        assert report[k]["limbs_equal"], (k, report[k])
        assert report[k]["info"][0] == report[k]["info"][1], k           # size, limbs and scale (to the last bit)
    # the bootstrap of the reference's own code, with the polynomial of its own Remez: |x| <= 1 comes back within 5e-5
    assert report["bootstrap_error"][0] == report["bootstrap_error"][1] < 5e-5
    assert report["bootstrap_sparse_real"]["info"][0] == [2, 17, 2.0 ** 46]


def test_reference_remez_polynomial_against_the_committed_table(app):
    """The EvalMod polynomial shipped in host/ckks_bootstrapping/evalmod_table.inc was made with mpmath because NTL is
    absent; the reference's own multi-interval Remez (common/Remez.cpp, RemezCos.h) now runs on the NTL stand-in.  Both
    must solve the same minimax problem: same error on the 49 intervals to 3 digits."""
    import ctypes as C
    import re

    s = app.session(10, [40, 30, 30, 40], hamming_weight=32)
    boot = s.bootstrapper(7, total_level=2)
    buf, n, sc = np.zeros(64), C.c_int(), C.c_double()
    app.ck(app.L.bkr_evalmod_chebyshev(boot.h, buf.ctypes.data_as(C.POINTER(C.c_double)), 64, C.byref(n), C.byref(sc)))
    assert n.value == 60
    txt = open(os.path.join(ROOT, "fhe-gpt-2_b200", "host", "ckks_bootstrapping", "evalmod_table.inc")).read()
    txt = re.sub(r"/\*.*?\*/", "", re.sub(r"//.*", "", txt))
    vals = [float(v.rstrip("L")) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?L", txt)]
    slope, table = vals[0], np.array(vals[1:61])
    assert abs(sc.value - slope ** 0.25) < 1e-15          # scale_inverse_coeff = c1^(1/2^r), ModularReducer.cpp:45-50
    ref = buf[:60] / sc.value

    def max_err(cheb):
        worst = 0.0
        for k in range(-24, 25):
            x = k + np.linspace(-2.0 ** -10, 2.0 ** -10, 257)
            p = np.polynomial.chebyshev.chebval(x / 25.0, cheb)
            worst = max(worst, np.abs(p - np.cos(2 * np.pi * (x - 0.25) / 4)).max())
        return worst

    e_ref, e_tab = max_err(ref), max_err(table)
    assert 1.8e-10 < e_ref < 2.1e-10 and abs(e_ref - e_tab) < 2e-12, (e_ref, e_tab)
    s.close()
