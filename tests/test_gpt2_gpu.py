"""GPT-2 operators on the B200 engine through the C ABI (include/b200ckks_app.h: bka_gpt2_call).

(1) The cases of tests/gpt2_cases.py - the same bodies tests/test_gpt2_cpu.py runs on the reference's own SEAL - on a
small ring.  (2) The reference's full-size doctest cases (gpt2_ckks/run/run_approx_test.cpp: N = 2^16, the 37-prime
chain {49, 46 x 21, 49 x 14, 60}) and its microbenchmark shapes, checked against the float64 slot model."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "fhe-gpt-2_b200", "python"))
import gpt2_cases as cases
import gpt2_model as gm

pytestmark = pytest.mark.gpu
SCALE = cases.SCALE


@pytest.fixture(scope="module")
def app():
    from b200ckks.app import App

    a = App()
    assert a.backend == "engine"
    return a


@pytest.fixture(scope="module")
def poly_session(app):
    s = app.session(cases.SMALL_LOG_N, cases.POLY_BITS, hamming_weight=64, rotation_steps=cases.pow2_steps(2048))
    yield s
    s.close()


def test_fold_and_mask(poly_session):
    cases.case_fold_and_mask(poly_session)


def test_polynomials(poly_session):
    cases.case_polynomials(poly_session)


def test_iterations(poly_session):
    cases.case_iterations(poly_session)


def test_max(poly_session):
    cases.case_max(poly_session)


def test_gelu_level_trajectory(poly_session):
    limbs, scale = cases.case_gelu_levels(poly_session)
    assert limbs == 23 - 19 and 0.5 < scale / SCALE < 2.0


def test_col_matmul(poly_session):
    cases.case_col_matmul(poly_session)


def test_row_matmul_small(app):
    steps = cases.row_matmul_steps(2048, 64, 16, 1, 1)
    s = app.session(cases.SMALL_LOG_N, cases.SHORT_BITS, hamming_weight=64, rotation_steps=steps)
    cases.case_row_matmul(s, W_rows=64, rows=16)
    s.close()


# ---- full size: N = 2^16, the INIT macro's chain and rotation-step list --------------------------------------------
@pytest.fixture(scope="module")
def big(app):
    bits, steps = app.gpt2_init_chain()
    extra = cases.row_matmul_steps(32768, 2048, 8, 1, 1)                 # RowMatMul placements
    extra += [-(k * 768 % 32768) for k in range(128)] + [256, 512]       # pack_tight moves
    extra += [-1024, -128, 32768 - 128, 32768 - 64, 32768 - 1024]
    s = app.session(16, bits, hamming_weight=192, rotation_steps=sorted(set(steps + extra) - {0}))
    yield s
    s.close()


def drop_to(sess, ct, limbs):
    sess.mod_switch_to(ct, limbs)
    return ct


def test_full_size_polynomial_vectors(big):
    """SignFunctionF/G, SignFunction, GeluP, Exp, QuickSum of run_approx_test.cpp at the reference's own parameters.
    The inputs are first switched down to the 22 limbs below the bootstrapping primes, as the reference's QuickMax case
    and microbenchmarks do (run_approx_test.cpp:697-699): rescaling a 2^46-scale product by the 49-bit primes of the
    top 14 levels divides the scale by 8 per level, and after four levels nothing is left on either backend."""
    fn = {"sign_f": ("sign_f", []), "sign_g": ("sign_g", []), "sign": ("sign", [2, 2]), "gelu_p": ("gelu_p", []),
          "exp": ("exp", [6])}
    for name, (op, ip) in fn.items():
        v, want = cases.KAT[name]
        _, ct = cases.enc(big, v)
        drop_to(big, ct, 22)
        out, = big.gpt2(op, [ct], i=ip)
        got = big.decrypt(out).real[:len(want)]
        # doctest::Approx(default epsilon): |got - want| < 1.19e-5 * (1 + max(|got|, |want|))
        assert np.all(np.abs(got - want) < 1.19e-5 * (1 + np.maximum(np.abs(got), np.abs(want)))), (name, got)
    _, ct = cases.enc(big, [1, 2, 3, 4, 5, 6, 7, 8] * 2)
    out, = big.gpt2("quickSum", [ct], i=[8])
    assert np.abs(big.decrypt(out).real[:8] - 36.0).max() < 1e-6


def test_full_size_row_matmul_vector(big):
    """RowMatMul of run_approx_test.cpp:233-303: 8 x 2048 ones times its transpose -> 2048 at res[16 i + j]."""
    ones = gm.pack_plain_row(np.ones((8, 2048)))
    ca, cw = big.encrypt(ones[0], SCALE, limbs=6), big.encrypt(ones[0], SCALE, limbs=6)
    bias = big.encrypt(np.full(32768, 1e-7), SCALE)
    out0 = big.gpt2("init_output", i=[1])
    outs = big.gpt2("row_matmul", [ca, cw, bias] + out0, i=[1, 1, 1, 8, 2048, 2048, 8])
    res = big.decrypt(outs[0]).real
    got = np.array([[res[i * 16 + j] for j in range(8)] for i in range(8)])
    assert np.abs(got - 2048.0).max() < 2048 * 1.19e-5
    want = gm.row_matmul([ones[0]], [ones[0]], np.full(32768, 1e-7), [np.zeros(32768)], 2048, 8)[0]
    assert np.abs(res - want).max() < 1e-3


def test_full_size_pack_from_row_and_tight(big):
    """PackFromRow of run_approx_test.cpp:104-170: A[i][j] = 768 i + j packed to fold format, then tight: slot k of
    the three outputs reads k, 32768 + k, 65536 + k."""
    A = np.arange(128 * 768, dtype=float).reshape(128, 768)
    fold = big.gpt2("pack_from_row", i=[128, 768], d=A)
    assert len(fold) == 8
    for ct, want in zip(fold[:2], gm.pack_plain_row(A)[:2]):
        assert np.abs(big.decrypt(ct).real - want).max() < 1e-4
    for ct in fold:
        drop_to(big, ct, 4)
    out0 = big.gpt2("init_output", i=[3])
    tight = big.gpt2("pack_tight", fold + out0)
    got = np.concatenate([big.decrypt(ct).real for ct in tight])
    # values up to 98303 at scale 2^46: absolute error ~1e-4; Approx(98303) tolerates 1.17
    assert np.abs(got - np.arange(3 * 32768)).max() < 1e-2


def test_full_size_smax(big):
    """compute_smax (microbenchmark.cpp:117-148 shape: one ciphertext of 128 x 128 scores in 256-slot chunks, switched
    down to TOTAL_LEVEL - BOOT_LEVEL - 1 = 20 limbs) against the model: exp 7 levels, mask 1, inverse 5, product."""
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, 32768)
    ct = drop_to(big, big.encrypt(x, SCALE), 20)
    out, = big.gpt2("smax", [ct], i=[6, 0])
    want = gm.smax(x, 0)
    got = big.decrypt(out).real
    assert np.abs(got - want).max() < 1e-6
    rows = np.array([got[i * 256:i * 256 + 128].sum() for i in range(128)])
    assert np.abs(rows - 1).max() < 0.15        # four Goldschmidt steps: the model's own distance from a softmax
    assert out.limbs == 20 - 7 - 1 - 5 - 1


def test_full_size_quick_max_with_bootstrap(big):
    """QuickMax of run_approx_test.cpp:674-709: {.1 .. .8} twice, n = 8 -> 0.8 in the first 8 slots; every round ends
    below 18 limbs and is followed by a full-slot bootstrap_3 (logn = 15)."""
    boot = big.bootstrapper(15, total_level=35, final_scale=SCALE)
    x, ct = cases.enc(big, [.1, .2, .3, .4, .5, .6, .7, .8] * 2)
    drop_to(big, ct, 21)
    out, = big.gpt2("quickMax", [ct], i=[8], boot=boot)
    got = big.decrypt(out).real[:8]
    want = gm.quick_max(x, 8)[:8]
    assert np.abs(got - want).max() < 1e-3      # three bootstraps at ~1e-4 each on values <= 0.8
    assert np.abs(got - 0.8).max() < 2e-3       # the reference's expected vector; the composite sign costs ~7e-4
    assert out.limbs >= 18


def test_full_size_attention_projection(big):
    """attn_proj_row_seal at the AttnProjRow shape of run_approx_test.cpp:305-392 (16 x 1024 times 1024 x 16), against
    the model of the operator as written; the per-head accumulation counts follow the reference's index arithmetic."""
    rng = np.random.default_rng(9)
    A1 = np.ones((16, 1024))
    Wt = rng.uniform(-1, 1, (16, 1024))
    a, w = gm.pack_plain_row(A1), gm.pack_plain_row(Wt)
    ca, cw = big.encrypt(a[0], SCALE, limbs=5), big.encrypt(w[0], SCALE, limbs=5)
    bias = big.encrypt(np.zeros(32768), SCALE)
    out0 = big.gpt2("init_output", i=[12])
    outs = big.gpt2("attn_proj_row", [ca, cw, bias] + out0, i=[1, 1, 12, 16, 1024, 1024, 16])
    want, heads = gm.attn_proj([a[0]], [w[0]], np.zeros(32768), [np.zeros(32768)] * 12, 16, 16, False)
    assert heads.sum() == 256
    for o, wv in zip(outs, want):
        assert np.abs(big.decrypt(o).real - wv).max() < 1e-5


def test_full_size_layernorm_as_written(big):
    """compute_layernorm at the microbenchmark shape (16 rows of 768 in 2048-slot chunks, 21 limbs), inputs sized so
    that the sum of squares is near the hard-coded Newton guess 323251."""
    rng = np.random.default_rng(10)
    x = np.zeros(32768)
    for i in range(16):
        x[i * 2048:i * 2048 + 768] = rng.uniform(-0.046, 0.046, 768)
    ct = drop_to(big, big.encrypt(x, SCALE), 21)
    gamma, beta = np.full(768, 0.8), np.full(768, 0.9)
    out, = big.gpt2("layernorm", [ct], i=[768], d=np.concatenate([gamma, beta]))
    want, folded = gm.layernorm_as_written(x, gamma, beta, 768)
    assert 1e5 < folded[0] < 1e6
    got = big.decrypt(out).real
    assert np.abs(got - want).max() < 1e-6 * np.abs(want).max() + 1e-4
