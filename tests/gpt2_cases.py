"""Test bodies for the GPT-2 operators (fhe-gpt-2_b200/host/gpt2), shared by tests/test_gpt2_cpu.py (the host code on
the reference's own SEAL, oracle/_ref/libapp_ref.so) and tests/test_gpt2_gpu.py (the same host code on the B200 engine
through the C ABI).  Expected values come from the float64 slot model oracle/gpt2_model.py, which
tests/test_gpt2_cpu.py pins to the expected vectors of the reference's own doctest cases (run/run_approx_test.cpp);
level consumption is asserted next to the values.  Tolerances are stated next to each assertion."""
import json
import os

import numpy as np

import gpt2_model as gm

SMALL_LOG_N = 12                            # 2048 slots
POLY_BITS = [49] + [46] * 22 + [60]          # 23 data limbs: one composite sign (16 levels) or a gelu (19) fits
SHORT_BITS = [49] + [46] * 6 + [60]
SCALE = 2.0 ** 46

# expected vectors of the reference's doctest cases (gpt2_ckks/run/run_approx_test.cpp), committed as a fixture
_GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpt2_doctest_vectors.json")))
KAT = {k: (v["input"], v["expected"]) for k, v in _GOLDEN.items() if not k.startswith("_")}


def pow2_steps(slots):
    return [1 << i for i in range(slots.bit_length() - 1)]


def enc(sess, v, limbs=0):
    x = np.zeros(sess.slots)
    x[:len(v)] = v
    return x, sess.encrypt(x, SCALE, limbs=limbs)


def close(sess, ct, want, tol, n=None):
    got = sess.decrypt(ct).real
    n = n or len(want)
    err = np.abs(got[:n] - want[:n]).max()
    assert err < tol, f"max error {err:.3e} >= {tol:.1e}"
    return err


def case_fold_and_mask(sess):
    """quickSum (run_approx_test.cpp:616-640: eight 36s from {1..8, 1..8}) and mask_out."""
    x, ct = enc(sess, [1, 2, 3, 4, 5, 6, 7, 8] * 2)
    out, = sess.gpt2("quickSum", [ct], i=[8])
    assert np.abs(sess.decrypt(out).real[:8] - 36.0).max() < 1e-6
    assert out.limbs == ct.limbs                       # rotations and additions only
    rng = np.random.default_rng(3)
    x, ct = enc(sess, rng.uniform(-1, 1, sess.slots))
    out, = sess.gpt2("quickSum", [ct], i=[16])
    close(sess, out, gm.quick_sum(x, 16), 1e-6)
    out, = sess.gpt2("mask_out", [ct], i=[5, 7])
    close(sess, out, gm.mask_out(x, 5, 7), 1e-7)
    assert out.limbs == ct.limbs - 1


def case_polynomials(sess):
    """sign f / g, the composite sign, the two GELU pieces and exp: the reference's expected vectors, then random
    inputs against the model.  Level use: 4 per sign polynomial, 2 for gelu_p and gelu_q, r + 1 for exp."""
    rng = np.random.default_rng(4)
    levels = {"sign_f": 4, "sign_g": 4, "gelu_p": 2}
    domain = {"sign_f": (-1, 1), "sign_g": (-1, 1), "gelu_p": (-4, -1.95)}
    model = {"sign_f": gm.sign_f, "sign_g": gm.sign_g, "gelu_p": gm.gelu_p}
    for op in ("sign_f", "sign_g", "gelu_p"):
        v, want = KAT[op]
        _, ct = enc(sess, v)
        out, = sess.gpt2(op, [ct])
        # doctest::Approx default: |a - b| < 1.19e-5 * (1 + max(|a|, |b|)); here the CKKS error is ~1e-8 on top of the
        # 10 printed digits of the expected vectors
        close(sess, out, np.array(want), 2e-6 * (1 + np.abs(want).max()))
        assert ct.limbs - out.limbs == levels[op], (op, ct.limbs, out.limbs)
        x, ct = enc(sess, rng.uniform(*domain[op], sess.slots))
        out, = sess.gpt2(op, [ct])
        close(sess, out, model[op](x), 1e-5)
    # compute_gelu_q: the reference adds its last product without a rescale onto an operand of the same level, which
    # replaces the accumulated scale (evaluator.cpp:316-321); the code therefore returns high(x) T4(x) at scale 2^92
    # and not the polynomial of its doctest vector (KAT["gelu_q"], which the model's gelu_q reproduces).  The
    # restatement keeps the reference's sequence, so that is what both backends must show.
    x, ct = enc(sess, rng.uniform(-1.95, 3, sess.slots))
    out, = sess.gpt2("gelu_q", [ct])
    close(sess, out, gm.gelu_q_as_written(x), 1e-5)
    assert ct.limbs - out.limbs == 2 and abs(out.scale / SCALE ** 2 - 1) < 1e-3
    v, want = KAT["sign"]
    _, ct = enc(sess, v)
    out, = sess.gpt2("sign", [ct], i=[2, 2])
    close(sess, out, np.array(want), 1e-5)
    assert ct.limbs - out.limbs == 16
    v, want = KAT["exp"]
    _, ct = enc(sess, v)
    out, = sess.gpt2("exp", [ct], i=[6])
    close(sess, out, np.array(want), 2e-6 * 10848)     # relative 2e-6 on the largest value (e^10)
    assert ct.limbs - out.limbs == 7


def case_iterations(sess):
    """Goldschmidt inverse, the Taylor starting value, and the Newton inverse square root - against the model of the
    C++ sequence.  (The doctest vectors at run_approx_test.cpp:534-587 belong to another normalisation: with the 0.001
    normaliser of IterApprox.cpp:24 eight iterations cannot reach 1 / 0.0035, and taylor_expand has no constant
    term - the model shows the same values as the reference's code, not the comment's.)"""
    rng = np.random.default_rng(5)
    x, ct = enc(sess, rng.uniform(300, 1500, sess.slots))        # 0.001 x in (0.3, 1.5): inside the convergence disc
    out, = sess.gpt2("inverse", [ct], i=[4])
    want = gm.inverse(x, 4)
    close(sess, out, want, 1e-7)
    assert np.abs(want * x - 1).max() < 0.07                     # four iterations: (1 - d)^16 with |1 - d| <= 0.7 -> 0.0033..
    assert ct.limbs - out.limbs == 5                             # one level for 0.001 x, one per iteration
    x, ct = enc(sess, rng.uniform(0.5, 2.0, sess.slots))
    out, = sess.gpt2("taylor", [ct], i=[3], d=[1.3])
    close(sess, out, gm.taylor_expand(x, 1.3), 1e-6)
    x, ct = enc(sess, rng.uniform(0.9, 1.1, sess.slots))
    out, = sess.gpt2("inv_sqrt", [ct], i=[2], d=[4.0])
    close(sess, out, gm.inv_sqrt(x, 2, 4.0), 1e-5)
    assert out.limbs == sess.top_limbs                            # ends with the decrypt / re-encrypt refresh


def case_max(sess):
    """computeMax (run_approx_test.cpp:642-672)."""
    a = [0.1, 0.5, 0.003, 0.4, -0.2]
    b = [0.3, 0.1, 0.1, -0.6, 0.0001]
    xa, ca = enc(sess, a)
    xb, cb = enc(sess, b)
    out, = sess.gpt2("max", [ca, cb])
    close(sess, out, gm.compute_max(xa, xb), 1e-5)
    # the composite sign of a difference of 0.1 / 10 is 0.98, so the reference's expected {0.3, 0.5, 0.1, 0.4, 0.0001}
    # is met to 1e-3 only
    assert np.abs(sess.decrypt(out).real[:5] - np.maximum(a, b)).max() < 1e-3
    assert ca.limbs - out.limbs == 19


def case_gelu_levels(sess):
    """compute_gelu feeds x - 3, x + 1.95 and x + 4 straight into the composite sign, whose domain is [-1, 1]; no x
    satisfies all three, so the reference's output is not a GELU for any input (it has no doctest case either).  What
    is comparable is the level and scale trajectory: 19 levels, scale back at 2^46-ish."""
    x, ct = enc(sess, [0.0])
    out, = sess.gpt2("gelu", [ct])
    assert ct.limbs - out.limbs == 19
    return out.limbs, out.scale


def case_col_matmul(sess):
    """ColMatMul of run_approx_test.cpp:176-231: a = {1,2,3}, b = {4,5,6,4,5,6} -> three rotated Hadamard products."""
    _, ca = enc(sess, [1.0, 2.0, 3.0])
    _, cb = enc(sess, [4.0, 5.0, 6.0, 4.0, 5.0, 6.0])
    outs = sess.gpt2("col_matmul", [ca, cb], i=[1, 3])
    expected = [[4.0, 10.0, 18.0], [5.0, 12.0, 12.0], [6.0, 8.0, 15.0]]
    assert len(outs) == 3
    for o, want in zip(outs, expected):
        close(sess, o, np.array(want), 1e-6)
        assert o.info()[0] == 3            # the accumulators are never relinearised in the reference


def row_matmul_steps(slots, W_rows, W_cols, n_left, n_weights):
    """Every rotation step row_matrix_multiplication_seal issues for these shapes (the reference relies on SEAL's
    power-of-two fallback; explicit keys keep the small CPU case cheap)."""
    wr, wc = gm.round_to_2(W_rows), gm.round_to_2(W_cols)
    chunk, out_chunk = wr * 2, wc * 2
    num_chunks = slots // chunk
    steps = set(pow2_steps(slots)) | {slots - wr}
    for rots in range(num_chunks):
        steps.add(rots * chunk)
    for i in range(n_left):
        for j in range(n_weights):
            for rots in range(num_chunks):
                for pos in range(num_chunks):
                    row = i * num_chunks + pos
                    col = j * num_chunks + ((rots + pos) % num_chunks)
                    cchunk = ((row * out_chunk) % slots) // out_chunk
                    steps.add(-(cchunk * out_chunk + col - pos * chunk))
    steps.discard(0)
    return sorted(steps)


def case_row_matmul(sess, W_rows, rows):
    """row_matrix_multiplication_seal (RowMatMul of run_approx_test.cpp:233-303 uses 8 x 2048 ones at 32768 slots):
    A (rows x W_rows) times W^T (rows x W_rows), both in fold format in one ciphertext, random entries."""
    slots = sess.slots
    rng = np.random.default_rng(6)
    A = rng.uniform(-1, 1, (rows, W_rows))
    W = rng.uniform(-1, 1, (rows, W_rows))
    a = gm.pack_plain_row(A, slots)
    w = gm.pack_plain_row(W, slots)
    assert a.shape[0] == 1 and w.shape[0] == 1
    ca = sess.encrypt(a[0], SCALE)
    cw = sess.encrypt(w[0], SCALE)
    bias = sess.encrypt(np.full(slots, 1e-7), SCALE)
    out0 = sess.gpt2("init_output", i=[1])
    outs = sess.gpt2("row_matmul", [ca, cw, bias] + out0, i=[1, 1, 1, rows, W_rows, W_rows, rows])
    want = gm.row_matmul([a[0]], [w[0]], np.full(slots, 1e-7), [np.zeros(slots)], W_rows, rows, slots)[0]
    close(sess, outs[0], want, 1e-5, n=slots)
    # and the model itself against plain linear algebra: element (i, j) of A W^T sits at i * 2 * round_to_2(rows) + j
    oc = 2 * gm.round_to_2(rows)
    got = np.array([[want[i * oc + j] for j in range(rows)] for i in range(rows)])
    assert np.abs(got - A @ W.T).max() < 1e-6
    assert ca.limbs - outs[0].limbs == 2        # product rescale + mask rescale
