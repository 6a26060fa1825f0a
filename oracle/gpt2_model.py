"""Float64 slot-level model of the reference's GPT-2 operators - TEST INFRASTRUCTURE ONLY.

Each function restates what the corresponding C++ operator of gpt2_ckks/gpt2-ckks/single-key/gpt2 computes on the
decrypted slot vector, call for call (rotations become np.roll, ciphertext products become element-wise products,
rescales and level bookkeeping disappear).  It follows the C++ - not the numpy sketches in /root/reference/plain_approx,
which differ from the C++ in loop counts and in a few formulas - and is pinned to the expected vectors of the
reference's own doctest cases (run/run_approx_test.cpp) by tests/test_gpt2_cpu.py.

Only tests/ may import this module; the product path never does.
"""
import math

import numpy as np

SLOTS = 32768


def rot(v, k):
    """Evaluator::rotate_vector(ct, k): slot i receives slot i + k."""
    return np.roll(v, -k)


def round_to_2(x):
    return int(2 ** math.ceil(math.log2(x)))


# ---- Fold.cpp -----------------------------------------------------------------------------------------------------
def quick_sum(v, n):
    """Fold.cpp:21-46."""
    out = v + rot(v, 1)
    stride, i = 2, 0
    while i < math.log2(n) - 1:
        out = out + rot(out, stride)
        stride *= 2
        i += 1
    return out


def compute_max(a, b):
    """Fold.cpp:49-86."""
    diff = a - b
    s = sign_function(0.1 * diff, 2, 2)
    return 0.5 * (diff * s + a + b)


def quick_max(v, n):
    """Fold.cpp:89-107 (the bootstraps are identity on the slots)."""
    cur, stride, i = v.copy(), 1, 0
    while i < math.log2(n):
        cur = compute_max(cur, rot(cur, stride))
        stride *= 2
        i += 1
    return cur


# ---- PolyApprox.cpp -----------------------------------------------------------------------------------------------
def cheby_basis(x, n):
    """PolyApprox.cpp:15-101: [T0, T1, T2, T3, T4, T8, ...]."""
    t2 = 2 * x * x - 1
    t3 = (2 * x) * t2 - x
    basis = [np.ones_like(x), x, t2, t3]
    p = t2
    for _ in range(n - 2):
        p = 2 * p * p - 1
        basis.append(p)
    return basis


SIGN_F = (-0.6767578125, 1.563049316, -0.02685546875, 0.1384277344, 0.002136230469)
SIGN_G = (-1.121704102, 1.978370667, -0.6178588867, 0.403533935, 0.3557052612)


def sign_poly(c, x):
    q1, r1, q2t3, q2x, q3 = c
    b = cheby_basis(x, 4)
    return (q1 * x) * b[2] + r1 * x + (q2t3 * b[3] + q2x * x) * b[4] + (q3 * x) * b[5]


def sign_f(x):
    return sign_poly(SIGN_F, x)


def sign_g(x):
    return sign_poly(SIGN_G, x)


def sign_function(x, df, dg):
    for _ in range(dg // 2):
        x = sign_g(sign_g(x))
    for _ in range(df // 2):
        x = sign_f(sign_f(x))
    return x


def gelu_p(x):
    b = cheby_basis(x, 2)
    return (-0.005337069175 * x + -0.05745879353) * b[2] + (-0.4187418723 * x + -0.55528939)


def gelu_q(x):
    b = cheby_basis(x, 4)
    out = (-0.00324699876 * x + 0.1634058825) * b[2] + (0.5027208006 * x + 0.1750485092)
    high = 0.0002609111473 * x + 0.0001533078376 * (x * x) + -0.004401064777
    return out + high * b[4]


def gelu_q_as_written(x):
    """What compute_gelu_q returns in the reference (PolyApprox.cpp:347-409): the last product (high * T4) is added
    without a rescale; both operands are on the same level, so add_inplace_reduced_error overwrites the accumulated
    scale 2^46 by 2^92 (evaluator.cpp:316-321) and the low-order part shrinks by 2^-46.  gelu_q above is the
    polynomial the reference's doctest vector describes."""
    b = cheby_basis(x, 4)
    out = (-0.00324699876 * x + 0.1634058825) * b[2] + (0.5027208006 * x + 0.1750485092)
    high = 0.0002609111473 * x + 0.0001533078376 * (x * x) + -0.004401064777
    return out * 2.0 ** -46 + high * b[4]


def gelu(x):
    s2 = 0.5 * sign_function(x - 3.0, 2, 2)
    s1 = 0.5 * sign_function(x + 1.95, 2, 2)
    s0 = 0.5 * sign_function(x + 4.0, 2, 2)
    b1, b2, b3 = s0 - s1, s1 - s2, 0.5 * s2
    return b1 * gelu_p(x) + b2 * gelu_q_as_written(x) + b3 * x


def exp(x, r):
    out = x / 2.0 ** r + 1
    for _ in range(r):
        out = out * out
    return out


def padding_mask(on_scores, on_padding, slots=SLOTS):
    m = np.full(slots, float(on_scores))
    for i in range(128):
        m[i * 256 + 128:i * 256 + 256] = on_padding
    return m


def smax(x, gamma):
    """PolyApprox.cpp:577-634 (gamma is an int in the reference's signature)."""
    x = x + padding_mask(0.0, -float(int(gamma)))
    exps = exp(x, 6) * padding_mask(1.0, 0.0)
    rolled = rot(exps, SLOTS - 128) + exps
    summed = quick_sum(rolled, 128)
    return exps * inverse(summed, 4)


# ---- IterApprox.cpp -----------------------------------------------------------------------------------------------
def inverse(x, iters):
    """IterApprox.cpp:16-57."""
    n = np.full_like(x, 0.001)
    d = 0.001 * x
    for _ in range(iters):
        f = 2.0 - d
        n = n * f
        d = d * f
    return n


def taylor_expand(x, guess):
    """IterApprox.cpp:70-122."""
    coeffs = [-0.5, -0.5 * -1.5, -2.5 * -1.5 * -0.5]
    powers = [-1.5, -2.5, -3.5]
    fact, total = 1, None
    for i in range(3):
        c = coeffs[i] * 1 / fact
        base = x * guess ** (powers[i] / (i + 1))
        p = base
        for _ in range(i):
            p = p * base
        p = p * c
        total = p if total is None else total + p
        fact *= i + 2
    return total


def inv_sqrt(x, iters, guess):
    """IterApprox.cpp:131-171."""
    y = taylor_expand(x, guess)
    mhx = -0.5 * x
    for _ in range(iters):
        y = y * ((y * y) * mhx + 1.5)
    return y


# ---- MatrixMul.cpp ------------------------------------------------------------------------------------------------
def mask_out(v, start, length):
    m = np.zeros_like(v)
    m[start:start + length] = 1.0
    return v * m


def pack_plain_row(mat, slots=SLOTS):
    rows, cols = mat.shape
    chunk = round_to_2(cols) * 2
    n = max(1, rows * chunk // slots)
    out = np.zeros(n * slots)
    for i in range(rows):
        out[i * chunk:i * chunk + cols] = mat[i]
    return out.reshape(n, slots)


def row_matmul(left, weights, bias, outputs, W_rows, W_cols, slots=SLOTS):
    """MatrixMul.cpp:124-193 on lists of slot vectors; returns the new outputs."""
    wr, wc = round_to_2(W_rows), round_to_2(W_cols)
    chunk, out_chunk = wr * 2, wc * 2
    num_chunks = slots // chunk
    outputs = [o.copy() for o in outputs]
    for i, a in enumerate(left):
        for j, w in enumerate(weights):
            for rots in range(num_chunks):
                prod = a * rot(w, rots * chunk)
                prod = prod + rot(prod, slots - wr)
                folded = quick_sum(prod, wr)
                for pos in range(num_chunks):
                    row = i * num_chunks + pos
                    col = j * num_chunks + ((rots + pos) % num_chunks)
                    piece = mask_out(folded, pos * chunk, 1)
                    idx = (row * out_chunk) // slots
                    cchunk = ((row * out_chunk) % slots) // out_chunk
                    shift = cchunk * out_chunk + col - pos * chunk
                    outputs[idx] += rot(piece, -shift)
    return [o + bias for o in outputs]


def col_matmul(left, right, cols):
    """MatrixMul.cpp:26-112."""
    right = [r.copy() for r in right]
    outs = []
    for _ in range(cols):
        acc = np.zeros_like(left[0])
        for a, b in zip(left, right):
            acc = acc + a * b
        outs.append(acc)
        right = [rot(b, 1) for b in right]
    return outs


def pack_tight(inputs, outputs, slots=SLOTS):
    """pack.cpp:9-57."""
    inputs = [v.copy() for v in inputs]
    outputs = [v.copy() for v in outputs]
    written = 0
    for i in range(8):
        straddled = False
        for j in range(16):
            if (j == 15 and straddled) or written == 98304:
                break
            outputs[written // slots] += rot(mask_out(inputs[i], 0, 768), -(written % slots))
            written += 768
            inputs[i] = rot(inputs[i], 2048)
            leftover = slots - written % slots
            if leftover < 768:
                outputs[written // slots] += rot(mask_out(inputs[i], 0, leftover), -(written % slots))
                written += leftover
                outputs[written // slots] += rot(mask_out(inputs[i], leftover, 768 - leftover), leftover)
                written += 768 - leftover
                inputs[i] = rot(inputs[i], 2048)
                straddled = True
    return outputs


def attn_proj(left, weights, bias, outputs, A_rows, W_cols, column_layout, slots=SLOTS):
    """attn_proj_row_seal / attn_proj_col_seal as written (MatrixMul.cpp:243-466): the activations never enter, every
    working ciphertext uses the chunk-0 mask, and the placement rotation is by 0."""
    outputs = [o.copy() for o in outputs]
    heads = np.zeros(len(outputs), dtype=int)
    for i in range(len(left)):
        for j, w in enumerate(weights):
            working = quick_sum(rot(mask_out(w, 0, 1), -1024), 1024)
            for rots in range(16):
                for pos in range(16):
                    row, col = i * 16 + pos, j * 16 + ((rots + pos) % 16)
                    abs_pos = row * (768 if column_layout else W_cols) + col
                    head = (abs_pos // 64) % 12
                    outputs[head] += mask_out(working, pos * 2048, 1)
                    heads[head] += 1
    return [o + bias for o in outputs], heads


def layernorm_as_written(x, gamma, beta, row_size, slots=SLOTS):
    """compute_layernorm as written (IterApprox.cpp:173-252): the inverse square root is computed and dropped; the
    result is mask * z^2 * z * gamma * sqrt(row_size) + beta (beta untiled, in the first row only)."""
    rr = round_to_2(row_size)
    mask, mul = np.zeros(slots), np.zeros(slots)
    g = np.asarray(gamma) * math.sqrt(row_size)
    for i in range(16):
        mask[i * 2 * rr:i * 2 * rr + rr] = 1.0
        mul[i * 2 * rr:i * 2 * rr + len(g)] = g
    folded = quick_sum(rot(x, -rr) + x, rr)
    z = row_size * x - folded
    y = (z * z) * mask
    folded2 = quick_sum(rot(y, slots - rr) + y, rr)
    y = y * z * mul
    b = np.zeros(slots)
    b[:len(beta)] = beta
    return y + b, folded2
