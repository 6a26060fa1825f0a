"""TEST INFRASTRUCTURE ONLY - float64 plaintext model of the reference's encrypted ResNet (cnn_ckks).

What each encrypted operator of cnn_ckks/cpu-ckks/single-key/cnn/cnn_seal.cpp computes on the packed slots, stated on
ordinary (C, H, W) arrays, so that decrypted results of the engine can be compared with a tolerance:

  pack / unpack         multiplexed packing, cnn_seal.cpp:358,395 (slot = k^2 h w (c // k^2) + k w (k y + (c % k^2) // k)
                        + k x + c % k, p copies with period n / p)
  conv_bn_scale         multiplexed_parallel_convolution_seal :284-530: 3x3 convolution, zero padding 1, stride st,
                        times weight / sqrt(var + eps) per output channel (folded into the select masks, :379)
  bn_shift              multiplexed_parallel_batch_norm_seal :531-576: subtract (mean w / sqrt(var + eps) - bias) / B
  downsample            multiplexed_parallel_downsampling_seal :610-679: x[:, ::2, ::2] with c/2 zero channels either side
  avgpool / fc          averagepooling_seal_scale :680-746 (times B / (h w)), matrix_multiplication_seal :747-787 (no bias:
                        the reference builds the bias vector but never adds it)
  minimax_relu          comp/SEALcomp.cpp:3-60 + SEALfunc.cpp:59-193: the alpha = 13 composite polynomial evaluated over
                        the same evaluation trees in float64 (approximation error of the network's ReLU, not just ReLU)
  resnet_forward        ResNet_cifar10_seal_sparse, infer_seal.cpp:404-575 (bootstrapping is the identity on values)
"""
import math
import os

import numpy as np

B = 40.0
EPS = 1e-5


# ---------------------------------------------------------------------------------------------- packing
def slot_index(k, h, w, c):
    """(c, h, w) integer array of slot positions inside one copy."""
    ch, y, x = np.meshgrid(np.arange(c), np.arange(h), np.arange(w), indexing="ij")
    return k * k * h * w * (ch // (k * k)) + k * w * (k * y + (ch % (k * k)) // k) + k * x + ch % k


def pack(t, k, p, n):
    c, h, w = t.shape
    v = np.zeros(n)
    idx = slot_index(k, h, w, c)
    for copy in range(p):
        v[copy * (n // p) + idx] = t
    return v


def unpack(v, k, h, w, c, copy=0, p=1):
    n = len(v)
    return np.asarray(v)[copy * (n // p) + slot_index(k, h, w, c)]


# ------------------------------------------------------------------------------------------------ layers
def conv3x3(x, weight, stride):
    """x: (ci, h, w); weight: flat in the reference's order data[9 ci oc + 9 ic + 3 dy + dx]."""
    ci, h, w = x.shape
    co = weight.size // (9 * ci)
    W = np.asarray(weight, dtype=np.float64).reshape(co, ci, 3, 3)
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1)))
    out = np.zeros((co, h, w))
    for dy in range(3):
        for dx in range(3):
            out += np.einsum("oi,ihw->ohw", W[:, :, dy, dx], xp[:, dy:dy + h, dx:dx + w])
    return out[:, ::stride, ::stride]


def conv_bn_scale(x, weight, running_var, bn_weight, stride, eps=EPS):
    s = np.asarray(bn_weight) / np.sqrt(np.asarray(running_var) + eps)
    return conv3x3(x, weight, stride) * s[:, None, None]


def bn_shift(x, bias, mean, var, weight, eps=EPS, b=B):
    g = (np.asarray(mean) * np.asarray(weight) / np.sqrt(np.asarray(var) + eps) - np.asarray(bias)) / b
    return x - g[:, None, None]


def downsample(x):
    c = x.shape[0]
    out = np.zeros((2 * c, x.shape[1] // 2, x.shape[2] // 2))
    out[c // 2:c // 2 + c] = x[:, ::2, ::2]
    return out


def avgpool(x, b=B):
    return x.mean(axis=(1, 2)) * b


def fc(x, matrix, q=10, r=64):
    return np.asarray(matrix, dtype=np.float64).reshape(q, r) @ x


# ---------------------------------------------------------------------------------------- minimax ReLU
def _pow2(n):
    return 1 << n


def oddbaby_tree(n):
    """upgrade_oddbaby (comp/program.cpp:3-60): returns (heap list, depth, m, l)."""
    d = math.ceil(math.log(n) / math.log(2.0) - 1e-12)
    INF = 10000
    best = (INF, None, 0, 0)

    def merge(rem, quo, g):
        depth = max(rem[1], quo[1]) + 1
        tree = [-1] * _pow2(depth + 1)
        tree[1] = g
        for sub, side in ((rem, 0), (quo, 1)):
            for i in range(1, _pow2(sub[1] + 1)):
                row = 1 << (i.bit_length() - 1)
                tree[i + (side + 1) * row] = sub[0][i]
        return (tree, depth)

    l = 1
    while _pow2(l) - 1 <= n:
        m = 1
        while _pow2(m - 1) < n:
            leaf = ([-1, 0], 0)
            f = [[0] * (d + 1) for _ in range(n + 1)]
            G = [[leaf] * (d + 1) for _ in range(n + 1)]
            for i in range(3, n + 1, 2):
                f[i][1] = INF
            for j in range(2, d + 1):
                for i in range(1, n + 1, 2):
                    if i <= _pow2(l) - 1 and i <= _pow2(j - 1):
                        f[i][j] = 0
                        continue
                    mn, mt = INF, ([-1, 0], 0)
                    k = 1
                    while k <= m - 1 and _pow2(k) < i and k < j:
                        g = _pow2(k)
                        c = f[i - g][j - 1] + f[g - 1][j] + 1
                        if c < mn:
                            mn, mt = c, merge(G[g - 1][j], G[i - g][j - 1], g)
                        k += 1
                    f[i][j], G[i][j] = mn, mt
            total = f[n][d] + _pow2(l - 1) + m - 2
            if total < best[0]:
                best = (total, G[n][d], m, l)
            m += 1
        l += 1
    return best[1][0], best[1][1], best[2], best[3]


def eval_decomposed(x, deg, coeff, tree):
    """eval_polynomial_integrate, odd-baby branch (comp/SEALfunc.cpp:59-193), on floats."""
    heap, depth = tree[0], tree[1]
    nodes = _pow2(depth + 1)
    degree = [-1] * nodes
    degree[1] = deg
    for j in range(2, nodes):
        g = heap[j // 2]
        degree[j] = g - 1 if j % 2 == 0 else degree[j // 2] - g
    first, cur = {}, 1
    for j in range(1, nodes):
        if heap[j] == 0:
            first[j] = cur
            cur += degree[j] + 1
    T = {0: np.ones_like(x), 1: x}

    def cheb(k):
        if k not in T:
            T[k] = 2 * x * cheb(k - 1) - cheb(k - 2)
        return T[k]

    def node(j):
        if heap[j] == 0:
            idx = first[j]
            acc = np.zeros_like(x)
            for k in range(1, degree[j] + 1, 2):
                acc = acc + coeff[idx] * cheb(k)
                idx += 2
            return acc
        return cheb(heap[j]) * node(2 * j + 1) + node(2 * j)

    return node(1)


def coeff_number(deg, tree):
    heap, depth = tree[0], tree[1]
    nodes = _pow2(depth + 1)
    degree = [-1] * nodes
    degree[1] = deg
    for j in range(2, nodes):
        g = heap[j // 2]
        degree[j] = g - 1 if j % 2 == 0 else degree[j // 2] - g
    return sum(degree[j] + 1 for j in range(nodes) if heap[j] == 0)


_D13 = None


def d13_coefficients():
    """the alpha = 13 table that ships with the engine (host/comp/minimax_relu_alpha13.inc)"""
    global _D13
    if _D13 is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fhe-gpt-2_b200", "host", "comp",
                            "minimax_relu_alpha13.inc")
        _D13 = [float(l.strip().rstrip(",")) for l in open(path) if l.strip() and not l.startswith("//")]
    return _D13


def minimax_relu(x, degs=(15, 15, 27), scaled_val=1.7):
    x = np.asarray(x, dtype=np.float64)
    table, pos, y = d13_coefficients(), 0, x
    trees = [oddbaby_tree(d) for d in degs]
    for i, (d, tr) in enumerate(zip(degs, trees)):
        cnt = coeff_number(d, tr)
        rng = (scaled_val if i + 1 == len(degs) - 1 else 2.0) if i + 1 < len(degs) else 2.0
        c = [v / rng for v in table[pos:pos + cnt]]
        pos += cnt
        y = eval_decomposed(y, d, c, tr)
    return (y + 0.5) * x


# ------------------------------------------------------------------------------------------- the network
import sys as _sys

_sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fhe-gpt-2_b200", "python")))
from b200ckks.synthetic import random_weights, resnet_shapes, synthetic_image  # noqa: E402,F401  (shared synthetic inputs)


def conv1x1(x, weight, stride):
    """x: (ci, h, w); weight: flat data[ci oc + ic] (the reference's order with fh = fw = 1)."""
    ci = x.shape[0]
    W = np.asarray(weight, dtype=np.float64).reshape(-1, ci)
    return np.einsum("oi,ihw->ohw", W, x)[:, ::stride, ::stride]


def resnet_forward(layer_num, w, image, relu=minimax_relu, collect=None):
    """Values are carried divided by B exactly like the ciphertext slots.  collect: optional list receiving
    (op, array) after every operation in the order of the engine's trace.  Weights with "shortcut_weight" select the
    CIFAR-100 network (B = 65, 1x1 stride-2 shortcut convolutions + batch norm, 100 logits; infer_seal.cpp:585-891)."""
    end_num, _ = resnet_shapes(layer_num)
    cifar100 = "shortcut_weight" in w
    B = 65.0 if cifar100 else 40.0
    x = np.asarray(image, dtype=np.float64).reshape(3, 32, 32) / B
    log = (lambda op, a: collect.append((op, a.copy()))) if collect is not None else (lambda op, a: None)

    def conv(stage, st):
        nonlocal x
        x = conv_bn_scale(x, w["conv_weight"][stage], w["bn_var"][stage], w["bn_weight"][stage], st)
        log("conv", x)

    def bn(stage):
        nonlocal x
        x = bn_shift(x, w["bn_bias"][stage], w["bn_mean"][stage], w["bn_var"][stage], w["bn_weight"][stage], b=B)
        log("bn", x)

    def act():
        nonlocal x
        x = relu(x)
        log("relu", x)

    conv(0, 1)
    bn(0)
    act()
    for j in range(3):
        for k in range(end_num + 1):
            stage = 2 * ((end_num + 1) * j + k) + 1
            short = x
            conv(stage, 2 if (j >= 1 and k == 0) else 1)
            bn(stage)
            log("bootstrap", x)
            act()
            conv(stage + 1, 1)
            bn(stage + 1)
            if j >= 1 and k == 0 and cifar100:
                sc = j - 1
                g = np.asarray(w["shortcut_bn_weight"][sc]) / np.sqrt(np.asarray(w["shortcut_bn_var"][sc]) + EPS)
                short = conv1x1(short, w["shortcut_weight"][sc], 2) * g[:, None, None]
                log("conv", short)
                short = bn_shift(short, w["shortcut_bn_bias"][sc], w["shortcut_bn_mean"][sc], w["shortcut_bn_var"][sc],
                                 w["shortcut_bn_weight"][sc], b=B)
                log("bn", short)
            elif j >= 1 and k == 0:
                short = downsample(short)
                log("downsample", short)
            x = short + x
            log("add", x)
            log("bootstrap", x)
            act()
    pooled = avgpool(x, b=B)
    log("avgpool", pooled)
    logits = fc(pooled, w["linear_weight"], q=100 if cifar100 else 10)
    log("fc", logits)
    return logits
