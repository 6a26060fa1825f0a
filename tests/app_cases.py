"""Test bodies for the application layers (bootstrapping, ReLU, multiplexed CNN operators), shared by
tests/test_app_cpu.py (the host code running on the reference's own SEAL, oracle/_ref/libapp_ref.so) and
tests/test_app_gpu.py (the same host code on the B200 engine through the C ABI).  Expected values come from the float64
plaintext model oracle/plain_model.py; tolerances are stated next to each assertion."""
import numpy as np

import plain_model as pm

SMALL_LOG_N = 12                       # 2048 slots: small tensors, seconds on one CPU core
CNN_SMALL_BITS = [51] + [46] * 6 + [51]
RELU_BITS = [51] + [46] * 16 + [51]    # the 16 "remaining levels" of infer_seal.cpp:300 without the boot levels
BOOT_BITS = [51] + [46] * 16 + [51] * 14 + [51]   # the full CNN chain (infer_seal.cpp:306-311)


def fit_copies(k, h, w, t, n):
    p = 1
    while k * k * h * w * t * p * 2 <= n:
        p *= 2
    return p


def make_tensor(sess, rng, k, h, w, c, limbs, mag=0.25):
    n = sess.slots
    t = -(-c // (k * k))
    p = fit_copies(k, h, w, t, n)
    x = rng.uniform(-mag, mag, (c, h, w))
    ct = sess.encrypt(pm.pack(x, k, p, n), 2.0 ** 46, limbs=limbs)
    return x, ct, [k, h, w, c, t, p, sess.log_n - 1]


def unpack_all(sess, ct, parms):
    k, h, w, c, t, p, _ = parms
    y = sess.decrypt(ct).real
    return [pm.unpack(y, k, h, w, c, copy=cp, p=p) for cp in range(p)]


def case_conv(sess, k, h, w, c, co, st, seed=1):
    rng = np.random.default_rng(seed)
    x, ct, parms = make_tensor(sess, rng, k, h, w, c, limbs=4)
    wt = rng.normal(0, 0.3, 9 * c * co)
    var, bw = rng.uniform(0.5, 1.5, co), rng.uniform(0.5, 1.0, co)
    out, op = sess.conv(ct, parms, co, st, wt, var, bw)
    want = pm.conv_bn_scale(x, wt, var, bw, st)
    assert op[:4] == [k * st, h // st, w // st, co]
    assert out.limbs == 2          # two rescales: 4 -> 2 limbs (conv: "remaining level" drops by 2)
    for got in unpack_all(sess, out, op):
        assert np.abs(got - want).max() < 1e-6      # fresh-ciphertext noise ~1e-8 at scale 2^46


def case_bn_add_downsample_pool_fc(sess, seed=2):
    rng = np.random.default_rng(seed)
    # batch norm shift
    x, ct, parms = make_tensor(sess, rng, 1, 8, 8, 8, limbs=3)
    c = 8
    bias, mean = rng.normal(0, 0.3, c), rng.normal(0, 0.3, c)
    var, bw = rng.uniform(0.5, 1.5, c), rng.uniform(0.5, 1.0, c)
    out = sess.bn(ct, parms, bias, mean, var, bw)
    want = pm.bn_shift(x, bias, mean, var, bw)
    for got in unpack_all(sess, out, parms):
        assert np.abs(got - want).max() < 1e-6
    # residual add
    x2, ct2, _ = make_tensor(sess, rng, 1, 8, 8, 8, limbs=3)
    s = sess.tensor_add(ct, ct2)
    for got in unpack_all(sess, s, parms):
        assert np.abs(got - (x + x2)).max() < 1e-6
    # down-sampling shortcut (needs t % 8 == 0)
    xd, ctd, pd = make_tensor(sess, rng, 1, 8, 8, 8, limbs=3)
    out, op = sess.downsample(ctd, pd)
    assert op[:5] == [2, 4, 4, 16, 4]
    want = pm.downsample(xd)
    for got in unpack_all(sess, out, op):
        assert np.abs(got - want).max() < 1e-6
    # average pooling (k = 2 input as after the network's last stage) and the fully connected layer
    xa, cta, pa = make_tensor(sess, rng, 2, 4, 4, 16, limbs=3)
    out, op = sess.avgpool(cta, pa, B=40.0)
    pooled = pm.avgpool(xa)
    got = sess.decrypt(out).real[:16]
    assert np.abs(got - pooled).max() < 1e-5
    W = rng.normal(0, 0.3, (10, 16))
    logits = sess.fc(out, op, W, np.zeros(10), 10, 16)
    got = sess.decrypt(logits).real[:10]
    assert np.abs(got - W @ pooled).max() < 1e-4


def case_relu(sess, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, sess.slots)
    x[:5] = [-1.0, 1.0, 0.0, 2.0 ** -10, -2.0 ** -10]
    ct = sess.encrypt(x, 2.0 ** 46)
    assert ct.limbs == 17
    out = sess.relu(ct)
    assert out.limbs == 3                      # 14 levels: 16 -> 2 (result/resnet20_cifar10_image0.txt:14-15)
    y = sess.decrypt(out).real
    assert np.abs(y - pm.minimax_relu(x)).max() < 1e-6     # same polynomial as the float model
    assert np.abs(y - np.maximum(x, 0)).max() < 2.0 ** -13  # the alpha = 13 guarantee (run_compare.cpp: ShowFailure_ReLU)


def case_bootstrap(sess, logn, real=True, seed=4, tol=5e-5, hoisting=None):
    rng = np.random.default_rng(seed)
    n = 1 << logn
    if real:
        x = rng.uniform(-1, 1, n)
    else:
        x = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    xs = np.tile(x, sess.slots // n)
    boot = sess.bootstrapper(logn)
    if hoisting is not None:
        boot.set_hoisting(hoisting)
    ct = sess.encrypt(xs, 2.0 ** 46, limbs=1)
    out = boot.bootstrap(ct, real_message=real)
    size, limbs, scale = out.info()
    assert (size, limbs) == (2, 17)             # "remaining level : 16" after every bootstrap of the reference log
    assert scale == 2.0 ** 46                   # scale forced to final_scale (Bootstrapper.cpp:3232)
    y = sess.decrypt(out)
    err = np.abs(y - xs)
    assert err.max() < tol, err.max()           # reference log: ~1e-7 on values ~3e-2; here |x| <= 1
    return err.max()


def case_conv1x1_shortcut(sess, k, h, w, c, co, seed=7):
    """The CIFAR-100 shortcut: 1x1 stride-2 multiplexed convolution with folded batch-norm scale (infer_seal.cpp:826-829)."""
    rng = np.random.default_rng(seed)
    x, ct, parms = make_tensor(sess, rng, k, h, w, c, limbs=4)
    wt = rng.normal(0, 0.3, c * co)
    var, bw = rng.uniform(0.5, 1.5, co), rng.uniform(0.5, 1.0, co)
    out, op = sess.conv(ct, parms, co, 2, wt, var, bw, fh=1, fw=1)
    g = bw / np.sqrt(var + 1e-5)
    want = pm.conv1x1(x, wt, 2) * g[:, None, None]
    assert op[:4] == [k * 2, h // 2, w // 2, co]
    for got in unpack_all(sess, out, op):
        assert np.abs(got - want).max() < 1e-6
