"""Child process of tests/test_refapp_cpu.py::test_restated_layers_equal_the_references_limb_by_limb.

Runs the same seeded program twice on the reference's SEAL (CPU): once through the REFERENCE's own application code
(oracle/_ref/libcnn_ref.so) and once through this repo's restated layers in their reference-sequence mode
(oracle/_ref/libapp_ref.so with B200CKKS_ENCRYPT_CONSTANTS=1, B200CKKS_NO_HOIST=1), with B200CKKS_REF_SEED making both
sessions draw identical keys and encryption randomness, and prints which results have identical ciphertext limbs."""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.join(HERE, "..", "oracle")]
os.environ["B200CKKS_REF_SEED"] = "20261019"
os.environ["B200CKKS_ENCRYPT_CONSTANTS"] = "1"
os.environ["B200CKKS_NO_HOIST"] = "1"

import app_cases as c   # noqa: E402
import appref           # noqa: E402
import cnnref           # noqa: E402


def program(a, heap):
    out = {}
    s = a.session(c.SMALL_LOG_N, c.CNN_SMALL_BITS, hamming_weight=64, rotation_steps=list(range(1, 2048)))
    for name, (k, h, w, ci, co, st) in {"conv_s1": (1, 8, 8, 4, 4, 1), "conv_s2": (1, 8, 8, 4, 8, 2), "conv_k2": (2, 4, 4, 8, 8, 1),
                                        "conv_3ch": (1, 8, 8, 3, 4, 1)}.items():
        rng = np.random.default_rng(1)
        x, ct, parms = c.make_tensor(s, rng, k, h, w, ci, limbs=4)
        wt = rng.normal(0, 0.3, 9 * ci * co)
        var, bw = rng.uniform(0.5, 1.5, co), rng.uniform(0.5, 1.0, co)
        o, op = s.conv(ct, parms, co, st, wt, var, bw)
        out[name] = (o.download(), o.info(), op)
    # With the seeded generator every fresh encryption carries the same randomness, so operands that meet a fresh
    # ciphertext inside a layer (batch-norm shift, residual add) are de-correlated by a rotation first - as in the
    # network, where they are convolution outputs.
    rng = np.random.default_rng(2)
    x, ct, parms = c.make_tensor(s, rng, 1, 8, 8, 8, limbs=3)
    s.rotate(ct, 1)
    bias, mean = rng.normal(0, 0.3, 8), rng.normal(0, 0.3, 8)
    var, bw = rng.uniform(0.5, 1.5, 8), rng.uniform(0.5, 1.0, 8)
    o = s.bn(ct, parms, bias, mean, var, bw)
    out["bn"] = (o.download(), o.info())
    x2, ct2, _ = c.make_tensor(s, rng, 1, 8, 8, 8, limbs=2)
    s.rotate(ct2, 3)
    o = s.tensor_add(ct, ct2)                       # mismatched levels: the walk-down branch of add_inplace_reduced_error
    out["add_mismatched_levels"] = (o.download(), o.info())
    o, op = s.downsample(ct, parms)
    out["downsample"] = (o.download(), o.info(), op)
    xa, cta, pa = c.make_tensor(s, rng, 2, 4, 4, 16, limbs=3)
    o, op = s.avgpool(cta, pa, B=40.0)
    out["avgpool"] = (o.download(), o.info(), op)
    o = s.fc(o, op, rng.normal(0, 0.3, (10, 16)), rng.normal(0, 0.3, 10), 10, 16)
    out["fc"] = (o.download(), o.info())
    s.close()

    s = a.session(c.SMALL_LOG_N, c.RELU_BITS, hamming_weight=64)
    ct = s.encrypt(np.random.default_rng(3).uniform(-1, 1, s.slots), 2.0 ** 46, limbs=17)
    o = s.relu(ct)
    out["relu"] = (o.download(), o.info())
    s.close()

    s = a.session(c.SMALL_LOG_N, c.BOOT_BITS, hamming_weight=64)
    b = s.bootstrapper(9)
    if heap is None:      # the reference's own Remez (common/Remez.cpp on the NTL stand-in) made this polynomial
        buf, n = np.zeros(8192), C.c_int()
        a.ck(a.L.bkr_evalmod_heap(b.h, buf.ctypes.data_as(C.POINTER(C.c_double)), len(buf), C.byref(n)))
        heap = buf[:n.value].copy()
    else:                 # the restated layers evaluate the same doubles
        a.ck(a.L.bka_bootstrapper_set_evalmod_heap(b.h, heap.ctypes.data_as(C.POINTER(C.c_double)), len(heap)))
    out["boot_rotation_steps"] = (np.array(b.rotation_steps(), dtype=np.uint64), None)
    for i in range(6):
        out[f"boot_lt_coefficients_{i}"] = (b.lt_coefficients(i).view(np.uint64), None)
    xs = np.tile(np.random.default_rng(4).uniform(-1, 1, 512), s.slots // 512)
    o = b.bootstrap(s.encrypt(xs, 2.0 ** 46, limbs=1), real_message=True)
    out["bootstrap_sparse_real"] = (o.download(), o.info())
    out["bootstrap_error"] = float(np.abs(s.decrypt(o) - xs).max())
    s.close()
    return out, heap


def main():
    with cnnref.workdir():
        ref, heap = program(cnnref.app(), None)
        res, _ = program(appref.app(), heap)
    report = {}
    for k in ref:
        if k == "bootstrap_error":
            report[k] = [ref[k], res[k]]
            continue
        same = ref[k][0].shape == res[k][0].shape and bool(np.array_equal(ref[k][0], res[k][0]))
        report[k] = {"limbs_equal": same, "info": [list(ref[k][1]) if ref[k][1] else None, list(res[k][1]) if res[k][1] else None],
                     "parms": [ref[k][2], res[k][2]] if len(ref[k]) > 2 else None}
    print("REPORT " + json.dumps(report))


if __name__ == "__main__":
    main()
