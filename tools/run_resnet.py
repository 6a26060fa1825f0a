"""Bootstrapped ResNet on one synthetic CIFAR-10 image through the engine, checked against the float64 plaintext model.

  python tools/run_resnet.py [--layers 20] [--images 1] [--out gpurun_out/resnet20.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import plain_model as pm
from b200ckks.app import App

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=20)
ap.add_argument("--images", type=int, default=1)
ap.add_argument("--out", default=None)
args = ap.parse_args()

bits = [51] + [46] * 16 + [51] * 14 + [51]
app = App()
t0 = time.time()
sess = app.session(16, bits, hamming_weight=192)
w = pm.random_weights(args.layers, seed=0)
net = sess.resnet(args.layers, w)
print(f"setup {time.time() - t0:.1f}s", flush=True)
report = {"layers": args.layers, "images": []}
for i in range(args.images):
    img = pm.synthetic_image(i)
    sess.stats(reset=True)
    t0 = time.time()
    logits, trace = net.infer(img)
    dt = time.time() - t0
    ref = []
    want = pm.resnet_forward(args.layers, w, img, collect=ref)
    err = float(np.abs(logits - want).max())
    by_op = {}
    for r in trace:
        by_op.setdefault(r["op"], [0, 0.0])
        by_op[r["op"]][0] += 1
        by_op[r["op"]][1] += r["ms"]
    kb, kg = sess.key_residency()
    print(f"image {i}: {dt:.2f}s  logits max err {err:.3e}  argmax {int(np.argmax(logits))} vs {int(np.argmax(want))}", flush=True)
    print("  logits", np.round(logits, 4).tolist())
    print("  model ", np.round(want, 4).tolist())
    print("  per op:", {k: (v[0], round(v[1], 1)) for k, v in by_op.items()})
    print(f"  galois keys resident {kb / 2**30:.1f} GiB after {kg} generations; stats {sess.stats()}", flush=True)
    report["images"].append({"seconds": dt, "logit_err": err, "logits": logits.tolist(), "model": want.tolist(),
                             "per_op_ms": {k: v for k, v in by_op.items()}, "trace": trace, "stats": sess.stats(),
                             "key_gib": kb / 2**30, "key_generations": kg})
if args.out:
    json.dump(report, open(args.out, "w"))
