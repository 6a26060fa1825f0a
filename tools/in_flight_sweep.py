"""Throughput of the bootstrapped ResNet-20 against the number of images in flight on ONE GPU (one host thread and CUDA
stream per image, shared keys and plaintext cache): `python tools/in_flight_sweep.py [K ...]`.  Device-timed with CUDA
events on the calling thread's stream, which the batch call orders around the workers' streams."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
from b200ckks import synthetic
from b200ckks.app import App

CNN_BITS = [51] + [46] * 16 + [51] * 14 + [51]


def main():
    ks = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
    os.environ.setdefault("B200CKKS_SEED", "0x5EA1C0DE")
    app = App()
    sess = app.session(16, CNN_BITS, hamming_weight=192)
    eng = sess.engine()
    net = sess.resnet(20, synthetic.random_weights(20, seed=0))
    ref = [net.infer(synthetic.synthetic_image(i), trace=False)[0] for i in range(2)]     # warm-up: keys, plans, cache
    out = {}
    for k in ks:
        n = 2 * k if k > 1 else 3
        enc = [net.encrypt_image(synthetic.synthetic_image(i % 2)) for i in range(n)]
        net.infer_encrypted_batch(enc[:k], k)            # worker threads, their streams and staging rings
        sess.sync()
        eng.timer_begin()
        t0 = time.perf_counter()
        res = net.infer_encrypted_batch(enc, k)
        ms = eng.timer_end()
        wall = time.perf_counter() - t0
        logits = [net.decrypt_logits(r) for r in res]
        err = max(float(np.abs(l - ref[i % 2]).max()) for i, l in enumerate(logits))
        out[k] = {"images": n, "device_ms": round(ms, 1), "wall_s": round(wall, 3), "images_per_s": round(n / (ms * 1e-3), 4),
                  "s_per_image_amortised": round(ms * 1e-3 / n, 4), "max_logit_diff_vs_sequential": err}
        print(k, json.dumps(out[k]), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
