"""Latency of one rotation / one multiply+relinearize+rescale at low levels (where a ResNet's convolutions and ReLUs
live): device time per call in a dependent chain, and the host time it takes just to enqueue the call (GPU idle
waiting for the host when that is the larger one).  usage: python tools/small_ops.py"""
import os
import sys
import time

os.environ.setdefault("B200CKKS_HYBRID_KS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fhe-gpt-2_b200", "python")]
import numpy as np
from b200ckks.app import App

BITS = [51] + [46] * 16 + [51] * 14 + [51]
s = App().session(16, BITS, hamming_weight=192, rotation_steps=[1, 2, 4, 8])
eng = s.engine()
x = np.random.default_rng(0).uniform(-1, 1, s.slots)
print(f"PDL {'off' if os.environ.get('B200CKKS_NO_PDL') else 'on'}")
print(f"{'limbs':>5} | {'rotate us':>10} {'enqueue us':>10} {'launches':>8} | {'mul+relin+rescale us':>20} {'enqueue us':>10}")
for limbs in (2, 3, 5, 9, 13, 17):
    ct = s.encrypt(x, 2.0 ** 46)
    s.mod_switch_to(ct, limbs)
    for _ in range(20):
        s.rotate(ct, 1)
    s.sync()
    reps = 200
    l0 = eng.launch_count()
    eng.timer_begin()
    t0 = time.perf_counter()
    for _ in range(reps):
        s.rotate(ct, 1)
    host = (time.perf_counter() - t0) / reps * 1e6
    dev = eng.timer_end() / reps * 1e3
    launches = (eng.launch_count() - l0) / reps
    a = s.encrypt(x, 2.0 ** 46)
    s.mod_switch_to(a, limbs)
    mdev = mhost = float("nan")
    if limbs > 2:
        for _ in range(5):
            s.multiply_relin_rescale(a.clone(), a)
        mreps = 50
        cs = [a.clone() for _ in range(mreps)]
        s.sync()
        eng.timer_begin()
        t0 = time.perf_counter()
        for c in cs:
            s.multiply_relin_rescale(c, a)
        mhost = (time.perf_counter() - t0) / mreps * 1e6
        mdev = eng.timer_end() / mreps * 1e3
    print(f"{limbs:5d} | {dev:10.1f} {host:10.1f} {launches:8.1f} | {mdev:20.1f} {mhost:10.1f}", flush=True)
s.close()
