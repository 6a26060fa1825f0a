"""Bootstrapping time and error at N = 2^16 on the CNN chain in hybrid mode: hoisted baby steps with one ModDown per
rotation against double-hoisted inner sums (one ModDown per giant step).  usage: python tools/boot_time.py [logn ...]"""
import os
import sys

os.environ["B200CKKS_HYBRID_KS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fhe-gpt-2_b200", "python")]
import numpy as np
from b200ckks.app import App

BITS = [51] + [46] * 16 + [51] * 14 + [51]
if os.environ.get("BOOT_CHAIN") == "gpt2":  # the GPT-2 chain (gpt2/util.h:37-75): 60-bit special prime, mixed butterflies
    BITS = [49] + [46] * 21 + [49] * 14 + [60]
logns = [int(a) for a in sys.argv[1:]] or [14, 13, 12]
s = App().session(16, BITS, hamming_weight=192)
eng = s.engine()
for logn in logns:
    n = 1 << logn
    xs = np.tile(np.random.default_rng(4).uniform(-1, 1, n), s.slots // n)
    boot = s.bootstrapper(logn)
    for name, (on, double) in ({"double hoisted": (True, True)} if os.environ.get("B200CKKS_NO_DOUBLE_HOIST") is None else {"hoisted": (True, False)}).items():
        boot.set_hoisting(on, double=double)
        for _ in range(3):
            out = boot.bootstrap(s.encrypt(xs, 2.0 ** 46, limbs=1), real_message=True)
        ms = 1e30
        for rep in range(3):
            cts = [s.encrypt(xs, 2.0 ** 46, limbs=1) for _ in range(5)]
            s.sync()
            c0 = eng.kernel_counters()
            eng.timer_begin()
            for ct in cts:
                out = boot.bootstrap(ct, real_message=True)
            ms = min(ms, eng.timer_end() / len(cts))
            c1 = eng.kernel_counters()
        fwd = (c1["fwd_cols"][1] - c0["fwd_cols"][1]) // len(cts)
        inv = (c1["inv_cols"][1] - c0["inv_cols"][1]) // len(cts)
        eng.profile_begin_all()
        boot.bootstrap(s.encrypt(xs, 2.0 ** 46, limbs=1), real_message=True)
        prof = eng.profile_end_all()
        print("      per family (launches, ms with events around every launch): " +
              ", ".join(f"{k} {v[0]}/{v[1]:.1f}" for k, v in prof.items() if v[0]), flush=True)
        err = np.abs(s.decrypt(out) - xs)
        print(f"logn {logn:2d} {name:15s} {ms:7.2f} ms per bootstrap, forward NTTs {fwd}, inverse NTTs {inv}, "
              f"error max {err.max():.2e} rms {np.sqrt((err ** 2).mean()):.2e}", flush=True)
print(f"plaintext cache {s.plain_cache()}")
s.close()
