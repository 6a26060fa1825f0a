"""Bootstrapping error (max / rms over the slots, three sparse-slot sizes) under the three key-switching settings:
reference-exact, hybrid, hybrid with wide digits at the two levels below the top.  python tools/boot_precision.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
    import numpy as np
    from b200ckks.app import App

    bits = [51] + [46] * 16 + [51] * 14 + [51]
    s = App().session(16, bits, hamming_weight=192)
    for logn in (14, 13, 12):
        boot = s.bootstrapper(logn)
        rng = np.random.default_rng(logn)
        errs = []
        for rep in range(3):
            x = np.tile(rng.uniform(-1, 1, 1 << logn), s.slots >> logn)
            out = boot.bootstrap(s.encrypt(x, 2.0 ** 46, limbs=1), real_message=True)
            e = s.decrypt(out).real - x
            errs.append((np.abs(e).max(), np.sqrt((e ** 2).mean()), e.mean()))
        print(f"  logn={logn}: " + "  ".join(f"max {a:.2e} rms {b:.2e} mean {c:+.1e}" for a, b, c in errs), flush=True)
    s.close()
else:
    for name, env in (("reference-exact", {}), ("hybrid", {"B200CKKS_HYBRID_KS": "1"}),
                      ("hybrid + wide top", {"B200CKKS_HYBRID_KS": "1", "B200CKKS_HYBRID_WIDE_TOP": "1"})):
        print(name, flush=True)
        subprocess.run([sys.executable, __file__, "--one"], env=dict(os.environ, B200CKKS_SEED="0x5EA1C0DE", **env), check=True)
