"""ncu target: one double-hoisted sparse-slot bootstrap (logn = 14) at N = 2^16 in hybrid mode between
cudaProfilerStart / cudaProfilerStop (`ncu --profile-from-start off ...`).  Keys, plaintext cache and scratch are warm.
  python tools/ncu_bootstrap.py [logn]"""
import ctypes
import os
import sys

os.environ["B200CKKS_HYBRID_KS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fhe-gpt-2_b200", "python")]
import numpy as np
from b200ckks.app import App

cudart = ctypes.CDLL("libcudart.so")
BITS = [51] + [46] * 16 + [51] * 14 + [51]
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 14
s = App().session(16, BITS, hamming_weight=192)
n = 1 << logn
xs = np.tile(np.random.default_rng(4).uniform(-1, 1, n), s.slots // n)
boot = s.bootstrapper(logn)
for _ in range(2):
    out = boot.bootstrap(s.encrypt(xs, 2.0 ** 46, limbs=1), real_message=True)
ct = s.encrypt(xs, 2.0 ** 46, limbs=1)
s.sync()
cudart.cudaProfilerStart()
out = boot.bootstrap(ct, real_message=True)
s.sync()
cudart.cudaProfilerStop()
print("bootstrap error", float(np.abs(s.decrypt(out) - xs).max()))
s.close()
