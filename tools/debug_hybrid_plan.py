"""debug helper: the key-plan workload in hybrid mode with a device wait after every launch"""
import os, sys
os.environ["B200CKKS_HYBRID_KS"] = "1"
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "fhe-gpt-2_b200", "python")]
import numpy as np
from b200ckks.app import App
BITS = [50] + [40] * 8 + [50]
s = App().session(13, BITS, hamming_weight=64, rotation_steps=[1, 3, 0])
rng = np.random.default_rng(5)
x, y = rng.uniform(-1, 1, s.slots), rng.uniform(-1, 1, s.slots)
a = s.encrypt(x, 2.0 ** 40, limbs=7); print("enc", flush=True)
s.rotate(a, 1); s.sync(); print("rot7", flush=True)
b = s.encrypt(y, 2.0 ** 40, limbs=7)
s.multiply_relin_rescale(a, b); s.sync(); print("mul", flush=True)
s.rotate(a, 3); s.sync(); print("rot6", flush=True)
s.mod_switch_to(a, 3); s.rotate(a, 1); s.sync(); print("rot3", flush=True)
print(s.key_plan())
