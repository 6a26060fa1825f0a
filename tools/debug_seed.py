import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import app_cases as cases
from b200ckks.app import App
app = App()
s = app.session(16, cases.BOOT_BITS, hamming_weight=192)
rng = np.random.default_rng(0)
for logn in (14, 12):
    n = 1 << logn
    boot = s.bootstrapper(logn)
    for mag in (1.0, 0.05):
        x = rng.uniform(-mag, mag, n); xs = np.tile(x, s.slots // n)
        out = boot.bootstrap(s.encrypt(xs, 2.0 ** 46, limbs=1), real_message=True)
        e = np.abs(s.decrypt(out) - xs)
        print(os.environ.get("B200CKKS_SEED"), "logn", logn, "mag", mag, "max", e.max(), "mean", e.mean(), "real-part bias", (s.decrypt(out).real - xs).mean(), flush=True)
# relu precision
x = rng.uniform(-1, 1, s.slots)
r = s.relu(s.encrypt(x, 2.0 ** 46, limbs=17))
import plain_model as pm
print("relu vs model", np.abs(s.decrypt(r).real - pm.minimax_relu(x)).max())
