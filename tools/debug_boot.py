import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import app_cases as cases
from b200ckks.app import App
app = App()
for log_n in (13, 14, 15, 16):
    s = app.session(log_n, cases.BOOT_BITS, hamming_weight=64 if log_n < 16 else 192)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, s.slots) + 1j * rng.uniform(-1, 1, s.slots)
    v = rng.uniform(-1, 1, s.slots) + 1j * rng.uniform(-1, 1, s.slots)
    ct = s.encrypt(x, 2.0 ** 46)
    print(log_n, "enc/dec", np.abs(s.decrypt(ct) - x).max(), flush=True)
    c2 = ct.clone(); s.multiply_vector_rescale(c2, v)
    print(log_n, "mulvec", np.abs(s.decrypt(c2) - x * v).max(), c2.info(), flush=True)
    s.add_rotation_steps([1, s.slots // 2, 3])
    for st in (1, s.slots // 2, 3):
        c3 = ct.clone(); s.rotate(c3, st)
        print(log_n, "rot", st, np.abs(s.decrypt(c3) - np.roll(x, -st)).max(), flush=True)
    c4 = s.encrypt(x, 2.0 ** 46, limbs=1)
    print(log_n, "enc1/dec", np.abs(s.decrypt(c4) - x).max(), flush=True)
    c5 = s.encrypt(x.real, 2.0 ** 46, limbs=17); r = s.relu(c5)
    print(log_n, "relu", np.abs(s.decrypt(r).real - np.maximum(x.real, 0)).max(), flush=True)
    for logn in (log_n - 3, log_n - 2):
        try:
            e = cases.case_bootstrap(s, logn=logn, real=True, tol=1e-3)
            print(log_n, "boot logn", logn, "err", e, flush=True)
        except AssertionError as ex:
            print(log_n, "boot logn", logn, "FAILED", str(ex)[:80], flush=True)
    s.close()
