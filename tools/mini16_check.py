import sys, os
sys.path[:0] = ['/root/repo/tests', '/root/repo/oracle', '/root/repo/fhe-gpt-2_b200/python']
import numpy as np
from b200ckks.app import App
import app_cases as c
a = App()
s = a.session(16, c.BOOT_BITS, hamming_weight=192, rotation_steps=[1, 5])
x = np.random.default_rng(0).uniform(-1, 1, s.slots)
for limbs in (31, 17, 3, 2):
    ct = s.encrypt(x, 2.0**46, limbs=limbs)
    s.rotate(ct, 1)
    print(limbs, 'rotate err', np.abs(s.decrypt(ct).real - np.roll(x, -1)).max(), flush=True)
    ct2 = s.encrypt(x, 2.0**46, limbs=limbs)
    if limbs > 2:
        s.multiply_relin_rescale(ct2, s.encrypt(x, 2.0**46, limbs=limbs))
        print(limbs, 'mul err', np.abs(s.decrypt(ct2).real - x*x).max(), flush=True)
s.close()
print('done')
