"""Merged ModDown + rescale (bk_relinearize_rescale_inplace) against relinearize_inplace + rescale_to_next_inplace on the
same product, level by level: decrypted difference between the two, and each against x * x.

  python tools/merged_rescale_check.py [log_n]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
import numpy as np

import b200ckks as bk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bits = [51] + [46] * 16 + [51] * 14 + [51]
primes = bk.coeff_modulus_create(log_n, bits)
eng = bk.Context(log_n, primes)
eng.set_hybrid(True)
sk = eng.generate_secret_key(192, 11)
pk = eng.create_public_key(sk)
rk = eng.create_relin_key(sk)
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, eng.slots) + 1j * rng.uniform(-1, 1, eng.slots)
ct = eng.encrypt(pk, eng.encode(x, len(bits) - 1, 2.0 ** 46))
dec = lambda c: eng.decode(eng.decrypt(sk, c))
print("level  two calls vs x*x   merged vs x*x   merged vs two calls")
for l in range(31, 2, -1):
    c = ct.copy()
    eng.mod_switch_to_inplace(c, l)
    a = c.copy()
    eng.multiply_inplace(a, c)
    b = a.copy()
    eng.relinearize_inplace(a, rk)
    eng.rescale_to_next_inplace(a)
    eng.relinearize_rescale_inplace(b, rk)
    da, db = dec(a), dec(b)
    print(f"{l:5d}  {np.abs(da - x * x).max():.3e}  {np.abs(db - x * x).max():.3e}  {np.abs(da - db).max():.3e}"
          f"  mean of the difference {np.mean(db - da):.3e}")

print("level  rotation by 5, hybrid, against the plain roll")
gk = eng.create_galois_keys(sk, [5])
for l in (31, 30, 29, 28, 24, 20, 17, 12, 8, 6, 5, 3):
    c = ct.copy()
    eng.mod_switch_to_inplace(c, l)
    eng.rotate_vector_inplace(c, 5, gk)
    print(f"{l:5d}  {np.abs(dec(c) - np.roll(x, -5)).max():.3e}")
