"""Decryption error of rotate / relinearize per level: SEAL-layout path vs level-aware hybrid key switching, and the time
of one rotation in each mode.  python tools/hybrid_noise.py [log_n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
import b200ckks as bk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
# --ncu-level L: bracket ONE hybrid rotation at level L with cudaProfilerStart/Stop (ncu --profile-from-start off)
ncu_level = int(sys.argv[sys.argv.index("--ncu-level") + 1]) if "--ncu-level" in sys.argv else 0
bits = [51] + [46] * 16 + [51] * 14 + [51]
levels = [31, 30, 29, 28, 24, 20, 17, 12, 8, 5, 3, 2, 1]
primes = bk.coeff_modulus_create(log_n, bits)
for hybrid in (False, True):
    eng = bk.Context(log_n, primes)
    eng.set_hybrid(hybrid)
    sk = eng.generate_secret_key(192, 11)
    pk = eng.create_public_key(sk)
    rk = eng.create_relin_key(sk)
    gk = eng.create_galois_keys(sk, [5])
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, eng.slots) + 1j * rng.uniform(-1, 1, eng.slots)
    ct = eng.encrypt(pk, eng.encode(x, 31, 2.0 ** 46))
    dec = lambda c: eng.decode(eng.decrypt(sk, c))
    print("hybrid" if hybrid else "seal-layout", "fresh error", np.abs(dec(ct) - x).max())
    for l in levels:
        c = ct.copy()
        eng.mod_switch_to_inplace(c, l)
        r = c.copy()
        eng.rotate_vector_inplace(r, 5, gk)      # generates the key of this level on first use
        e_rot = np.abs(dec(r) - np.roll(x, -5)).max()
        reps = 20
        eng.sync()
        eng.timer_begin()
        for _ in range(reps):
            eng.rotate_vector_inplace(r, 5, gk)
        us = eng.timer_end() * 1e3 / reps
        e_mul = float("nan")
        if l >= 3:
            m = c.copy()
            eng.multiply_inplace(m, c)
            eng.relinearize_inplace(m, rk)
            eng.rescale_to_next_inplace(m)
            e_mul = np.abs(dec(m) - x * x).max()
        print(f"  l={l:2d} rotate err {e_rot:.2e}  mul+relin err {e_mul:.2e}  rotate {us:8.1f} us")
    print("  ", eng.hybrid_info())
    if hybrid and ncu_level:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        c = ct.copy()
        eng.mod_switch_to_inplace(c, ncu_level)
        eng.rotate_vector_inplace(c, 5, gk)
        eng.sync()
        cudart.cudaProfilerStart()
        eng.rotate_vector_inplace(c, 5, gk)
        eng.sync()
        cudart.cudaProfilerStop()
    eng.close()
