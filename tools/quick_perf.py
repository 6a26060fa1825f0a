"""Quick wall-clock probe of the hot ops (host timer around a synced batch). Not the bench."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
import numpy as np
import b200ckks as bk

bits = [51] + [46] * 16 + [51] * 14 + [51]
primes = bk.coeff_modulus_create(16, bits)
t0 = time.time()
eng = bk.Context(16, primes)
print("context", time.time() - t0)
t0 = time.time()
sk = eng.generate_secret_key(192, 1)
pk = eng.create_public_key(sk)
rk = eng.create_relin_key(sk)
gk = eng.create_galois_keys(sk, [1])
eng.sync()
print("keygen", time.time() - t0)
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, 32768)
pt = eng.encode(x, 31, 2.0 ** 46)
ct0 = eng.encrypt(pk, pt)
err = np.max(np.abs(eng.decode(eng.decrypt(sk, ct0)) - x))
print("roundtrip err", err)
for chunk in (2, 4, 8):
    eng.set_ks_chunk(chunk)
    for limbs in (31, 17, 3):
        ct = ct0.copy()
        eng.mod_switch_to_inplace(ct, limbs)
        for name, fn in (("rotate", lambda c: eng.rotate_vector_inplace(c, 1, gk)),
                         ("rescale", None), ("mulplain", None), ("ntt", None)):
            if name != "rotate" and chunk != 4:
                continue
            reps = 20
            cts = [ct.copy() for _ in range(reps)]
            ptl = eng.encode(x, limbs, 2.0 ** 46)
            eng.sync()
            t = time.perf_counter()
            for c in cts:
                if name == "rotate":
                    fn(c)
                elif name == "rescale":
                    if limbs > 1:
                        eng.rescale_to_next_inplace(c)
                elif name == "mulplain":
                    eng.multiply_plain_inplace(c, ptl)
                else:
                    eng.transform_from_ntt_inplace(c)
                    eng.transform_to_ntt_inplace(c)
            eng.sync()
            dt = (time.perf_counter() - t) / reps
            print(f"chunk={chunk} limbs={limbs:2d} {name:8s} {dt*1e6:10.1f} us")
y = eng.decode(eng.decrypt(sk, cts[0]))
print("launches", eng.launch_count())
