"""One pass over every kernel family of the engine at the CNN parameter set (N = 2^16, 31 limbs), bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off --set full` captures exactly one launch sequence of each:
NTT / INTT (all load-store functors), key-switch inner product, element-wise, tensor, rescale, ModRaise, encoder FFT,
decode, samplers / encryption.  Usage: python tools/kernel_tour.py  (plain run prints per-op CUDA-event times)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
import numpy as np

import b200ckks as bk

bits = [51] + [46] * 16 + [51] * 14 + [51]
primes = bk.coeff_modulus_create(16, bits)
eng = bk.Context(16, primes)
sk = eng.generate_secret_key(192, 1)
pk = eng.create_public_key(sk)
rk = eng.create_relin_key(sk)
gk = eng.create_galois_keys(sk, [1, 0])
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, eng.slots) + 1j * rng.uniform(-1, 1, eng.slots)
cudart = ctypes.CDLL("libcudart.so")


def tour(timed):
    out = {}

    def t(name, fn):
        eng.sync()
        eng.timer_begin()
        r = fn()
        out[name] = eng.timer_end()
        return r

    pt = t("encode (31 limbs)", lambda: eng.encode(x, 31, 2.0 ** 46))
    ct = t("encrypt (public key)", lambda: eng.encrypt(pk, pt))
    a = ct.copy()
    t("rotate_vector l=31", lambda: eng.rotate_vector_inplace(a, 1, gk))
    t("complex_conjugate l=31", lambda: eng.complex_conjugate_inplace(a, gk))
    hs = t("apply_galois_hoisted x2 l=31", lambda: eng.apply_galois_hoisted(ct, [bk.galois_elt_from_step(16, 1), 2 * eng.n - 1], gk))
    b = ct.copy()
    t("multiply (tensor) l=31", lambda: eng.multiply_inplace(b, a))
    t("relinearize l=31", lambda: eng.relinearize_inplace(b, rk))
    t("rescale_to_next l=31", lambda: eng.rescale_to_next_inplace(b))
    t("square l=30", lambda: eng.square_inplace(b))
    eng.relinearize_inplace(b, rk)
    c = ct.copy()
    t("multiply_plain l=31", lambda: eng.multiply_plain_inplace(c, pt))
    t("add l=31", lambda: eng.add_inplace(c, c))
    t("multiply_const l=31", lambda: eng.multiply_const_inplace(a, 0.37))
    t("add_const l=31", lambda: eng.add_const_inplace(a, -1.0))
    c = ct.copy()
    t("mod_switch_to 31->1", lambda: eng.mod_switch_to_inplace(c, 1))
    t("modraise 1->31", lambda: eng.modraise_inplace(c))
    t("transform_from_ntt l=31", lambda: eng.transform_from_ntt_inplace(a))
    t("transform_to_ntt l=31", lambda: eng.transform_to_ntt_inplace(a))
    d = t("decrypt l=31", lambda: eng.decrypt(sk, ct))
    t("decode l=31", lambda: eng.decode(d))
    if timed:
        for k, v in out.items():
            print(f"{k:32s} {v * 1e3:9.1f} us")
    return hs


tour(False)
eng.sync()
cudart.cudaProfilerStart()
tour(True)
eng.sync()
cudart.cudaProfilerStop()
eng.close()
