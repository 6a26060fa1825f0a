"""ctypes wrapper around oracle/_ref/libseal_ref.so (the reference's own modified SEAL 3.6.6,
compiled in place by oracle/Makefile, driven through oracle/ref_shim.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product path never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libseal_ref.so")

OPS = dict(
    add=1, sub=2, multiply=3, square=4, relinearize=5, rescale=6, mod_switch_next=7, rotate=8,
    conjugate=9, negate=10, add_plain=11, sub_plain=12, multiply_plain=13, add_const=14,
    multiply_const=15, add_reduced_error=16, sub_reduced_error=17, multiply_reduced_error=18,
    ntt_fwd=19, ntt_inv=20, mod_switch_to=21, multiply_vector=22,
)

CNN_BITS = [51] + [46] * 16 + [51] * 14 + [51]  # infer_seal.cpp:288-322
GPT2_BITS = [49] + [46] * 21 + [49] * 14 + [60]  # gpt2 util.h:22-27


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefSeal:
    """One SEALContext + keys of the reference, raw-limb access."""

    def __init__(self, log_n, bits, hamming_weight=192, sparse_slots=0, seed=1):
        self.L = lib()
        arr = (C.c_int * len(bits))(*bits)
        self.h = C.c_void_p(self.L.ref_create(log_n, arr, len(bits), hamming_weight, sparse_slots, C.c_uint64(seed)))
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        self.log_n = log_n
        self.n = 1 << log_n
        self.n_primes = self.L.ref_n_primes(self.h)
        p = np.zeros(self.n_primes, dtype=np.uint64)
        self.L.ref_get_primes(self.h, _p(p))
        self.primes = p

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    # --- kernels -------------------------------------------------------------------------
    def ntt(self, prime_idx, data, inverse=False):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        self._ck(self.L.ref_ntt(self.h, prime_idx, _p(d), int(inverse)))
        return d

    def root_powers(self, prime_idx, inverse=False):
        out = np.zeros(self.n, dtype=np.uint64)
        self._ck(self.L.ref_root_powers(self.h, prime_idx, _p(out), int(inverse)))
        return out

    def galois_elt(self, step):
        e = C.c_uint32(0)
        self._ck(self.L.ref_galois_elt_from_step(self.h, step, C.byref(e)))
        return e.value

    def apply_galois_ntt(self, elt, limb):
        d = np.ascontiguousarray(limb, dtype=np.uint64)
        out = np.zeros_like(d)
        self._ck(self.L.ref_apply_galois_ntt(self.h, C.c_uint32(elt), _p(d), _p(out)))
        return out

    # --- keys ------------------------------------------------------------------------------
    def secret_key(self):
        out = np.zeros((self.n_primes, self.n), dtype=np.uint64)
        self._ck(self.L.ref_get_secret_key(self.h, _p(out)))
        return out

    def public_key(self):
        out = np.zeros((2, self.n_primes, self.n), dtype=np.uint64)
        self._ck(self.L.ref_get_public_key(self.h, _p(out)))
        return out

    def relin_key(self, dump=True):
        if not dump:
            self._ck(self.L.ref_make_relin_key(self.h, None))
            return None
        out = np.zeros((self.n_primes - 1, 2, self.n_primes, self.n), dtype=np.uint64)
        self._ck(self.L.ref_make_relin_key(self.h, _p(out)))
        return out

    def make_galois_keys(self, steps):
        arr = (C.c_int * len(steps))(*steps)
        self._ck(self.L.ref_make_galois_keys(self.h, arr, len(steps)))

    def galois_key(self, elt):
        out = np.zeros((self.n_primes - 1, 2, self.n_primes, self.n), dtype=np.uint64)
        self._ck(self.L.ref_get_galois_key(self.h, C.c_uint32(elt), _p(out)))
        return out

    # --- ct / pt registry ------------------------------------------------------------------
    def ct_new(self):
        return self.L.ref_ct_new(self.h)

    def ct_free(self, i):
        self.L.ref_ct_free(self.h, i)

    def ct_info(self, i):
        size, limbs, ntt = C.c_int(), C.c_int(), C.c_int()
        scale = C.c_double()
        self._ck(self.L.ref_ct_info(self.h, i, C.byref(size), C.byref(limbs), C.byref(scale), C.byref(ntt)))
        return size.value, limbs.value, scale.value, bool(ntt.value)

    def ct_get(self, i):
        size, limbs, _, _ = self.ct_info(i)
        out = np.zeros((size, limbs, self.n), dtype=np.uint64)
        self._ck(self.L.ref_ct_get(self.h, i, _p(out)))
        return out

    def ct_set(self, i, data, scale, is_ntt=True):
        d = np.ascontiguousarray(data, dtype=np.uint64)
        size, limbs, n = d.shape
        assert n == self.n
        self._ck(self.L.ref_ct_set(self.h, i, _p(d), size, limbs, C.c_double(scale), int(is_ntt)))

    def ct_set_scale(self, i, scale):
        self._ck(self.L.ref_ct_set_scale(self.h, i, C.c_double(scale)))

    def ct_copy(self, dst, src):
        self._ck(self.L.ref_ct_copy(self.h, dst, src))

    def pt_new(self):
        return self.L.ref_pt_new(self.h)

    def pt_free(self, i):
        self.L.ref_pt_free(self.h, i)

    def pt_info(self, i):
        limbs = C.c_int()
        scale = C.c_double()
        self._ck(self.L.ref_pt_info(self.h, i, C.byref(limbs), C.byref(scale)))
        return limbs.value, scale.value

    def pt_get(self, i):
        limbs, _ = self.pt_info(i)
        out = np.zeros((limbs, self.n), dtype=np.uint64)
        self._ck(self.L.ref_pt_get(self.h, i, _p(out)))
        return out

    # --- encoder / encryptor -----------------------------------------------------------------
    def encode(self, pt, values, limbs, scale):
        v = np.asarray(values)
        if np.iscomplexobj(v):
            v = np.ascontiguousarray(v, dtype=np.complex128)
            self._ck(self.L.ref_encode(self.h, pt, _p(v.view(np.float64)), len(v), 1, limbs, C.c_double(scale)))
        else:
            v = np.ascontiguousarray(v, dtype=np.float64)
            self._ck(self.L.ref_encode(self.h, pt, _p(v), len(v), 0, limbs, C.c_double(scale)))

    def encode_scalar(self, pt, value, limbs, scale):
        self._ck(self.L.ref_encode_scalar(self.h, pt, C.c_double(value), limbs, C.c_double(scale)))

    def decode(self, pt):
        out = np.zeros(self.n // 2, dtype=np.complex128)
        self._ck(self.L.ref_decode(self.h, pt, _p(out.view(np.float64))))
        return out

    def encrypt(self, pt, ct):
        self._ck(self.L.ref_encrypt(self.h, pt, ct))

    def decrypt(self, ct, pt):
        self._ck(self.L.ref_decrypt(self.h, ct, pt))

    # --- evaluator -----------------------------------------------------------------------------
    def op(self, name, a, b=0, iarg=0, darg=0.0):
        self._ck(self.L.ref_op(self.h, OPS[name], a, b, iarg, C.c_double(darg)))

    def multiply_vector_reduced_error(self, a, values):
        v = np.asarray(values)
        if np.iscomplexobj(v):
            v = np.ascontiguousarray(v, dtype=np.complex128)
            self._ck(self.L.ref_multiply_vector_reduced_error(self.h, a, _p(v.view(np.float64)), len(v), 1))
        else:
            v = np.ascontiguousarray(v, dtype=np.float64)
            self._ck(self.L.ref_multiply_vector_reduced_error(self.h, a, _p(v), len(v), 0))

    def time_op(self, name, a, b=0, iarg=0, darg=0.0, threads=1, reps=1):
        wall, mean = C.c_double(), C.c_double()
        self._ck(self.L.ref_time_op(self.h, OPS[name], a, b, iarg, C.c_double(darg), threads, reps,
                                    C.byref(wall), C.byref(mean)))
        return wall.value, mean.value

    WHAT = dict(ciphertext=0, plaintext=1, relin_keys=2, galois_keys=3, secret_key=4, public_key=5)

    def save(self, what, path, ident=0):
        """the reference's own save() (compr_mode_type::none) of an object into a file"""
        self._ck(self.L.ref_save(self.h, self.WHAT[what], ident, path.encode()))

    def load(self, what, path, ident=0):
        self._ck(self.L.ref_load(self.h, self.WHAT[what], ident, path.encode()))

    def max_threads(self):
        return self.L.ref_max_threads()
