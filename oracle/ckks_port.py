"""ctypes wrapper around oracle/_ref/libckks_port.so (oracle/ckks_port.c, our plain-C
restatement of the reference's hot-path algorithms).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libckks_port.so")


class _Tables(C.Structure):
    _fields_ = [("log_n", C.c_int), ("n", C.c_size_t), ("q", C.c_uint64), ("root", C.c_uint64),
                ("root_powers", C.POINTER(C.c_uint64)), ("root_powers_q", C.POINTER(C.c_uint64)),
                ("inv_root_powers", C.POINTER(C.c_uint64)), ("inv_root_powers_q", C.POINTER(C.c_uint64)),
                ("inv_n", C.c_uint64), ("inv_n_q", C.c_uint64)]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        for f in ("port_barrett_reduce_128", "port_barrett_reduce_64", "port_mulmod", "port_shoup_quotient",
                  "port_mulmod_operand", "port_minimal_primitive_root"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_uint64] * {"port_barrett_reduce_128": 3, "port_barrett_reduce_64": 2,
                                                      "port_mulmod": 3, "port_shoup_quotient": 2,
                                                      "port_mulmod_operand": 4, "port_minimal_primitive_root": 2}[f]
        L.port_galois_elt_from_step.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Tables:
    """NTTTables for a list of primes (one per key-level limb)."""

    def __init__(self, log_n, primes):
        self.L = lib()
        self.log_n, self.n = log_n, 1 << log_n
        self.primes = np.ascontiguousarray(primes, dtype=np.uint64)
        self.arr = (_Tables * len(primes))()
        for i, q in enumerate(primes):
            if self.L.port_ntt_tables_init(C.byref(self.arr[i]), log_n, C.c_uint64(int(q))) != 0:
                raise ValueError(f"no primitive root for q={q}")

    def __del__(self):
        for t in self.arr:
            self.L.port_ntt_tables_free(C.byref(t))

    def root_powers(self, i, inverse=False):
        t = self.arr[i]
        src = t.inv_root_powers if inverse else t.root_powers
        return np.ctypeslib.as_array(src, shape=(self.n,)).copy()

    def ntt(self, i, data, inverse=False, lazy=False):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        name = ("port_intt" if inverse else "port_ntt") + ("_lazy" if lazy else "")
        getattr(self.L, name)(_p(d), C.byref(self.arr[i]))
        return d

    def divide_and_round_q_last_ntt(self, poly):
        d = np.ascontiguousarray(poly, dtype=np.uint64).copy()
        limbs = d.shape[0]
        self.L.port_divide_and_round_q_last_ntt(_p(d), limbs, self.arr)
        return d[:limbs - 1]

    def switch_key(self, ct, target, key):
        c = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        t = np.ascontiguousarray(target, dtype=np.uint64)
        k = np.ascontiguousarray(key, dtype=np.uint64)
        l = c.shape[1]
        self.L.port_switch_key(_p(c), _p(t), _p(k), l, len(self.arr), self.arr)
        return c

    def apply_galois_ct(self, ct, elt, key):
        c = np.ascontiguousarray(ct, dtype=np.uint64).copy()
        k = np.ascontiguousarray(key, dtype=np.uint64)
        self.L.port_apply_galois_ct(_p(c), c.shape[1], C.c_uint32(elt), _p(k), len(self.arr), self.log_n, self.arr)
        return c

    def relinearize(self, ct3, key):
        c = np.ascontiguousarray(ct3, dtype=np.uint64)
        out = c[:2].copy()
        return self.switch_key(out, c[2], key)

    def rescale(self, ct):
        c = np.ascontiguousarray(ct, dtype=np.uint64)
        return np.stack([self.divide_and_round_q_last_ntt(c[p]) for p in range(c.shape[0])])

    def multiply(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        l = a.shape[1]
        out = np.zeros((3, l, self.n), dtype=np.uint64)
        self.L.port_ckks_multiply(_p(a), _p(b), _p(out), l, C.c_size_t(self.n), _p(self.primes))
        return out


def galois_elt_from_step(log_n, step):
    return lib().port_galois_elt_from_step(log_n, step)


def galois_table_ntt(log_n, elt):
    out = np.zeros(1 << log_n, dtype=np.uint32)
    lib().port_galois_table_ntt(log_n, C.c_uint32(elt), _p(out))
    return out


def apply_galois_ntt(data, log_n, elt):
    d = np.ascontiguousarray(data, dtype=np.uint64)
    out = np.zeros_like(d)
    lib().port_apply_galois_ntt(_p(d), log_n, C.c_uint32(elt), _p(out))
    return out


def apply_galois(data, log_n, elt, q):
    d = np.ascontiguousarray(data, dtype=np.uint64)
    out = np.zeros_like(d)
    lib().port_apply_galois(_p(d), log_n, C.c_uint32(elt), C.c_uint64(q), _p(out))
    return out


def dyadic_product(a, b, q):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    lib().port_dyadic_product(_p(a), _p(b), C.c_size_t(a.size), C.c_uint64(q), _p(out))
    return out


def modraise_coeffs(src, primes):
    s = np.ascontiguousarray(src, dtype=np.uint64)
    p = np.ascontiguousarray(primes, dtype=np.uint64)
    out = np.zeros((len(p), s.size), dtype=np.uint64)
    lib().port_modraise_coeffs(_p(s), _p(out), len(p), C.c_size_t(s.size), _p(p))
    return out
