"""ctypes binding of libb200ckks.so (include/b200ckks.h) - the harness used by tests/ and bench.py.

This is NOT a second implementation: every method is one call through the C ABI into the
sm_100a CUDA library.  If the library is missing, importing this module raises - there is no
CPU path to fall back to.  Method names follow the reference's seal::Evaluator /
CKKSEncoder / KeyGenerator / Encryptor / Decryptor members they stand for.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libb200ckks.so"))
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "..", "include", "b200ckks.h"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU fallback)")

_L = C.CDLL(LIB_PATH)
_L.bk_last_error.restype = C.c_char_p
_L.bk_version.restype = C.c_char_p

BK_OK, BK_INVALID_ARGUMENT, BK_LOGIC_ERROR, BK_OUT_OF_RANGE, BK_CUDA_ERROR, BK_NO_DEVICE = range(6)


class InvalidArgument(ValueError):
    """std::invalid_argument in the reference"""


class LogicError(RuntimeError):
    """std::logic_error in the reference"""


class CudaError(RuntimeError):
    pass


class NoDevice(RuntimeError):
    pass


_EXC = {BK_INVALID_ARGUMENT: InvalidArgument, BK_LOGIC_ERROR: LogicError, BK_OUT_OF_RANGE: IndexError,
        BK_CUDA_ERROR: CudaError, BK_NO_DEVICE: NoDevice}


def _ck(rc):
    if rc != BK_OK:
        raise _EXC.get(rc, RuntimeError)(_L.bk_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    return _L


def version():
    return _L.bk_version().decode()


KERNEL_TAGS = dict(fwd_cols=0, fwd_blocks=1, inv_blocks=2, inv_cols=3, ks_mac=4, elementwise=5)


# ---- host-only helpers -------------------------------------------------------------------------
def coeff_modulus_create(log_n, bits):
    arr = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    _ck(_L.bk_coeff_modulus_create(log_n, arr, len(bits), _ptr(out)))
    return out


def minimal_primitive_root(log_n, q):
    r = C.c_uint64()
    _ck(_L.bk_minimal_primitive_root(log_n, C.c_uint64(int(q)), C.byref(r)))
    return r.value


def galois_elt_from_step(log_n, step):
    e = C.c_uint32()
    _ck(_L.bk_galois_elt_from_step(log_n, step, C.byref(e)))
    return e.value


def galois_table_ntt(log_n, elt):
    out = np.zeros(1 << log_n, dtype=np.uint32)
    _ck(_L.bk_galois_table_ntt(log_n, C.c_uint32(elt), _ptr(out)))
    return out


def ntt_root_powers(log_n, q, inverse=False):
    out = np.zeros(1 << log_n, dtype=np.uint64)
    _ck(_L.bk_ntt_root_powers(log_n, C.c_uint64(int(q)), int(inverse), _ptr(out)))
    return out


# ---- objects ---------------------------------------------------------------------------------------
class Ciphertext:
    def __init__(self, ctx):
        self.ctx = ctx
        self.h = C.c_void_p()
        _ck(_L.bk_ct_create(ctx.h, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            _L.bk_ct_destroy(self.h)
            self.h = None

    def info(self):
        size, limbs, ntt = C.c_int(), C.c_int(), C.c_int()
        scale = C.c_double()
        _ck(_L.bk_ct_info(self.h, C.byref(size), C.byref(limbs), C.byref(scale), C.byref(ntt)))
        return size.value, limbs.value, scale.value, bool(ntt.value)

    size = property(lambda s: s.info()[0])
    limbs = property(lambda s: s.info()[1])
    is_ntt_form = property(lambda s: s.info()[3])

    @property
    def scale(self):
        return self.info()[2]

    @scale.setter
    def scale(self, v):
        _ck(_L.bk_ct_set_scale(self.h, C.c_double(v)))

    def upload(self, data, scale, is_ntt=True):
        d = np.ascontiguousarray(data, dtype=np.uint64)
        size, limbs, n = d.shape
        assert n == self.ctx.n
        _ck(_L.bk_ct_upload(self.h, _ptr(d), size, limbs, C.c_double(scale), int(is_ntt)))
        return self

    def upload_ptr(self, host_ptr, size, limbs, scale, is_ntt=True):
        """upload from a raw (e.g. pinned) host pointer"""
        _ck(_L.bk_ct_upload(self.h, C.c_void_p(host_ptr), size, limbs, C.c_double(scale), int(is_ntt)))
        return self

    def download_ptr(self, host_ptr):
        _ck(_L.bk_ct_download(self.h, C.c_void_p(host_ptr)))

    def download(self):
        size, limbs, _, _ = self.info()
        out = np.zeros((size, limbs, self.ctx.n), dtype=np.uint64)
        _ck(_L.bk_ct_download(self.h, _ptr(out)))
        return out

    def copy(self):
        c = Ciphertext(self.ctx)
        _ck(_L.bk_ct_copy(c.h, self.h))
        return c

    def resize(self, size, limbs):
        _ck(_L.bk_ct_resize(self.h, size, limbs))

    def device_ptr(self):
        p = C.c_void_p()
        _ck(_L.bk_ct_device_ptr(self.h, C.byref(p)))
        return p.value


class Plaintext:
    def __init__(self, ctx):
        self.ctx = ctx
        self.h = C.c_void_p()
        _ck(_L.bk_pt_create(ctx.h, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            _L.bk_pt_destroy(self.h)
            self.h = None

    def info(self):
        limbs = C.c_int()
        scale = C.c_double()
        _ck(_L.bk_pt_info(self.h, C.byref(limbs), C.byref(scale)))
        return limbs.value, scale.value

    limbs = property(lambda s: s.info()[0])

    @property
    def scale(self):
        return self.info()[1]

    @scale.setter
    def scale(self, v):
        _ck(_L.bk_pt_set_scale(self.h, C.c_double(v)))

    def upload(self, data, scale):
        d = np.ascontiguousarray(data, dtype=np.uint64)
        limbs, n = d.shape
        _ck(_L.bk_pt_upload(self.h, _ptr(d), limbs, C.c_double(scale)))
        return self

    def download(self):
        out = np.zeros((self.limbs, self.ctx.n), dtype=np.uint64)
        _ck(_L.bk_pt_download(self.h, _ptr(out)))
        return out

    def mod_switch_to(self, limbs):
        _ck(_L.bk_pt_mod_switch_to(self.h, limbs))

    def copy(self):
        p = Plaintext(self.ctx)
        _ck(_L.bk_pt_copy(p.h, self.h))
        return p


class KSwitchKey:
    """One std::vector<PublicKey> of the reference's KSwitchKeys (kswitchkeys.h:340)."""

    def __init__(self, ctx, h, owned=True):
        self.ctx, self.h, self.owned = ctx, h, owned

    def __del__(self):
        if getattr(self, "owned", False) and self.h and self.ctx.h:
            _L.bk_kskey_destroy(self.h)
            self.h = None

    def info(self):
        d, l = C.c_int(), C.c_int()
        b = C.c_uint64()
        _ck(_L.bk_kskey_info(self.h, C.byref(d), C.byref(l), C.byref(b)))
        return d.value, l.value, b.value

    def export_device(self, dev_ptr):
        _ck(_L.bk_kskey_export_device(self.h, C.c_void_p(dev_ptr)))

    def download(self):
        d, l, _ = self.info()
        out = np.zeros((d, 2, l + 1, self.ctx.n), dtype=np.uint64)
        _ck(_L.bk_kskey_download(self.h, _ptr(out)))
        return out


class GaloisKeys:
    def __init__(self, ctx):
        self.ctx = ctx
        self.h = C.c_void_p()
        _ck(_L.bk_gkeys_create(ctx.h, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            _L.bk_gkeys_destroy(self.h)
            self.h = None

    def set(self, elt, key):
        _ck(_L.bk_gkeys_set(self.h, C.c_uint32(elt), key.h))
        key.owned = False

    def has(self, elt):
        r = C.c_int()
        _ck(_L.bk_gkeys_has(self.h, C.c_uint32(elt), C.byref(r)))
        return bool(r.value)


class SecretKey:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            _L.bk_sk_destroy(self.h)
            self.h = None

    def download(self):
        out = np.zeros((self.ctx.n_primes, self.ctx.n), dtype=np.uint64)
        _ck(_L.bk_sk_download(self.h, _ptr(out)))
        return out


class Context:
    """SEALContext + Evaluator + CKKSEncoder (+ key generation) on one GPU."""

    def __init__(self, log_n, primes, device=0):
        p = np.ascontiguousarray(primes, dtype=np.uint64)
        self.h = C.c_void_p()
        _ck(_L.bk_context_create(log_n, _ptr(p), len(p), device, C.byref(self.h)))
        self.log_n, self.n, self.n_primes = log_n, 1 << log_n, len(p)
        self.primes = p.copy()
        self.slots = self.n // 2
        self.top_limbs = self.n_primes - 1

    def close(self):
        if self.h:
            _L.bk_context_destroy(self.h)
            self.h = None

    # -- plumbing
    def sync(self):
        _ck(_L.bk_sync(self.h))

    def stream(self):
        s = C.c_void_p()
        _ck(_L.bk_stream(self.h, C.byref(s)))
        return s.value or 0

    def launch_count(self):
        n = C.c_uint64()
        _ck(_L.bk_launch_count(self.h, C.byref(n)))
        return n.value

    def timer_begin(self):
        _ck(_L.bk_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_double()
        _ck(_L.bk_timer_end(self.h, C.byref(ms)))
        return ms.value

    def profile_begin(self, tag):
        _ck(_L.bk_profile_begin(self.h, KERNEL_TAGS[tag] if isinstance(tag, str) else tag))

    def profile_end(self):
        n, ms = C.c_uint64(), C.c_double()
        _ck(_L.bk_profile_end(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def flush_l2(self):
        _ck(_L.bk_flush_l2(self.h))

    def import_kskey_device(self, dev_ptr, digits, limbs):
        h = C.c_void_p()
        _ck(_L.bk_kskey_import_device(self.h, C.c_void_p(dev_ptr), digits, limbs, C.byref(h)))
        return KSwitchKey(self, h)

    def set_hybrid(self, on=True):
        """Level-aware hybrid key switching for keys generated from now on (tolerance mode, include/b200ckks.h)."""
        _ck(_L.bk_context_set_hybrid(self.h, int(bool(on))))

    def set_key_compression(self, on=True):
        """level keys generated afterwards keep only their non-uniform halves resident (bk_context_set_key_compression)"""
        _ck(_L.bk_context_set_key_compression(self.h, int(bool(on))))

    def hybrid_info(self):
        on, nbytes, keys = C.c_int(), C.c_uint64(), C.c_uint64()
        _ck(_L.bk_context_hybrid(self.h, C.byref(on), C.byref(nbytes), C.byref(keys)))
        return bool(on.value), nbytes.value, keys.value

    def set_ks_chunk(self, chunk):
        _ck(_L.bk_context_set_ks_chunk(self.h, chunk))

    # -- containers
    def ciphertext(self, data=None, scale=1.0, is_ntt=True):
        c = Ciphertext(self)
        if data is not None:
            c.upload(data, scale, is_ntt)
        return c

    def plaintext(self, data=None, scale=1.0):
        p = Plaintext(self)
        if data is not None:
            p.upload(data, scale)
        return p

    # -- keys
    def upload_kskey(self, data, max_limbs=0):
        d = np.ascontiguousarray(data, dtype=np.uint64)
        digits = d.shape[0]
        assert d.shape[1:] == (2, self.n_primes, self.n), d.shape
        h = C.c_void_p()
        _ck(_L.bk_kskey_upload(self.h, _ptr(d), digits, max_limbs, C.byref(h)))
        return KSwitchKey(self, h)

    def galois_keys(self):
        return GaloisKeys(self)

    def generate_secret_key(self, hamming_weight=192, seed=1):
        h = C.c_void_p()
        _ck(_L.bk_sk_generate(self.h, hamming_weight, C.c_uint64(seed), C.byref(h)))
        return SecretKey(self, h)

    def upload_secret_key(self, data):
        d = np.ascontiguousarray(data, dtype=np.uint64)
        assert d.shape == (self.n_primes, self.n)
        h = C.c_void_p()
        _ck(_L.bk_sk_upload(self.h, _ptr(d), C.byref(h)))
        return SecretKey(self, h)

    def create_public_key(self, sk, seed=2):
        pk = Ciphertext(self)
        _ck(_L.bk_pk_generate(self.h, sk.h, C.c_uint64(seed), pk.h))
        return pk

    def create_relin_key(self, sk, seed=3, max_limbs=0):
        h = C.c_void_p()
        _ck(_L.bk_relin_key_generate(self.h, sk.h, C.c_uint64(seed), max_limbs, C.byref(h)))
        return KSwitchKey(self, h)

    def create_galois_key(self, sk, elt, seed=4, max_limbs=0):
        h = C.c_void_p()
        _ck(_L.bk_galois_key_generate(self.h, sk.h, C.c_uint32(elt), C.c_uint64(seed), max_limbs, C.byref(h)))
        return KSwitchKey(self, h)

    def create_galois_keys(self, sk, steps, seed=4, max_limbs=0):
        """KeyGenerator::create_galois_keys(steps, gk) (keygenerator.h:213); step 0 = conjugation."""
        gk = GaloisKeys(self)
        for i, st in enumerate(steps):
            elt = galois_elt_from_step(self.log_n, st)
            if not gk.has(elt):
                gk.set(elt, self.create_galois_key(sk, elt, seed + 7919 * i, max_limbs))
        return gk

    # -- encoder
    def encode(self, values, limbs, scale, top_dropped=False):
        v = np.asarray(values)
        pt = Plaintext(self)
        fn = _L.bk_encode_top_dropped if top_dropped else _L.bk_encode
        if np.iscomplexobj(v):
            v = np.ascontiguousarray(v, dtype=np.complex128)
            _ck(fn(self.h, _ptr(v.view(np.float64)), len(v), 1, limbs, C.c_double(scale), pt.h))
        else:
            v = np.ascontiguousarray(v, dtype=np.float64)
            _ck(fn(self.h, _ptr(v), len(v), 0, limbs, C.c_double(scale), pt.h))
        return pt

    def encode_scalar(self, value, limbs, scale):
        pt = Plaintext(self)
        _ck(_L.bk_encode_scalar(self.h, C.c_double(value), limbs, C.c_double(scale), pt.h))
        return pt

    def decode(self, pt):
        out = np.zeros(self.slots, dtype=np.complex128)
        _ck(_L.bk_decode(self.h, pt.h, _ptr(out.view(np.float64))))
        return out

    def set_sparse_slots(self, s):
        _ck(_L.bk_set_sparse_slots(self.h, s))

    # -- encryptor / decryptor
    def encrypt(self, pk, pt, seed=5):
        ct = Ciphertext(self)
        _ck(_L.bk_encrypt(self.h, pk.h, pt.h, C.c_uint64(seed), ct.h))
        return ct

    def encrypt_symmetric(self, sk, pt, seed=6):
        ct = Ciphertext(self)
        _ck(_L.bk_encrypt_symmetric(self.h, sk.h, pt.h, C.c_uint64(seed), ct.h))
        return ct

    def decrypt(self, sk, ct):
        pt = Plaintext(self)
        _ck(_L.bk_decrypt(self.h, sk.h, ct.h, pt.h))
        return pt

    # -- evaluator (all in place on `a`)
    def add_inplace(self, a, b):
        _ck(_L.bk_add_inplace(self.h, a.h, b.h))

    def sub_inplace(self, a, b):
        _ck(_L.bk_sub_inplace(self.h, a.h, b.h))

    def negate_inplace(self, a):
        _ck(_L.bk_negate_inplace(self.h, a.h))

    def multiply_inplace(self, a, b):
        _ck(_L.bk_multiply_inplace(self.h, a.h, b.h))

    def square_inplace(self, a):
        _ck(_L.bk_square_inplace(self.h, a.h))

    def relinearize_inplace(self, a, rk):
        _ck(_L.bk_relinearize_inplace(self.h, a.h, rk.h))

    def relinearize_rescale_inplace(self, a, rk):
        """relinearize_inplace + rescale_to_next_inplace as one call; one division by q_last * P_S in hybrid mode at a
        level with idle primes (bk_relinearize_rescale_inplace), exactly the two calls elsewhere"""
        _ck(_L.bk_relinearize_rescale_inplace(self.h, a.h, rk.h))

    def rescale_to_next_inplace(self, a):
        _ck(_L.bk_rescale_to_next_inplace(self.h, a.h))

    def mod_switch_to_next_inplace(self, a):
        _ck(_L.bk_mod_switch_to_next_inplace(self.h, a.h))

    def mod_switch_to_inplace(self, a, limbs):
        _ck(_L.bk_mod_switch_to_inplace(self.h, a.h, limbs))

    def apply_galois_inplace(self, a, elt, gk):
        _ck(_L.bk_apply_galois_inplace(self.h, a.h, C.c_uint32(elt), gk.h))

    def rotate_vector_inplace(self, a, steps, gk):
        _ck(_L.bk_rotate_vector_inplace(self.h, a.h, steps, gk.h))

    def apply_galois_hoisted(self, a, elts, gk):
        """outs[k] = apply_galois(a, elts[k]) with one shared decomposition (tolerance mode, see include/b200ckks.h)"""
        outs = [Ciphertext(self) for _ in elts]
        arr = (C.c_uint32 * len(elts))(*elts)
        hs = (C.c_void_p * len(elts))(*[o.h for o in outs])
        _ck(_L.bk_apply_galois_hoisted(self.h, a.h, arr, len(elts), gk.h, hs))
        return outs

    def complex_conjugate_inplace(self, a, gk):
        _ck(_L.bk_complex_conjugate_inplace(self.h, a.h, gk.h))

    def add_plain_inplace(self, a, p):
        _ck(_L.bk_add_plain_inplace(self.h, a.h, p.h))

    def sub_plain_inplace(self, a, p):
        _ck(_L.bk_sub_plain_inplace(self.h, a.h, p.h))

    def multiply_plain_inplace(self, a, p):
        _ck(_L.bk_multiply_plain_inplace(self.h, a.h, p.h))

    def transform_to_ntt_inplace(self, a):
        _ck(_L.bk_transform_to_ntt_inplace(self.h, a.h))

    def transform_from_ntt_inplace(self, a):
        _ck(_L.bk_transform_from_ntt_inplace(self.h, a.h))

    def add_const_inplace(self, a, v):
        _ck(_L.bk_add_const_inplace(self.h, a.h, C.c_double(v)))

    def multiply_const_inplace(self, a, v):
        _ck(_L.bk_multiply_const_inplace(self.h, a.h, C.c_double(v)))

    def modraise_inplace(self, a):
        _ck(_L.bk_modraise_inplace(self.h, a.h))

    def scalar_linear_combination(self, cts, values, constant, target_scale):
        """constant + sum_j values[j] * cts[j] at the lowest level among cts and at scale target_scale (one pass;
        term j's scalar is encoded at target_scale / cts[j].scale) - bk_scalar_linear_combination"""
        out = Ciphertext(self)
        hs = (C.c_void_p * len(cts))(*[c.h for c in cts])
        vs = (C.c_double * len(cts))(*[float(v) for v in values])
        _ck(_L.bk_scalar_linear_combination(self.h, out.h, hs, vs, len(cts), C.c_double(constant), C.c_double(target_scale)))
        return out

    # -- raw kernels
    def ntt_limbs_host(self, data, prime_idx, inverse=False):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        idx = (C.c_int * len(prime_idx))(*[int(i) for i in prime_idx])
        _ck(_L.bk_ntt_limbs_host(self.h, _ptr(d), idx, len(prime_idx), int(inverse)))
        return d

    def ntt_limbs_device(self, dev_ptr, prime_idx, inverse=False):
        idx = (C.c_int * len(prime_idx))(*[int(i) for i in prime_idx])
        _ck(_L.bk_ntt_limbs(self.h, C.c_void_p(dev_ptr), idx, len(prime_idx), int(inverse)))
