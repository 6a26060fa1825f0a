"""ctypes binding of the application-layer C ABI (include/b200ckks_app.h): bootstrapping, approximate ReLU,
multiplexed-packing CNN operators and the ResNet driver, as restated in fhe-gpt-2_b200/host/.

`App()` loads lib/libb200ckks_app.so (the engine; raises if it is missing - there is no CPU fallback).  `App(path)`
binds any other library exporting the same C ABI (the test suite uses that for its CPU checker build of the same host
sources against the reference's own SEAL).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
APP_LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libb200ckks_app.so"))
APP_HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "..", "include", "b200ckks_app.h"))

OPS = ["conv", "bn", "relu", "bootstrap", "add", "downsample", "avgpool", "fc"]
STAT_NAMES = ["key_switch_rotate", "key_switch_relin", "rescale", "multiply", "multiply_plain", "encode_vector", "add",
              "mod_switch", "scalar_op"]

_EXC = {1: ValueError, 2: RuntimeError, 3: IndexError, 4: RuntimeError}


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


KERNEL_FAMILIES = ["fwd_cols", "fwd_blocks", "inv_blocks", "inv_cols", "ks_mac", "elementwise", "fft", "other"]


class EngineView:
    """Measurement helpers of include/b200ckks.h on the bk_context_t behind a session."""

    def __init__(self, handle):
        from . import lib
        self.L, self.h = lib(), C.c_void_p(handle)

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.L.bk_last_error().decode())

    def launch_count(self):
        n = C.c_uint64()
        self._ck(self.L.bk_launch_count(self.h, C.byref(n)))
        return n.value

    def transfer_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.L.bk_transfer_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def kernel_counters(self):
        l, u = (C.c_uint64 * 8)(), (C.c_uint64 * 8)()
        self._ck(self.L.bk_kernel_counters(self.h, l, u))
        return {k: (int(l[i]), int(u[i])) for i, k in enumerate(KERNEL_FAMILIES)}

    def profile_begin_all(self):
        self._ck(self.L.bk_profile_begin(self.h, -2))

    def profile_end_all(self):
        l, ms = (C.c_uint64 * 8)(), (C.c_double * 8)()
        self._ck(self.L.bk_profile_end_all(self.h, l, ms))
        return {k: (int(l[i]), float(ms[i])) for i, k in enumerate(KERNEL_FAMILIES)}

    def timer_begin(self):
        self._ck(self.L.bk_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_double()
        self._ck(self.L.bk_timer_end(self.h, C.byref(ms)))
        return ms.value

    def imad_peak(self):
        v = C.c_double()
        self._ck(self.L.bk_measure_imad_peak(self.h, C.byref(v)))
        return v.value

    def int_pipe_rates(self):
        """thread-instructions per second of mad.lo.u32 / mad.wide.u32 / mad.hi.u32 on this GPU"""
        v = (C.c_double * 3)()
        self._ck(self.L.bk_measure_int_pipe(self.h, v))
        return dict(zip(("mad_lo", "mad_wide", "mad_hi"), (float(x) for x in v)))

    def flush_l2(self):
        self._ck(self.L.bk_flush_l2(self.h))

    def set_ks_chunk(self, chunk):
        self._ck(self.L.bk_context_set_ks_chunk(self.h, chunk))


class App:
    def __init__(self, lib_path=None):
        path = lib_path or APP_LIB_PATH
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        self.L = C.CDLL(path)
        self.L.bka_last_error.restype = C.c_char_p
        self.L.bka_backend.restype = C.c_char_p

    def ck(self, rc):
        if rc:
            raise _EXC.get(rc, RuntimeError)(self.L.bka_last_error().decode())

    @property
    def backend(self):
        return self.L.bka_backend().decode()

    def session(self, log_n, bits, hamming_weight=192, device=0, rotation_steps=(), secret_key=None):
        return Session(self, log_n, bits, hamming_weight, device, rotation_steps, secret_key)

    def gpt2_init_chain(self):
        """(bit sizes, rotation steps) of the reference's GPT-2 INIT macro (gpt2/util.h:37-75)."""
        bits, steps = np.zeros(64, dtype=np.int32), np.zeros(256, dtype=np.int32)
        nb, ns = C.c_int(), C.c_int()
        self.ck(self.L.bka_gpt2_init_chain(_iptr(bits), len(bits), C.byref(nb), _iptr(steps), len(steps), C.byref(ns)))
        return bits[:nb.value].tolist(), steps[:ns.value].tolist()

    def oddbaby_tree(self, deg):
        buf = np.zeros(4096, dtype=np.int32)
        n, depth, m, l = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.ck(self.L.bka_oddbaby_tree(deg, _iptr(buf), len(buf), C.byref(n), C.byref(depth), C.byref(m), C.byref(l)))
        return buf[:n.value].tolist(), depth.value, m.value, l.value


class Ct:
    def __init__(self, sess, h):
        self.s, self.h = sess, h

    def __del__(self):
        if getattr(self, "h", None) and self.s.h:
            self.s.app.L.bka_ct_free(self.h)
            self.h = None

    def info(self):
        size, limbs, scale = C.c_int(), C.c_int(), C.c_double()
        self.s.app.ck(self.s.app.L.bka_ct_info(self.h, C.byref(size), C.byref(limbs), C.byref(scale)))
        return size.value, limbs.value, scale.value

    limbs = property(lambda s: s.info()[1])

    @property
    def scale(self):
        return self.info()[2]

    @scale.setter
    def scale(self, v):
        self.s.app.ck(self.s.app.L.bka_ct_set_scale(self.h, C.c_double(v)))

    def clone(self):
        out = C.c_void_p()
        self.s.app.ck(self.s.app.L.bka_ct_clone(self.h, C.byref(out)))
        return Ct(self.s, out)

    def download(self):
        size, limbs, _ = self.info()
        out = np.zeros((size, limbs, 1 << self.s.log_n), dtype=np.uint64)
        self.s.app.ck(self.s.app.L.bka_ct_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out


class Session:
    """EncryptionParameters + SEALContext + keys + encoder/encryptor/evaluator/decryptor (infer_seal.cpp:288-342)."""

    def __init__(self, app, log_n, bits, hamming_weight, device, rotation_steps, secret_key=None):
        self.app, self.log_n, self.bits = app, log_n, list(bits)
        self.slots = 1 << (log_n - 1)
        self.top_limbs = len(self.bits) - 1      # data limbs of a fresh ciphertext
        arr = (C.c_int * len(bits))(*bits)
        st = (C.c_int * max(1, len(rotation_steps)))(*rotation_steps)
        self.h = C.c_void_p()
        sk = None
        if secret_key is not None:
            sk = np.ascontiguousarray(secret_key, dtype=np.uint64)
            assert sk.shape == (len(bits), 1 << log_n)
        app.ck(app.L.bka_session_create_with_secret(log_n, arr, len(bits), hamming_weight, device, st, len(rotation_steps),
                                                    sk.ctypes.data_as(C.c_void_p) if sk is not None else None,
                                                    C.byref(self.h)))

    def secret_key(self):
        out = np.zeros((len(self.bits), 1 << self.log_n), dtype=np.uint64)
        self.app.ck(self.app.L.bka_session_secret_key(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def engine(self):
        p = C.c_void_p()
        self.app.ck(self.app.L.bka_session_engine_context(self.h, C.byref(p)))
        if not p.value:
            raise RuntimeError("this backend has no engine context")
        return EngineView(p.value)

    def level_histogram(self, which, reset=False):
        out = (C.c_uint64 * 64)()
        self.app.ck(self.app.L.bka_session_level_histogram(self.h, which, out, int(reset)))
        return [int(x) for x in out]

    def close(self):
        if self.h:
            self.app.L.bka_session_destroy(self.h)
            self.h = None

    def add_rotation_steps(self, steps):
        st = (C.c_int * len(steps))(*steps)
        self.app.ck(self.app.L.bka_session_add_rotation_steps(self.h, st, len(steps)))

    def primes(self):
        out = np.zeros(len(self.bits), dtype=np.uint64)
        self.app.ck(self.app.L.bka_session_primes(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def sync(self):
        self.app.ck(self.app.L.bka_session_sync(self.h))

    def stats(self, reset=False):
        out = (C.c_uint64 * 9)()
        self.app.ck(self.app.L.bka_session_stats(self.h, out, int(reset)))
        return dict(zip(STAT_NAMES, [int(x) for x in out]))

    def key_residency(self):
        b, g = C.c_uint64(), C.c_uint64()
        self.app.ck(self.app.L.bka_session_key_residency(self.h, C.byref(b), C.byref(g)))
        return b.value, g.value

    def double_hoisted_groups(self):
        n = C.c_uint64()
        self.app.ck(self.app.L.bka_session_double_hoisted_groups(self.h, C.byref(n)))
        return n.value

    def key_plan(self):
        """text of the (Galois element, level) pairs and relinearization levels the keys cover so far"""
        n = C.c_int()
        self.app.ck(self.app.L.bka_session_key_plan(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        self.app.ck(self.app.L.bka_session_key_plan(self.h, buf, n.value + 1, C.byref(n)))
        return buf.value.decode()

    def apply_key_plan(self, text, detach_secret=True):
        """generate exactly the keys of a plan now; detach_secret: the evaluation keys forget the secret key"""
        self.app.ck(self.app.L.bka_session_apply_key_plan(self.h, text.encode(), int(detach_secret)))

    def plain_cache(self):
        b, h, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.app.ck(self.app.L.bka_session_plain_cache(self.h, C.byref(b), C.byref(h), C.byref(m)))
        return dict(bytes=b.value, hits=h.value, misses=m.value)

    # -- ciphertexts
    def encrypt(self, values, scale, limbs=0):
        v = np.asarray(values)
        out = C.c_void_p()
        if np.iscomplexobj(v):
            v = np.ascontiguousarray(v, dtype=np.complex128)
            self.app.ck(self.app.L.bka_encrypt(self.h, _dptr(v.view(np.float64)), len(v), 1, C.c_double(scale), limbs,
                                               C.byref(out)))
        else:
            v = np.ascontiguousarray(v, dtype=np.float64)
            self.app.ck(self.app.L.bka_encrypt(self.h, _dptr(v), len(v), 0, C.c_double(scale), limbs, C.byref(out)))
        return Ct(self, out)

    def decrypt(self, ct):
        out = np.zeros(self.slots, dtype=np.complex128)
        self.app.ck(self.app.L.bka_decrypt(self.h, ct.h, _dptr(out.view(np.float64))))
        return out

    def mod_switch_to(self, ct, limbs):
        self.app.ck(self.app.L.bka_ct_mod_switch_to(self.h, ct.h, limbs))

    def rotate(self, ct, steps):
        self.app.ck(self.app.L.bka_rotate(self.h, ct.h, steps))

    def multiply_relin_rescale(self, a, b):
        self.app.ck(self.app.L.bka_multiply_relin_rescale(self.h, a.h, b.h))

    def add_reduced_error(self, a, b):
        self.app.ck(self.app.L.bka_add_reduced_error(self.h, a.h, b.h))

    def reduced_error_op(self, which, a, b):
        """a <- a {add, sub, multiply}_inplace_reduced_error b (evaluator.cpp:312-486)"""
        self.app.ck(self.app.L.bka_reduced_error_op(self.h, {"add": 0, "sub": 1, "multiply": 2}[which], a.h, b.h))

    def upload(self, limbs_array, scale, ntt=True):
        """ciphertext from raw limbs [size][limbs][N] in the reference's layout"""
        d = np.ascontiguousarray(limbs_array, dtype=np.uint64)
        out = C.c_void_p()
        self.app.ck(self.app.L.bka_ct_upload(self.h, d.ctypes.data_as(C.c_void_p), d.shape[0], d.shape[1], C.c_double(scale),
                                             int(ntt), C.byref(out)))
        return Ct(self, out)

    def import_relin_key(self, key):
        """relinearization key in SEAL's layout [digits][2][n_primes][N]"""
        d = np.ascontiguousarray(key, dtype=np.uint64)
        self.app.ck(self.app.L.bka_session_import_relin_key(self.h, d.ctypes.data_as(C.c_void_p), d.shape[0]))

    WHAT = dict(ciphertext=0, relin_keys=2, galois_keys=3, secret_key=4, public_key=5)

    def save(self, what, path, ct=None):
        """SEAL 3.6 wire format (compr_mode none) of a ciphertext or of the session's keys, into a file"""
        self.app.ck(self.app.L.bka_save(self.h, self.WHAT[what], ct.h if ct is not None else None, path.encode()))

    def load(self, what, path):
        out = C.c_void_p()
        self.app.ck(self.app.L.bka_load(self.h, self.WHAT[what], path.encode(), C.byref(out)))
        return Ct(self, out) if what == "ciphertext" else None

    def multiply_vector_rescale(self, a, values):
        v = np.asarray(values)
        if np.iscomplexobj(v):
            v = np.ascontiguousarray(v, dtype=np.complex128)
            self.app.ck(self.app.L.bka_multiply_vector_rescale(self.h, a.h, _dptr(v.view(np.float64)), len(v), 1))
        else:
            v = np.ascontiguousarray(v, dtype=np.float64)
            self.app.ck(self.app.L.bka_multiply_vector_rescale(self.h, a.h, _dptr(v), len(v), 0))

    # -- application layers
    def bootstrapper(self, logn, loge=10, total_level=30, final_scale=2.0 ** 46, boundary_k=25, sin_cos_deg=59,
                     scale_factor=2, inverse_deg=1):
        return Bootstrapper(self, loge, logn, total_level, final_scale, boundary_k, sin_cos_deg, scale_factor, inverse_deg)

    def relu(self, ct):
        out = C.c_void_p()
        self.app.ck(self.app.L.bka_relu(self.h, ct.h, C.byref(out)))
        return Ct(self, out)

    def conv(self, ct, parms, co, st, weight, running_var, bn_weight, epsilon=1e-5, fh=3, fw=3, end=False):
        out, op = C.c_void_p(), (C.c_int * 7)()
        ip = (C.c_int * 7)(*parms)
        w = np.ascontiguousarray(weight, dtype=np.float64).reshape(-1)
        rv = np.ascontiguousarray(running_var, dtype=np.float64)
        bw = np.ascontiguousarray(bn_weight, dtype=np.float64)
        self.app.ck(self.app.L.bka_conv(self.h, ct.h, ip, co, st, fh, fw, _dptr(w), _dptr(rv), _dptr(bw),
                                        C.c_double(epsilon), int(end), C.byref(out), op))
        return Ct(self, out), list(op)

    def bn(self, ct, parms, bias, mean, var, weight, epsilon=1e-5, B=40.0):
        out = C.c_void_p()
        ip = (C.c_int * 7)(*parms)
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (bias, mean, var, weight)]
        self.app.ck(self.app.L.bka_bn(self.h, ct.h, ip, _dptr(a[0]), _dptr(a[1]), _dptr(a[2]), _dptr(a[3]),
                                      C.c_double(epsilon), C.c_double(B), C.byref(out)))
        return Ct(self, out)

    def downsample(self, ct, parms):
        out, op = C.c_void_p(), (C.c_int * 7)()
        self.app.ck(self.app.L.bka_downsample(self.h, ct.h, (C.c_int * 7)(*parms), C.byref(out), op))
        return Ct(self, out), list(op)

    def avgpool(self, ct, parms, B=40.0):
        out, op = C.c_void_p(), (C.c_int * 7)()
        self.app.ck(self.app.L.bka_avgpool(self.h, ct.h, (C.c_int * 7)(*parms), C.c_double(B), C.byref(out), op))
        return Ct(self, out), list(op)

    def fc(self, ct, parms, matrix, bias, q, r):
        out = C.c_void_p()
        m = np.ascontiguousarray(matrix, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(bias, dtype=np.float64)
        self.app.ck(self.app.L.bka_fc(self.h, ct.h, (C.c_int * 7)(*parms), _dptr(m), _dptr(b), q, r, C.byref(out)))
        return Ct(self, out)

    def tensor_add(self, a, b):
        out = C.c_void_p()
        self.app.ck(self.app.L.bka_tensor_add(self.h, a.h, b.h, C.byref(out)))
        return Ct(self, out)

    def resnet(self, layer_num, weights):
        return ResNet(self, layer_num, weights)

    def gpt2(self, op, cts=(), i=(), d=(), boot=None, out_cap=256):
        """One GPT-2 operator of gpt2/approx.h by name (see include/b200ckks_app.h); returns a list of Ct."""
        ins = (C.c_void_p * max(1, len(cts)))(*[c.h for c in cts])
        ip = np.ascontiguousarray(np.asarray(i, dtype=np.int32).ravel())
        dp = np.ascontiguousarray(np.asarray(d, dtype=np.float64).ravel())
        outs = (C.c_void_p * out_cap)()
        n = C.c_int()
        self.app.ck(self.app.L.bka_gpt2_call(self.h, boot.h if boot is not None else None, op.encode(), ins, len(cts),
                                             _dptr(dp) if dp.size else None, int(dp.size), _iptr(ip) if ip.size else None,
                                             int(ip.size), outs, out_cap, C.byref(n)))
        return [Ct(self, C.c_void_p(outs[k])) for k in range(n.value)]


class Bootstrapper:
    def __init__(self, sess, loge, logn, total_level, final_scale, boundary_k, sin_cos_deg, scale_factor, inverse_deg):
        self.s, self.logn = sess, logn
        self.h = C.c_void_p()
        sess.app.ck(sess.app.L.bka_bootstrapper_create(sess.h, loge, logn, total_level, C.c_double(final_scale), boundary_k,
                                                       sin_cos_deg, scale_factor, inverse_deg, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and self.s.h:
            self.s.app.L.bka_bootstrapper_destroy(self.h)
            self.h = None

    def set_hoisting(self, on, double=True):
        """hoisted baby steps on / off; double=False keeps one ModDown per baby rotation (no double hoisting)"""
        prev = C.c_int()
        self.s.app.ck(self.s.app.L.bka_bootstrapper_set_hoisting(self.h, (1 if on else 0) | (0 if double else 2), C.byref(prev)))
        return bool(prev.value & 1)

    def rotation_steps(self):
        buf = np.zeros(4096, dtype=np.int32)
        n = C.c_int()
        self.s.app.ck(self.s.app.L.bka_bootstrapper_rotation_steps(self.h, _iptr(buf), len(buf), C.byref(n)))
        return buf[:n.value].tolist()

    def lt_coefficients(self, which):
        nd, ln = C.c_int(), C.c_int()
        self.s.app.ck(self.s.app.L.bka_bootstrapper_lt_coefficients(self.h, which, C.byref(nd), C.byref(ln), None))
        out = np.zeros((nd.value, ln.value), dtype=np.complex128)
        self.s.app.ck(self.s.app.L.bka_bootstrapper_lt_coefficients(self.h, which, C.byref(nd), C.byref(ln),
                                                                    _dptr(out.view(np.float64))))
        return out

    def bootstrap(self, ct, real_message=True):
        out = C.c_void_p()
        self.s.app.ck(self.s.app.L.bka_bootstrap(self.h, ct.h, int(real_message), C.byref(out)))
        return Ct(self.s, out)

    def modular_reduction(self, ct):
        out = C.c_void_p()
        self.s.app.ck(self.s.app.L.bka_modular_reduction(self.h, ct.h, C.byref(out)))
        return Ct(self.s, out)


class ResNet:
    """weights: dict with conv_weight / bn_bias / bn_mean / bn_var / bn_weight (lists of arrays, layer_num - 1 each),
    linear_weight (10 x 64), linear_bias (10)."""

    def __init__(self, sess, layer_num, weights):
        self.s = sess
        cat = lambda k: np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in weights[k]]))
        self._keep = [cat("conv_weight"), cat("bn_bias"), cat("bn_mean"), cat("bn_var"), cat("bn_weight"),
                      np.ascontiguousarray(weights["linear_weight"], dtype=np.float64).reshape(-1),
                      np.ascontiguousarray(weights["linear_bias"], dtype=np.float64)]
        self.h = C.c_void_p()
        if "shortcut_weight" in weights:      # CIFAR-100 variant: 1x1 stride-2 shortcut convolutions, 100 classes, B = 65
            self._keep += [cat("shortcut_weight"), cat("shortcut_bn_bias"), cat("shortcut_bn_mean"), cat("shortcut_bn_var"),
                           cat("shortcut_bn_weight")]
            sess.app.ck(sess.app.L.bka_resnet_create_cifar100(sess.h, layer_num, *[_dptr(a) for a in self._keep], C.byref(self.h)))
        else:
            sess.app.ck(sess.app.L.bka_resnet_create(sess.h, layer_num, *[_dptr(a) for a in self._keep], C.byref(self.h)))
        n = C.c_int()
        sess.app.ck(sess.app.L.bka_resnet_classes(self.h, C.byref(n)))
        self.classes = n.value

    def __del__(self):
        if getattr(self, "h", None) and self.s.h:
            self.s.app.L.bka_resnet_destroy(self.h)
            self.h = None

    def encrypt_image(self, image):
        img = np.ascontiguousarray(image, dtype=np.float64).reshape(-1)
        assert img.size == 3072
        out = C.c_void_p()
        self.s.app.ck(self.s.app.L.bka_resnet_encrypt_image(self.h, _dptr(img), C.byref(out)))
        return Ct(self.s, out)

    def infer_encrypted(self, ct, trace=False):
        out, rows = C.c_void_p(), C.c_int()
        cap = 4096
        tr = np.zeros((cap, 4))
        self.s.app.ck(self.s.app.L.bka_resnet_infer_encrypted(self.h, ct.h, C.byref(out), _dptr(tr) if trace else None, cap,
                                                              C.byref(rows)))
        t = [dict(op=OPS[int(r[0])], level=int(r[1]), scale=float(r[2]), ms=float(r[3])) for r in tr[:rows.value]]
        return Ct(self.s, out), t

    def decrypt_logits(self, ct):
        logits = np.zeros(self.classes)
        self.s.app.ck(self.s.app.L.bka_resnet_decrypt_logits(self.h, ct.h, _dptr(logits)))
        return logits

    def infer(self, image, trace=True):
        img = np.ascontiguousarray(image, dtype=np.float64).reshape(-1)
        assert img.size == 3072
        logits = np.zeros(self.classes)
        cap = 4096
        tr = np.zeros((cap, 4))
        rows = C.c_int()
        self.s.app.ck(self.s.app.L.bka_resnet_infer(self.h, _dptr(img), _dptr(logits), _dptr(tr) if trace else None, cap,
                                                    C.byref(rows)))
        t = [dict(op=OPS[int(r[0])], level=int(r[1]), scale=float(r[2]), ms=float(r[3])) for r in tr[:rows.value]]
        return logits, t

    def infer_batch(self, images, in_flight):
        """bka_resnet_infer_batch: images (n, 3072) on the host -> logits (n, 10); `in_flight` images at a time, one host
        thread and CUDA stream each (the reference's OpenMP image loop, infer_seal.cpp:404)."""
        imgs = np.ascontiguousarray(images, dtype=np.float64).reshape(-1, 3072)
        logits = np.zeros((imgs.shape[0], self.classes))
        self.s.app.ck(self.s.app.L.bka_resnet_infer_batch(self.h, _dptr(imgs), imgs.shape[0], int(in_flight), _dptr(logits)))
        return logits

    def infer_encrypted_batch(self, cts, in_flight):
        """bka_resnet_infer_encrypted_batch: inputs and outputs stay in HBM."""
        ins = (C.c_void_p * len(cts))(*[c.h for c in cts])
        outs = (C.c_void_p * len(cts))()
        self.s.app.ck(self.s.app.L.bka_resnet_infer_encrypted_batch(self.h, ins, len(cts), int(in_flight), outs))
        return [Ct(self.s, C.c_void_p(outs[k])) for k in range(len(cts))]
