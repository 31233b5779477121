"""ctypes loader for oracle/libhhe_oracle.so (our CPU restatement) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

See oracle/hhe_oracle.h for scope and pinning. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libhhe_oracle.so")
_u64p = C.POINTER(C.c_uint64)
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hhe_oracle.c")
        if not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(src):
            build()
        l = C.CDLL(_PATH)
        l.hor_create.restype = C.c_void_p
        l.hor_create.argtypes = [C.c_uint64, C.c_uint64, _u64p, C.c_int]
        l.hor_galois_elt.restype = C.c_uint32
        _lib = l
    return _lib


class OracleError(RuntimeError):
    pass


class Oracle:
    def __init__(self, N, t, q):
        self.q = np.ascontiguousarray(q, dtype=np.uint64)
        self.N, self.t, self.K, self.L = int(N), int(t), len(self.q), len(self.q) - 1
        h = lib().hor_create(self.N, self.t, _p(self.q), self.K)
        if not h:
            raise OracleError("hor_create failed")
        self.h = C.c_void_p(h)

    def close(self):
        if self.h:
            lib().hor_destroy(self.h)
            self.h = None

    def _ct(self, size=2):
        return np.zeros((size, self.L, self.N), dtype=np.uint64)

    @staticmethod
    def _c(a):
        return _p(np.ascontiguousarray(a, dtype=np.uint64))

    def _chk(self, rc, what):
        if rc:
            raise OracleError(what + ": missing key or invalid argument")

    def ntt_roots(self):
        out = np.zeros(self.K + 1, dtype=np.uint64)
        lib().hor_ntt_roots(self.h, _p(out))
        return out[: self.K].copy(), int(out[self.K])

    def behz(self):
        out = np.zeros(3 + 2 * self.L + 1, dtype=np.uint64)
        lib().hor_behz(self.h, _p(out))
        return dict(m_sk=int(out[0]), gamma=int(out[1]), m_tilde=int(out[2]), base_B=out[3 : 3 + self.L].copy(),
                    bsk_roots=out[3 + self.L :].copy())

    def galois_elt(self, step):
        return int(lib().hor_galois_elt(self.h, step))

    def load_ksk(self, kind, elt, data):
        d = np.ascontiguousarray(data, dtype=np.uint64)
        assert d.size == self.L * 2 * self.K * self.N
        self._chk(lib().hor_load_ksk(self.h, kind, C.c_uint32(elt), _p(d)), "load_ksk")

    def ntt(self, limb, data, inverse=False):
        d = np.array(data, dtype=np.uint64, copy=True)
        lib().hor_ntt(self.h, limb, int(inverse), _p(d))
        return d

    def encode(self, slots):
        s = np.ascontiguousarray(slots, dtype=np.uint64)
        out = np.zeros(self.N, dtype=np.uint64)
        lib().hor_encode(self.h, _p(s), C.c_size_t(len(s)), _p(out))
        return out

    def add(self, a, b):
        o = self._ct()
        lib().hor_add(self.h, self._c(a), self._c(b), _p(o))
        return o

    def negate(self, a):
        o = self._ct()
        lib().hor_negate(self.h, self._c(a), _p(o))
        return o

    def add_plain(self, a, pt):
        o = self._ct()
        lib().hor_add_plain(self.h, self._c(a), self._c(pt), _p(o))
        return o

    def multiply_plain(self, a, pt):
        o = self._ct()
        lib().hor_multiply_plain(self.h, self._c(a), self._c(pt), _p(o))
        return o

    def apply_galois(self, a, elt, keys=0):
        o = self._ct()
        self._chk(lib().hor_apply_galois(self.h, self._c(a), C.c_uint32(elt), keys, _p(o)), "apply_galois")
        return o

    def rotate_rows(self, a, steps, keys=0):
        o = self._ct()
        self._chk(lib().hor_rotate_rows(self.h, self._c(a), steps, keys, _p(o)), "rotate_rows")
        return o

    def rotate_columns(self, a, keys=0):
        o = self._ct()
        self._chk(lib().hor_rotate_columns(self.h, self._c(a), keys, _p(o)), "rotate_columns")
        return o

    def multiply(self, a, b):
        o = self._ct(3)
        lib().hor_multiply(self.h, self._c(a), self._c(b), _p(o))
        return o

    def relinearize(self, a3):
        o = self._ct()
        self._chk(lib().hor_relinearize(self.h, self._c(a3), _p(o)), "relinearize")
        return o

    def exponentiate3(self, a):
        o = self._ct()
        self._chk(lib().hor_exponentiate3(self.h, self._c(a), _p(o)), "exponentiate3")
        return o

    def vec_sum(self, a, n, keys=1):
        o = self._ct()
        self._chk(lib().hor_vec_sum(self.h, self._c(a), C.c_size_t(n), keys, _p(o)), "vec_sum")
        return o

    def mask(self, a, mask):
        m = np.ascontiguousarray(mask, dtype=np.uint64)
        o = self._ct()
        lib().hor_mask(self.h, self._c(a), _p(m), C.c_size_t(len(m)), _p(o))
        return o

    def flatten(self, cts, keys=0):
        c = np.ascontiguousarray(cts, dtype=np.uint64)
        o = self._ct()
        self._chk(lib().hor_flatten(self.h, _p(c), C.c_size_t(c.shape[0]), keys, _p(o)), "flatten")
        return o

    def pasta_decompose(self, enc_key, sym_ct, use_bsgs=False, nonce=123456789, first_counter=0):
        s = np.ascontiguousarray(sym_ct, dtype=np.uint64)
        nblk = (len(s) + 127) // 128
        o = np.zeros((nblk, 2, self.L, self.N), dtype=np.uint64)
        self._chk(lib().hor_pasta_decompose(self.h, self._c(enc_key), _p(s), C.c_size_t(len(s)), C.c_uint64(nonce),
                                            C.c_uint64(first_counter), int(use_bsgs), _p(o)), "pasta_decompose")
        return o


def naf(value):
    t = (C.c_int * 40)()
    n = lib().hor_naf(value, t)
    return [t[i] for i in range(n)]


def pasta_layer_material(p, nonce, counter, layer):
    m1 = np.zeros((128, 128), dtype=np.uint64)
    m2 = np.zeros((128, 128), dtype=np.uint64)
    rc = np.zeros(256, dtype=np.uint64)
    lib().hor_pasta_layer_material(C.c_uint64(p), C.c_uint64(nonce), C.c_uint64(counter), layer, _p(m1), _p(m2), _p(rc))
    return m1, m2, rc


def pasta_plain(key256, p, data, decrypt=False):
    k = np.ascontiguousarray(key256, dtype=np.uint64)
    d = np.ascontiguousarray(data, dtype=np.uint64)
    o = np.zeros_like(d)
    lib().hor_pasta_plain(_p(k), C.c_uint64(p), _p(d), C.c_size_t(len(d)), int(decrypt), _p(o))
    return o


def shake128(data: bytes, outlen: int) -> bytes:
    out = (C.c_uint8 * outlen)()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if data else b"\0")
    lib().hor_shake128(buf, C.c_size_t(len(data)), out, C.c_size_t(outlen))
    return bytes(out)
