"""ctypes loader for oracle/_ref/libhhe_ref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library is the UNMODIFIED reference hot path (src/pasta/*.cpp, src/util/sealhelper.cpp, libs/keccak,
vendored libseal-4.0.a) behind the flat C ABI of oracle/ref_shim.cpp. It is built in the container by
`make -C oracle ref` and travels to the GPU box as a prebuilt .so. Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libhhe_ref.so")
_u64p = C.POINTER(C.c_uint64)


def available():
    return os.path.exists(_PATH)


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(_PATH)
        l.ref_create.restype = C.c_void_p
        l.ref_create.argtypes = [C.c_uint64, C.c_uint64, _u64p, C.c_int, C.c_uint64, C.POINTER(C.c_int), C.c_int, C.c_int]
        l.ref_last_error.restype = C.c_char_p
        l.ref_galois_elt.restype = C.c_uint32
        l.ref_bench_decompose.restype = C.c_double
        l.ref_bench_primitive.restype = C.c_double
        for f in ("ref_ct_save", "ref_ct_load", "ref_keys_save"):
            getattr(l, f).restype = C.c_longlong
        _lib = l
    return _lib


class RefError(RuntimeError):
    pass


class Ref:
    """One SEAL context + seeded key set of the reference (see ref_shim.cpp)."""

    def __init__(self, N, t=65537, q=None, seed=1, steps=(0, -1, 128), default_gk=False):
        l = lib()
        qa = np.asarray(q if q is not None else [], dtype=np.uint64)
        st = (C.c_int * max(1, len(steps)))(*steps)
        self.h = l.ref_create(N, t, _p(qa) if len(qa) else None, len(qa), seed, st, len(steps), int(default_gk))
        if not self.h:
            raise RefError(l.ref_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.seed = int(seed)
        info = np.zeros(4, dtype=np.uint64)
        l.ref_info(self.h, _p(info))
        self.N, self.L, self.K, self.t = (int(x) for x in info)
        self.q = np.zeros(self.K, dtype=np.uint64)
        l.ref_moduli(self.h, _p(self.q))
        self.ct_words = 2 * self.L * self.N

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc:
            raise RefError(lib().ref_last_error().decode())

    # ---- constants -------------------------------------------------------------------------
    def ntt_roots(self):
        out = np.zeros(self.K + 1, dtype=np.uint64)
        lib().ref_ntt_roots(self.h, _p(out))
        return out[: self.K].copy(), int(out[self.K])

    def behz(self):
        out = np.zeros(3 + 2 * self.L + 1, dtype=np.uint64)
        lib().ref_behz(self.h, _p(out))
        return dict(m_sk=int(out[0]), gamma=int(out[1]), m_tilde=int(out[2]), base_B=out[3 : 3 + self.L].copy(),
                    bsk_roots=out[3 + self.L :].copy())

    def galois_elt(self, step):
        return int(lib().ref_galois_elt(self.h, step))

    def ksk(self, kind, elt=0):
        """kind 0/1 = galois keyset, 2 = relin. Returns [L][2][K][N] or None if absent."""
        out = np.zeros((self.L, 2, self.K, self.N), dtype=np.uint64)
        rc = lib().ref_get_ksk(self.h, kind, C.c_uint32(elt), _p(out))
        return out if rc == 0 else None

    def list_galois(self, kind):
        n = lib().ref_list_galois(self.h, kind, None)
        arr = (C.c_uint32 * max(1, n))()
        lib().ref_list_galois(self.h, kind, arr)
        return [int(arr[i]) for i in range(n)]

    def public_key(self):
        out = np.zeros((2, self.K, self.N), dtype=np.uint64)
        lib().ref_public_key(self.h, _p(out))
        return out

    def prng_seed(self):
        """The prng_seed_type ref_create gives its Blake2xbPRNGFactory: every generator SEAL creates from it (one per encryption)
        starts from this seed."""
        return np.array([(self.seed * 0x9E3779B97F4A7C15 + i) % (1 << 64) for i in range(8)], dtype=np.uint64)

    def secret_key(self):
        out = np.zeros((self.K, self.N), dtype=np.uint64)
        lib().ref_secret_key(self.h, _p(out))
        return out

    # ---- SEAL wire format (checker for the codec) --------------------------------------------
    def parms_id(self, level=0):
        out = np.zeros(4, dtype=np.uint64)
        lib().ref_parms_id(self.h, level, _p(out))
        return out

    def ct_save(self, ct, compr=2):
        c = np.ascontiguousarray(ct, dtype=np.uint64)
        size = c.size // (self.L * self.N)
        buf = np.zeros(c.nbytes + 4096, dtype=np.uint8)
        n = lib().ref_ct_save(self.h, _p(c), size, compr, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size))
        if n < 0:
            raise RefError(lib().ref_last_error().decode())
        return buf[:n].tobytes()

    def ct_load(self, data):
        b = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros((3, self.L, self.N), dtype=np.uint64)
        size = C.c_int(0)
        n = lib().ref_ct_load(self.h, b.ctypes.data_as(C.c_void_p), C.c_size_t(b.size), _p(out), C.byref(size))
        if n < 0:
            raise RefError(lib().ref_last_error().decode())
        return out[: size.value].copy(), int(n)

    def keys_save(self, kind, compr=2):
        cap = lib().ref_keys_save(self.h, kind, compr, None, C.c_size_t(0))
        if cap < 0:
            raise RefError(lib().ref_last_error().decode())
        buf = np.zeros(cap, dtype=np.uint8)
        n = lib().ref_keys_save(self.h, kind, compr, buf.ctypes.data_as(C.c_void_p), C.c_size_t(cap))
        if n < 0:
            raise RefError(lib().ref_last_error().decode())
        return buf[:n].tobytes()

    def galois_load_count(self, data):
        b = np.frombuffer(data, dtype=np.uint8)
        n = lib().ref_galois_load_count(self.h, b.ctypes.data_as(C.c_void_p), C.c_size_t(b.size))
        if n < 0:
            raise RefError(lib().ref_last_error().decode())
        return n

    # ---- client side -----------------------------------------------------------------------
    def encode(self, slots):
        s = np.ascontiguousarray(slots, dtype=np.uint64)
        out = np.zeros(self.N, dtype=np.uint64)
        self._chk(lib().ref_encode(self.h, _p(s), C.c_size_t(len(s)), _p(out)))
        return out

    def encrypt(self, slots):
        s = np.ascontiguousarray(slots, dtype=np.uint64)
        out = np.zeros((2, self.L, self.N), dtype=np.uint64)
        self._chk(lib().ref_encrypt(self.h, _p(s), C.c_size_t(len(s)), _p(out)))
        return out

    def decrypt(self, ct):
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.zeros(self.N, dtype=np.uint64)
        b = C.c_int(0)
        self._chk(lib().ref_decrypt(self.h, _p(ct), ct.shape[0], _p(out), C.byref(b)))
        return out, b.value

    # ---- evaluator ops ---------------------------------------------------------------------
    def ntt(self, limb, data, inverse=False):
        d = np.array(data, dtype=np.uint64, copy=True)
        self._chk(lib().ref_ntt(self.h, limb, int(inverse), _p(d)))
        return d

    def ntt_bsk(self, idx, data, inverse=False):
        d = np.array(data, dtype=np.uint64, copy=True)
        self._chk(lib().ref_ntt_bsk(self.h, idx, int(inverse), _p(d)))
        return d

    def _ct_out(self, size=2):
        return np.zeros((size, self.L, self.N), dtype=np.uint64)

    def add(self, a, b):
        o = self._ct_out()
        self._chk(lib().ref_add(self.h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(o)))
        return o

    def negate(self, a):
        o = self._ct_out()
        self._chk(lib().ref_negate(self.h, _p(np.ascontiguousarray(a)), _p(o)))
        return o

    def add_plain(self, a, pt):
        o = self._ct_out()
        self._chk(lib().ref_add_plain(self.h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(pt)), _p(o)))
        return o

    def multiply_plain(self, a, pt):
        o = self._ct_out()
        self._chk(lib().ref_multiply_plain(self.h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(pt)), _p(o)))
        return o

    def rotate_rows(self, a, steps, keys=0):
        o = self._ct_out()
        self._chk(lib().ref_rotate_rows(self.h, _p(np.ascontiguousarray(a)), steps, keys, _p(o)))
        return o

    def rotate_columns(self, a, keys=0):
        o = self._ct_out()
        self._chk(lib().ref_rotate_columns(self.h, _p(np.ascontiguousarray(a)), keys, _p(o)))
        return o

    def multiply(self, a, b):
        o = self._ct_out(3)
        self._chk(lib().ref_multiply(self.h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(o)))
        return o

    def square(self, a):
        o = self._ct_out(3)
        self._chk(lib().ref_square(self.h, _p(np.ascontiguousarray(a)), _p(o)))
        return o

    def relinearize(self, a3):
        o = self._ct_out()
        self._chk(lib().ref_relinearize(self.h, _p(np.ascontiguousarray(a3)), _p(o)))
        return o

    def exponentiate3(self, a):
        o = self._ct_out()
        self._chk(lib().ref_exponentiate3(self.h, _p(np.ascontiguousarray(a)), _p(o)))
        return o

    def vec_sum(self, a, n, keys=1):
        o = self._ct_out()
        self._chk(lib().ref_vec_sum(self.h, _p(np.ascontiguousarray(a)), C.c_size_t(n), keys, _p(o)))
        return o

    def pasta_decompose(self, enc_key, sym_ct, use_bsgs=False):
        s = np.ascontiguousarray(sym_ct, dtype=np.uint64)
        nblk = (len(s) + 127) // 128
        o = np.zeros((nblk, 2, self.L, self.N), dtype=np.uint64)
        self._chk(lib().ref_pasta_decompose(self.h, _p(np.ascontiguousarray(enc_key)), _p(s), C.c_size_t(len(s)),
                                            int(use_bsgs), _p(o)))
        return o

    def mask(self, a, mask):
        m = np.ascontiguousarray(mask, dtype=np.uint64)
        o = self._ct_out()
        self._chk(lib().ref_mask(self.h, _p(np.ascontiguousarray(a)), _p(m), C.c_size_t(len(m)), _p(o)))
        return o

    def flatten(self, cts, keys=0):
        c = np.ascontiguousarray(cts, dtype=np.uint64)
        o = self._ct_out()
        self._chk(lib().ref_flatten(self.h, _p(c), C.c_size_t(c.shape[0]), keys, _p(o)))
        return o

    # ---- CPU baseline ----------------------------------------------------------------------
    def bench_decompose(self, enc_key, threads, blocks_per_thread, use_bsgs=False):
        s = lib().ref_bench_decompose(self.h, _p(np.ascontiguousarray(enc_key)), threads, blocks_per_thread, int(use_bsgs))
        if s < 0:
            raise RefError("reference decomposition failed")
        return s

    def bench_primitive(self, ct, op, reps):
        s = lib().ref_bench_primitive(self.h, _p(np.ascontiguousarray(ct)), op, reps)
        if s < 0:
            raise RefError(lib().ref_last_error().decode())
        return s


def pasta_plain(key256, p, data, decrypt=False):
    k = np.ascontiguousarray(key256, dtype=np.uint64)
    d = np.ascontiguousarray(data, dtype=np.uint64)
    o = np.zeros_like(d)
    rc = lib().ref_pasta_plain(_p(k), C.c_uint64(p), _p(d), C.c_size_t(len(d)), int(decrypt), _p(o))
    if rc:
        raise RefError(lib().ref_last_error().decode())
    return o


def pasta_layer_material(p, nonce, counter, layer):
    m1 = np.zeros((128, 128), dtype=np.uint64)
    m2 = np.zeros((128, 128), dtype=np.uint64)
    rc = np.zeros(256, dtype=np.uint64)
    r = lib().ref_pasta_layer_material(C.c_uint64(p), C.c_uint64(nonce), C.c_uint64(counter), layer, _p(m1), _p(m2), _p(rc))
    if r:
        raise RefError(lib().ref_last_error().decode())
    return m1, m2, rc
