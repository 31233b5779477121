"""hhe_b200: Python binding (ctypes) of libhhe_b200.so, the B200-native BFV engine for the reference's
PASTA-3 transciphering + encrypted-FC hot path.  See include/hhe_b200.h for the C ABI and the reference
interfaces each call replaces; `host.py` mirrors the reference's class/helper names on top of it.

The directory name is not a Python identifier; import it with
    importlib.import_module("privacy-preserving-ml-through-hhe_b200")

There is no CPU path: constructing a Context without the CUDA library or without a Blackwell GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HHE_B200_LIB", os.path.join(_HERE, "libhhe_b200.so"))  # override: kernel-variant experiments
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)

HHE_OK, HHE_ERR_INVALID, HHE_ERR_RUNTIME, HHE_ERR_NO_DEVICE, HHE_ERR_LOGIC = 0, -1, -2, -3, -4
KEYSET_0, KEYSET_1, RELIN = 0, 1, 2

# every symbol include/hhe_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "hhe_last_error", "hhe_version", "hhe_ctx_create", "hhe_ctx_destroy", "hhe_ctx_info", "hhe_ctx_stream",
    "hhe_set_batch", "hhe_galois_elt", "hhe_ctx_constants", "hhe_load_ksk", "hhe_has_ksk", "hhe_ntt", "hhe_encode",
    "hhe_add", "hhe_negate", "hhe_add_plain", "hhe_multiply_plain", "hhe_rotate_rows", "hhe_rotate_columns",
    "hhe_multiply", "hhe_square", "hhe_relinearize", "hhe_exponentiate3", "hhe_pasta3_decompose",
    "hhe_pasta3_decompose_records", "hhe_mask", "hhe_flatten", "hhe_vec_sum", "hhe_fc_rows", "hhe_csp_decompose",
    "hhe_csp_evaluate_model", "hhe_dev_alloc",
    "hhe_dev_free", "hhe_dev_upload", "hhe_dev_download", "hhe_sync", "hhe_dev_ntt", "hhe_dev_rotate_rows",
    "hhe_dev_relinearize", "hhe_dev_multiply", "hhe_dev_pasta3_decompose", "hhe_launch_count",
    "hhe_pasta_layer_material", "hhe_profile_enable", "hhe_profile_reset", "hhe_profile_report", "hhe_clear_keyset",
    "hhe_seal_parms_id", "hhe_seal_ct_save_bound", "hhe_seal_ct_save", "hhe_seal_ct_load", "hhe_seal_keys_unpack",
    "hhe_load_seal_keys", "hhe_pasta3_decompose_serialized", "hhe_pasta3_plain", "hhe_build_is_cuda", "hhe_encrypt", "hhe_encrypt_slots",
]


class HheError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"hhe_b200 status {status}: {msg}")
        self.status = status


class HheInvalidArgument(HheError, ValueError):
    """std::invalid_argument in the reference (missing Galois key, bad sizes, values >= t)"""


class HheLogicError(HheError):
    """std::logic_error in the reference (transparent ciphertext)"""


class HheNoDevice(HheError):
    """no usable sm_100a device: the engine has no CPU path"""


_libs = {}


def load_library(path=None, emulation_harness=False):
    """Load the shared library. Fails loudly if it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)
    or if it is not the CUDA build: the host emulation of the kernel bodies (tests/emul) is accepted only when the caller says
    it is the test harness (`emulation_harness=True`, which nothing in the package does)."""
    path = path or LIB_PATH
    if path in _libs:
        l = _libs[path]
        if not emulation_harness and not l.hhe_build_is_cuda():
            raise HheNoDevice(HHE_ERR_NO_DEVICE, f"{path} is a host-emulation test build, not the CUDA engine: there is no CPU path")
        return l
    if not os.path.exists(path):
        raise HheNoDevice(HHE_ERR_NO_DEVICE, f"{path} is missing: build the CUDA extension first; there is no CPU fallback")
    l = C.CDLL(path)
    l.hhe_version.restype = C.c_char_p
    if not hasattr(l, "hhe_build_is_cuda") or (not emulation_harness and not l.hhe_build_is_cuda()):
        raise HheNoDevice(HHE_ERR_NO_DEVICE, f"{path} is not the CUDA build of the engine ({l.hhe_version().decode()}): there is no CPU path")
    l.hhe_last_error.restype = C.c_char_p
    l.hhe_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_uint64, C.c_uint64, _u64p, C.c_int, C.c_int, C.c_void_p]
    l.hhe_ctx_destroy.argtypes = [C.c_void_p]
    l.hhe_ctx_destroy.restype = None
    l.hhe_ctx_stream.restype = C.c_void_p
    l.hhe_ctx_stream.argtypes = [C.c_void_p]
    l.hhe_galois_elt.restype = C.c_uint32
    l.hhe_galois_elt.argtypes = [C.c_void_p, C.c_int]
    l.hhe_launch_count.restype = C.c_uint64
    l.hhe_launch_count.argtypes = [C.c_void_p]
    l.hhe_has_ksk.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    l.hhe_clear_keyset.argtypes = [C.c_void_p, C.c_int]
    l.hhe_seal_ct_save_bound.restype = C.c_size_t
    l.hhe_load_seal_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    _libs[path] = l
    return l


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(_u64p)


class Context:
    """One engine context = one GPU, one stream, one BFV parameter set (the SEALContext + Evaluator of the reference)."""

    def __init__(self, N, t, q, device=0, stream=None, lib_path=None, emulation_harness=False):
        self.lib = load_library(lib_path, emulation_harness)
        self.q = np.ascontiguousarray(q, dtype=np.uint64)
        self.N, self.t, self.K, self.L = int(N), int(t), len(self.q), len(self.q) - 1
        h = C.c_void_p()
        rc = self.lib.hhe_ctx_create(C.byref(h), self.N, self.t, self.q.ctypes.data_as(_u64p), self.K, device,
                                     C.c_void_p(stream) if stream else None)
        self.h = h if rc == 0 else None
        self._chk(rc)
        self.ct_words = 2 * self.L * self.N

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _chk(self, rc):
        if rc == HHE_OK:
            return
        msg = self.lib.hhe_last_error().decode()
        cls = {HHE_ERR_INVALID: HheInvalidArgument, HHE_ERR_LOGIC: HheLogicError, HHE_ERR_NO_DEVICE: HheNoDevice}.get(rc, HheError)
        raise cls(rc, msg)

    def close(self):
        if self.h:
            self.lib.hhe_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        out = np.zeros(7, dtype=np.uint64)
        self._chk(self.lib.hhe_ctx_info(self.h, out.ctypes.data_as(_u64p)))
        return dict(N=int(out[0]), L=int(out[1]), K=int(out[2]), t=int(out[3]), batch=int(out[4]), sms=int(out[5]),
                    fp64_moduli=int(out[6]))

    def stream(self):
        return self.lib.hhe_ctx_stream(self.h)

    def set_batch(self, blocks):
        self._chk(self.lib.hhe_set_batch(self.h, int(blocks)))

    def launch_count(self):
        return int(self.lib.hhe_launch_count(self.h))

    def profile(self, on=True):
        self._chk(self.lib.hhe_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._chk(self.lib.hhe_profile_reset(self.h))

    def profile_report(self):
        import json
        buf = C.create_string_buffer(16384)
        self._chk(self.lib.hhe_profile_report(self.h, buf, C.c_size_t(len(buf))))
        return json.loads(buf.value.decode())

    def galois_elt(self, step):
        return int(self.lib.hhe_galois_elt(self.h, step))

    def constants(self):
        K, L = self.K, self.L
        out = np.zeros(K + 1 + 3 + 2 * L + 1, dtype=np.uint64)
        self._chk(self.lib.hhe_ctx_constants(self.h, out.ctypes.data_as(_u64p)))
        return dict(psi=out[:K].copy(), psi_t=int(out[K]), m_sk=int(out[K + 1]), gamma=int(out[K + 2]),
                    m_tilde=int(out[K + 3]), base_B=out[K + 4 : K + 4 + L].copy(), bsk_roots=out[K + 4 + L :].copy())

    def load_ksk(self, kind, elt, ksk):
        k, p = _arr(ksk)
        if k.size != self.L * 2 * self.K * self.N:
            raise HheInvalidArgument(HHE_ERR_INVALID, "key-switching key has the wrong size")
        self._chk(self.lib.hhe_load_ksk(self.h, kind, C.c_uint32(elt), p))

    def has_ksk(self, kind, elt):
        return bool(self.lib.hhe_has_ksk(self.h, kind, C.c_uint32(elt)))

    def pasta3_plain(self, key256, words, decrypt=False, nonce=123456789, first_counter=0):
        """pasta::PASTA::encrypt / decrypt (src/pasta/pasta_3_plain.cpp:9-47) on the GPU"""
        k, pk = _arr(key256)
        if k.size != 256:
            raise HheInvalidArgument(HHE_ERR_INVALID, "the PASTA-3 key has 256 words")
        w, pw = _arr(words)
        o = np.zeros(w.size, dtype=np.uint64)
        self._chk(self.lib.hhe_pasta3_plain(self.h, pk, pw, C.c_size_t(w.size), C.c_uint64(nonce), C.c_uint64(first_counter),
                                            int(decrypt), o.ctypes.data_as(_u64p)))
        return o

    def encrypt(self, pk, plain=None, slots=None, seeds=None):
        """seal::Encryptor::encrypt with the public key `pk` ([2][K][N], NTT form) on the GPU. Either `plain` ([count][N] plaintext
        coefficients) or `slots` ([count][n] batch slots, encoded first). `seeds` ([count][8] uint64): one SEAL prng_seed_type per
        ciphertext (None: from the operating system). Bit-identical to SEAL for the same seed."""
        k, pk_p = _arr(pk)
        if k.size != 2 * self.K * self.N:
            raise HheInvalidArgument(HHE_ERR_INVALID, "public key has the wrong size")
        src = np.ascontiguousarray(plain if slots is None else slots, dtype=np.uint64)
        src = src[None] if src.ndim == 1 else src
        count = src.shape[0]
        sp = None
        if seeds is not None:
            sd = np.ascontiguousarray(seeds, dtype=np.uint64).reshape(count, 8)
            sp = sd.ctypes.data_as(_u64p)
        o = np.zeros((count, 2, self.L, self.N), dtype=np.uint64)
        if slots is None:
            if src.shape[1] != self.N:
                raise HheInvalidArgument(HHE_ERR_INVALID, "plaintext must have N coefficients")
            rc = self.lib.hhe_encrypt(self.h, pk_p, sp, src.ctypes.data_as(_u64p), C.c_size_t(count), o.ctypes.data_as(_u64p))
        else:
            rc = self.lib.hhe_encrypt_slots(self.h, pk_p, sp, src.ctypes.data_as(_u64p), C.c_size_t(src.shape[1]), C.c_size_t(count),
                                            o.ctypes.data_as(_u64p))
        self._chk(rc)
        return o

    def clear_keyset(self, kind):
        self._chk(self.lib.hhe_clear_keyset(self.h, int(kind)))

    def load_seal_keys(self, kind, data):
        """GaloisKeys / RelinKeys::load from the serialized bytes (SEAL wire format), every key uploaded as it is parsed.
        Returns the number of keys loaded."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        n, used = C.c_size_t(0), C.c_size_t(0)
        self._chk(self.lib.hhe_load_seal_keys(self.h, int(kind), buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size),
                                              C.byref(n), C.byref(used)))
        return int(n.value)

    def pasta3_decompose_serialized(self, enc_key_bytes, sym_ct, use_bsgs=False, nonce=123456789, first_counter=0, compr_mode=2):
        """Serialized enc. key in, list of serialized result ciphertexts out (the CSP's gRPC request/response payloads)."""
        from . import seal_io
        kb = np.frombuffer(bytes(enc_key_bytes), dtype=np.uint8)
        s = np.ascontiguousarray(sym_ct, dtype=np.uint64)
        nblk = (s.size + 127) // 128
        ring = seal_io.Ring(self.N, self.t, self.q, lib=self.lib)
        cap = nblk * ring.ct_save_bound(2)
        out = np.zeros(cap, dtype=np.uint8)
        sizes = (C.c_size_t * max(1, nblk))()
        written = C.c_size_t(0)
        self._chk(self.lib.hhe_pasta3_decompose_serialized(
            self.h, kb.ctypes.data_as(C.c_void_p), C.c_size_t(kb.size), s.ctypes.data_as(_u64p), C.c_size_t(s.size), C.c_uint64(nonce),
            C.c_uint64(first_counter), int(use_bsgs), int(compr_mode), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), sizes,
            C.byref(written)))
        res, o = [], 0
        for b in range(nblk):
            res.append(out[o:o + sizes[b]].tobytes())
            o += sizes[b]
        return res

    def _cts(self, a, size=2):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        w = size * self.L * self.N
        if a.size % w:
            raise HheInvalidArgument(HHE_ERR_INVALID, "ciphertext array has the wrong size")
        return a, a.ctypes.data_as(_u64p), a.size // w

    def _out(self, count, size=2, like=None):
        shape = (count, size, self.L, self.N) if (like is None or like.ndim == 4) else (size, self.L, self.N)
        return np.zeros(shape, dtype=np.uint64)

    # -- primitives -------------------------------------------------------------------------------------------
    def ntt(self, limb, data, inverse=False):
        d = np.array(data, dtype=np.uint64, copy=True, order="C")
        self._chk(self.lib.hhe_ntt(self.h, limb, int(inverse), d.ctypes.data_as(_u64p), C.c_size_t(d.size // self.N)))
        return d

    def encode(self, slots):
        s = np.ascontiguousarray(slots, dtype=np.uint64)
        count = 1 if s.ndim == 1 else s.shape[0]
        n = s.shape[-1]
        out = np.zeros((count, self.N) if s.ndim == 2 else self.N, dtype=np.uint64)
        self._chk(self.lib.hhe_encode(self.h, s.ctypes.data_as(_u64p), C.c_size_t(n), out.ctypes.data_as(_u64p), C.c_size_t(count)))
        return out

    def add(self, a, b):
        a, pa, n = self._cts(a)
        b, pb, _ = self._cts(b)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_add(self.h, pa, pb, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def negate(self, a):
        a, pa, n = self._cts(a)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_negate(self.h, pa, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def add_plain(self, a, pt):
        a, pa, n = self._cts(a)
        p, pp = _arr(pt)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_add_plain(self.h, pa, pp, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def multiply_plain(self, a, pt):
        a, pa, n = self._cts(a)
        p, pp = _arr(pt)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_multiply_plain(self.h, pa, pp, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def rotate_rows(self, a, steps, keys=KEYSET_0):
        a, pa, n = self._cts(a)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_rotate_rows(self.h, pa, int(steps), keys, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def rotate_columns(self, a, keys=KEYSET_0):
        a, pa, n = self._cts(a)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_rotate_columns(self.h, pa, keys, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def multiply(self, a, b):
        a, pa, n = self._cts(a)
        b, pb, _ = self._cts(b)
        o = self._out(n, 3, like=a)
        self._chk(self.lib.hhe_multiply(self.h, pa, pb, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def square(self, a):
        a, pa, n = self._cts(a)
        o = self._out(n, 3, like=a)
        self._chk(self.lib.hhe_square(self.h, pa, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def relinearize(self, a3):
        a, pa, n = self._cts(a3, 3)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_relinearize(self.h, pa, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def exponentiate3(self, a):
        a, pa, n = self._cts(a)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_exponentiate3(self.h, pa, o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    # -- hot path ---------------------------------------------------------------------------------------------
    def _into(self, out, shape):
        """Caller-owned result buffer (e.g. pinned host memory, reused between calls) or a fresh array"""
        if out is None:
            return np.zeros(shape, dtype=np.uint64)
        if out.dtype != np.uint64 or not out.flags.c_contiguous or out.size != int(np.prod(shape)):
            raise HheInvalidArgument(HHE_ERR_INVALID, "`out` must be a C-contiguous uint64 array of the result's size")
        return out.reshape(shape)

    def pasta3_decompose(self, enc_key, sym_ct, use_bsgs=False, nonce=123456789, first_counter=0, records=1, out=None):
        k, pk, _ = self._cts(enc_key)
        s = np.ascontiguousarray(sym_ct, dtype=np.uint64)
        n_words = s.size // records
        nblk = (n_words + 127) // 128
        o = self._into(out, (records * nblk, 2, self.L, self.N))
        if records == 1:
            rc = self.lib.hhe_pasta3_decompose(self.h, pk, s.ctypes.data_as(_u64p), C.c_size_t(n_words), C.c_uint64(nonce),
                                               C.c_uint64(first_counter), int(use_bsgs), o.ctypes.data_as(_u64p))
        else:
            rc = self.lib.hhe_pasta3_decompose_records(self.h, pk, s.ctypes.data_as(_u64p), C.c_size_t(n_words),
                                                       C.c_size_t(records), C.c_uint64(nonce), C.c_uint64(first_counter),
                                                       int(use_bsgs), o.ctypes.data_as(_u64p))
        self._chk(rc)
        return o

    def mask(self, a, mask):
        a, pa, n = self._cts(a)
        m, pm = _arr(mask)
        o = self._out(n, like=a)
        self._chk(self.lib.hhe_mask(self.h, pa, pm, C.c_size_t(m.size), o.ctypes.data_as(_u64p), C.c_size_t(n)))
        return o

    def flatten(self, cts, keys=KEYSET_0, groups=1):
        a, pa, n = self._cts(cts)
        per = n // groups
        o = np.zeros((groups, 2, self.L, self.N) if groups > 1 else (2, self.L, self.N), dtype=np.uint64)
        self._chk(self.lib.hhe_flatten(self.h, pa, C.c_size_t(per), keys, o.ctypes.data_as(_u64p), C.c_size_t(groups)))
        return o

    def vec_sum(self, a, n, keys=KEYSET_1):
        a, pa, cnt = self._cts(a)
        o = self._out(cnt, like=a)
        self._chk(self.lib.hhe_vec_sum(self.h, pa, C.c_size_t(n), keys, o.ctypes.data_as(_u64p), C.c_size_t(cnt)))
        return o

    def fc_rows(self, x, w, n, keys=KEYSET_1, out=None):
        x, px, ns = self._cts(x)
        w, pw, nr = self._cts(w)
        o = self._into(out, (ns, nr, 2, self.L, self.N))
        self._chk(self.lib.hhe_fc_rows(self.h, px, C.c_size_t(ns), pw, C.c_size_t(nr), C.c_size_t(n), keys, o.ctypes.data_as(_u64p)))
        return o

    # -- service-level calls (BaseCSP::decompose / CSP_hhe_pktnn_1fc::evaluateModel, src/examples/CSP/CSP.cpp:235-323) --------
    def csp_decompose(self, enc_key, sym_ct, records=1, use_bsgs=False, apply_mask=False, flatten_keys=KEYSET_1, nonce=123456789):
        k, pk, _ = self._cts(enc_key)
        s = np.ascontiguousarray(sym_ct, dtype=np.uint64)
        n_words = s.size // records
        o = np.zeros((records, 2, self.L, self.N), dtype=np.uint64)
        self._chk(self.lib.hhe_csp_decompose(self.h, pk, s.ctypes.data_as(_u64p), C.c_size_t(n_words), C.c_size_t(records),
                                             C.c_uint64(nonce), int(use_bsgs), int(apply_mask), int(flatten_keys),
                                             o.ctypes.data_as(_u64p)))
        return o

    def csp_evaluate_model(self, records_ct, enc_weights, input_len, keys=KEYSET_1):
        x, px, ns = self._cts(records_ct)
        w, pw, nr = self._cts(enc_weights)
        o = np.zeros((ns, nr, 2, self.L, self.N), dtype=np.uint64)
        self._chk(self.lib.hhe_csp_evaluate_model(self.h, px, C.c_size_t(ns), pw, C.c_size_t(nr), C.c_size_t(input_len), keys,
                                                  o.ctypes.data_as(_u64p)))
        return o

    def pasta_layer_material(self, nonce, counter, layer):
        m1 = np.zeros((128, 128), dtype=np.uint32)
        m2 = np.zeros((128, 128), dtype=np.uint32)
        rc = np.zeros(256, dtype=np.uint32)
        self._chk(self.lib.hhe_pasta_layer_material(self.h, C.c_uint64(nonce), C.c_uint64(counter), layer,
                                                    m1.ctypes.data_as(_u32p), m2.ctypes.data_as(_u32p), rc.ctypes.data_as(_u32p)))
        return m1, m2, rc

    # -- device-resident API ----------------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._chk(self.lib.hhe_dev_alloc(self.h, C.c_size_t(nbytes), C.byref(p)))
        return p

    def dev_free(self, p):
        self._chk(self.lib.hhe_dev_free(self.h, p))

    def dev_upload(self, p, arr):
        a = np.ascontiguousarray(arr)
        self._chk(self.lib.hhe_dev_upload(self.h, p, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes)))

    def dev_download(self, p, arr):
        assert arr.flags["C_CONTIGUOUS"]
        self._chk(self.lib.hhe_dev_download(self.h, arr.ctypes.data_as(C.c_void_p), p, C.c_size_t(arr.nbytes)))

    def sync(self):
        self._chk(self.lib.hhe_sync(self.h))

    def dev_ntt(self, limb, inverse, dptr, count):
        self._chk(self.lib.hhe_dev_ntt(self.h, limb, int(inverse), C.cast(dptr, _u64p), C.c_size_t(count)))

    def dev_rotate_rows(self, d_a, steps, keys, d_out, count):
        self._chk(self.lib.hhe_dev_rotate_rows(self.h, C.cast(d_a, _u64p), steps, keys, C.cast(d_out, _u64p), C.c_size_t(count)))

    def dev_relinearize(self, d_a3, d_out, count):
        self._chk(self.lib.hhe_dev_relinearize(self.h, C.cast(d_a3, _u64p), C.cast(d_out, _u64p), C.c_size_t(count)))

    def dev_multiply(self, d_a, d_b, d_out3, count):
        self._chk(self.lib.hhe_dev_multiply(self.h, C.cast(d_a, _u64p), C.cast(d_b, _u64p), C.cast(d_out3, _u64p), C.c_size_t(count)))

    def dev_pasta3_decompose(self, d_key, d_sym, lens, counters, nonce, use_bsgs, d_out):
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        ctr = np.ascontiguousarray(counters, dtype=np.uint64)
        self._chk(self.lib.hhe_dev_pasta3_decompose(self.h, C.cast(d_key, _u64p), C.cast(d_sym, _u64p), lens.ctypes.data_as(_u32p),
                                                    ctr.ctypes.data_as(_u64p), C.c_size_t(len(ctr)), C.c_uint64(nonce), int(use_bsgs),
                                                    C.cast(d_out, _u64p)))
