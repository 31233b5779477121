"""SEAL 4.0 wire format at the engine's boundary -- Python binding of the stateless codec
entry points of include/hhe_b200.h (hhe_seal_*). Host-side byte work: no device is needed for these.

The reference moves ciphertexts and keys as seal::Ciphertext / GaloisKeys / RelinKeys ::save streams
(src/examples/CSP/CSP.cpp:131,201; Analyst.cpp:258-341; User.cpp:83,168) and checkpoints decompositions as
`size_t count` + `count` saved ciphertexts (CSP.cpp:495-547, read back at :520-547 and :575-595).
"""
import ctypes as C
import struct

import numpy as np

from . import HHE_ERR_INVALID, HHE_ERR_LOGIC, HHE_OK, HheError, HheInvalidArgument, HheLogicError, load_library

COMPR_NONE, COMPR_ZLIB, COMPR_ZSTD = 0, 1, 2
_u64p = C.POINTER(C.c_uint64)


class _RingStruct(C.Structure):
    _fields_ = [("N", C.c_uint64), ("t", C.c_uint64), ("q", _u64p), ("nq", C.c_int)]


class Ring:
    """The encryption parameters a serialized object is validated against (the role of SEALContext in load(context, ...))."""

    def __init__(self, N, t, q, lib=None):
        self.lib = lib or load_library()
        self.q = np.ascontiguousarray(q, dtype=np.uint64)
        self.N, self.t, self.K, self.L = int(N), int(t), len(self.q), len(self.q) - 1
        self._s = _RingStruct(self.N, self.t, self.q.ctypes.data_as(_u64p), self.K)

    def _chk(self, rc):
        if rc == HHE_OK:
            return
        msg = self.lib.hhe_last_error().decode()
        raise {HHE_ERR_INVALID: HheInvalidArgument, HHE_ERR_LOGIC: HheLogicError}.get(rc, HheError)(rc, msg)

    def parms_id(self, level=0):
        """level 0: SEALContext::first_parms_id(); 1: key_parms_id()"""
        out = np.zeros(4, dtype=np.uint64)
        self._chk(self.lib.hhe_seal_parms_id(C.byref(self._s), int(level), out.ctypes.data_as(_u64p)))
        return out

    def ct_save_bound(self, size=2):
        return int(self.lib.hhe_seal_ct_save_bound(C.byref(self._s), int(size)))

    def save_ciphertext(self, ct, compr_mode=COMPR_ZSTD):
        """seal::Ciphertext::save -> bytes. ct: uint64 [size][L][N]."""
        a = np.ascontiguousarray(ct, dtype=np.uint64)
        if a.size % (self.L * self.N):
            raise HheInvalidArgument(HHE_ERR_INVALID, "ciphertext array has the wrong size")
        size = a.size // (self.L * self.N)
        cap = self.ct_save_bound(size)
        out = np.zeros(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        self._chk(self.lib.hhe_seal_ct_save(C.byref(self._s), a.ctypes.data_as(_u64p), size, int(compr_mode),
                                            out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n)))
        return out[: n.value].tobytes()

    def load_ciphertext(self, data, offset=0):
        """seal::Ciphertext::load(context, ...) -> (uint64 [size][L][N], bytes consumed)"""
        buf = np.frombuffer(data, dtype=np.uint8)[offset:]
        out = np.zeros((3, self.L, self.N), dtype=np.uint64)
        size, used = C.c_int(0), C.c_size_t(0)
        self._chk(self.lib.hhe_seal_ct_load(C.byref(self._s), buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size),
                                            out.ctypes.data_as(_u64p), C.c_size_t(out.size), C.byref(size), C.byref(used)))
        return out[: size.value].copy(), int(used.value)

    def unpack_keys(self, data):
        """GaloisKeys / RelinKeys::load -> {index: ksk [L][2][K][N]}; Galois element = 2*index + 1, relin key at index 0."""
        buf = np.frombuffer(data, dtype=np.uint8)
        n = C.c_size_t(0)
        p = buf.ctypes.data_as(C.c_void_p)
        self._chk(self.lib.hhe_seal_keys_unpack(C.byref(self._s), p, C.c_size_t(buf.size), None, None, C.c_size_t(0), C.byref(n), None))
        keys = np.zeros((max(1, n.value), self.L, 2, self.K, self.N), dtype=np.uint64)
        idx = np.zeros(max(1, n.value), dtype=np.uint64)
        self._chk(self.lib.hhe_seal_keys_unpack(C.byref(self._s), p, C.c_size_t(buf.size), idx.ctypes.data_as(_u64p),
                                                keys.ctypes.data_as(_u64p), C.c_size_t(n.value), C.byref(n), None))
        return {int(idx[i]): keys[i] for i in range(n.value)}

    # ---- the decomposition checkpoint file of BaseCSP::write/readHHEDecompositionDataToFile (CSP.cpp:495-547) ----
    def save_ciphertext_vector(self, cts, compr_mode=COMPR_ZSTD):
        return struct.pack("<Q", len(cts)) + b"".join(self.save_ciphertext(c, compr_mode) for c in cts)

    def load_ciphertext_vector(self, data):
        if len(data) < 8:
            raise HheLogicError(HHE_ERR_LOGIC, "Failed to read the size of the ciphertext array.")
        (count,) = struct.unpack_from("<Q", data, 0)
        out, off = [], 8
        for _ in range(count):
            ct, used = self.load_ciphertext(data, off)
            out.append(ct)
            off += used
        return out
