"""Shared helpers for the test-suite (parameter sets, key loading)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "privacy-preserving-ml-through-hhe_b200"
T = 65537
NONCE = 123456789

# CoeffModulus::BFVDefault(16384) -- the BASELINE.json configuration
Q_16384 = [281474976546817, 281474976317441, 281474975662081, 562949952798721, 562949952700417, 562949952274433,
           562949951979521, 562949951881217, 562949951619073]
Q_8192 = [8796092858369, 8796092792833, 17592186028033, 17592185438209, 17592184717313]


def _is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def ntt_primes(N, bits, count):
    """`count` largest primes below 2^bits that are 1 mod 2N."""
    f = 2 * N
    v = ((1 << bits) - 1) // f * f + 1
    out = []
    while len(out) < count:
        if _is_prime(v):
            out.append(v)
        v -= f
    return out


def small_params(N=1024, data_primes=6, bits=50):
    """A test ring: `data_primes` primes of `bits` bits + one special prime one bit larger.
    bits=50 exercises the integer (Shoup) kernels; bits=48 (all moduli <= 2^49) the FP64-pipe kernels."""
    return ntt_primes(N, bits, data_primes) + ntt_primes(N, bits + 1, 1)


def package():
    return importlib.import_module(PKG)


def pack_key(key256, N):
    """pastahelper::encrypt_symmetric_key slot layout (src/util/pastahelper.cpp:355-377)."""
    kt = np.zeros(N // 2 + 128, dtype=np.uint64)
    kt[:128] = key256[:128]
    kt[N // 2:] = key256[128:]
    return kt


def load_keys_from_ref(dst, ref, keysets=(0, 1)):
    """Copy every key the reference generated into an engine/oracle context (`dst.load_ksk`)."""
    for kind in keysets:
        for elt in ref.list_galois(kind):
            dst.load_ksk(kind, elt, ref.ksk(kind, elt))
    dst.load_ksk(2, 0, ref.ksk(2))


# ---- a self-contained key generator for the test ring (keys are inputs; any valid RLWE key set will do) -------------
class ToyKeys:
    """Secret key + key-switching keys built with the oracle's NTT (symmetric-key RLWE, SEAL's key layout)."""

    def __init__(self, orc, seed):
        self.o, self.rng = orc, np.random.default_rng(seed)
        self.N, self.K, self.L, self.q = orc.N, orc.K, orc.L, [int(v) for v in orc.q]
        s = self.rng.integers(-1, 2, self.N)
        self.s_ntt = np.stack([orc.ntt(k, np.mod(s, self.q[k]).astype(np.uint64)) for k in range(self.K)])

    def _mul(self, a, b, k):
        return np.array([int(x) * int(y) % self.q[k] for x, y in zip(a, b)], dtype=np.uint64)

    def ksk(self, new_key_ntt):
        """key[J][c][k]: c0 = -(a s + e) + q_sp * new_key [only on limb J], c1 = a  (all NTT form)"""
        out = np.zeros((self.L, 2, self.K, self.N), dtype=np.uint64)
        qsp = self.q[-1]
        for J in range(self.L):
            e = self.rng.integers(-3, 4, self.N)
            for k in range(self.K):
                qk = self.q[k]
                a = self.rng.integers(0, qk, self.N, dtype=np.uint64)
                e_ntt = self.o.ntt(k, np.mod(e, qk).astype(np.uint64))
                c0 = (qk - (self._mul(a, self.s_ntt[k], k).astype(object) + e_ntt.astype(object)) % qk) % qk
                if k == J:
                    c0 = (c0 + (qsp % qk) * new_key_ntt[k].astype(object)) % qk
                out[J, 0, k] = np.array(c0, dtype=np.uint64)
                out[J, 1, k] = a
        return out

    def galois_key(self, elt):
        # s(X^elt) in NTT form: permute coefficient form then transform
        s_coeff = [self.o.ntt(k, self.s_ntt[k], inverse=True) for k in range(self.K)]
        g = np.zeros((self.K, self.N), dtype=np.uint64)
        for k in range(self.K):
            idx = (np.arange(self.N, dtype=np.int64) * elt) % (2 * self.N)
            val = s_coeff[k].copy()
            neg = idx >= self.N
            val[neg] = (self.q[k] - val[neg]) % self.q[k]
            tmp = np.zeros(self.N, dtype=np.uint64)
            tmp[idx % self.N] = val
            g[k] = self.o.ntt(k, tmp)
        return self.ksk(g)

    def relin_key(self):
        s2 = np.stack([self._mul(self.s_ntt[k], self.s_ntt[k], k) for k in range(self.K)])
        return self.ksk(s2)

    def encrypt_zero_plus(self, orc, pt):
        """(c0, c1) = (-(a s + e) , a) + Delta*pt on c0 -- a valid fresh BFV ciphertext of pt."""
        ct = np.zeros((2, self.L, self.N), dtype=np.uint64)
        e = self.rng.integers(-3, 4, self.N)
        for i in range(self.L):
            qi = self.q[i]
            a = self.rng.integers(0, qi, self.N, dtype=np.uint64)
            as_ = orc.ntt(i, self._mul(orc.ntt(i, a), self.s_ntt[i], i), inverse=True)
            ct[0, i] = np.array((qi - (as_.astype(object) + np.mod(e, qi).astype(object)) % qi) % qi, dtype=np.uint64)
            ct[1, i] = a
        return orc.add_plain(ct, pt)
