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
