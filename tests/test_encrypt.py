"""seal::Encryptor::encrypt on the engine: the Blake2xb generator stream, the samplers and the ciphertext
itself against the reference's Encryptor run with the same seed (oracle/_ref: libseal-4.0.a with a seeded Blake2xbPRNGFactory)."""
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest

import common
from oracle import refshim as R

pkg = common.package()
EMUL = os.environ.get("HHE_EMUL_LIB") or os.path.join(common.ROOT, "tests", "emul", "libhhe_emul.so")
BACKENDS = [pytest.param("emul"), pytest.param("cuda", marks=pytest.mark.gpu)]
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libhhe_ref.so not built")


def make_ctx(backend, N, q):
    if backend == "emul":
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(EMUL)])
        return pkg.Context(N, common.T, q, lib_path=EMUL, emulation_harness=True)
    return pkg.Context(N, common.T, q, device=0)


_IV = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
       0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]
_SIGMA = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
          [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
          [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
          [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
          [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0]]
_M = (1 << 64) - 1


def _compress(h, block, t, last):
    """BLAKE2b compression function F (RFC 7693 section 3.2), plain Python."""
    m = struct.unpack("<16Q", block)
    v = list(h) + _IV[:]
    v[12] ^= t
    if last:
        v[14] ^= _M
    rot = lambda x, n: ((x >> n) | (x << (64 - n))) & _M  # noqa: E731
    for r in range(12):
        s = _SIGMA[r % 10]
        for i, (a, b, c, d) in enumerate(((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
                                          (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14))):
            v[a] = (v[a] + v[b] + m[s[2 * i]]) & _M
            v[d] = rot(v[d] ^ v[a], 32)
            v[c] = (v[c] + v[d]) & _M
            v[b] = rot(v[b] ^ v[c], 24)
            v[a] = (v[a] + v[b] + m[s[2 * i + 1]]) & _M
            v[d] = rot(v[d] ^ v[a], 16)
            v[c] = (v[c] + v[d]) & _M
            v[b] = rot(v[b] ^ v[c], 63)
    return [h[i] ^ v[i] ^ v[8 + i] for i in range(8)]


def _init(digest, keylen, fanout, depth, leaf, node_offset, xof, node_depth, inner):
    p = struct.pack("<BBBBIIIBB", digest, keylen, fanout, depth, leaf, node_offset, xof, node_depth, inner) + bytes(46)
    return [_IV[i] ^ w for i, w in enumerate(struct.unpack("<8Q", p))]


def blake2xb_stream(seed8, refills):
    """SEAL's Blake2xbPRNG byte stream (seal/randomgen.h): refill k = blake2xb(4096 bytes, in = counter k, key = seed), restated from
    RFC 7693 + the BLAKE2X construction in plain Python (hashlib cannot express the depth-0 expansion nodes). The keyed root hash is
    cross-checked against hashlib, which can express it (the xof length sits in the upper half of hashlib's 64-bit node_offset)."""
    key = struct.pack("<8Q", *[int(v) for v in seed8])
    out = b""
    for counter in range(refills):
        h = _init(64, 64, 1, 1, 0, 0, 4096, 0, 0)
        h = _compress(h, key + bytes(64), 128, False)
        h = _compress(h, struct.pack("<Q", counter) + bytes(120), 136, True)
        root = struct.pack("<8Q", *h)
        assert root == hashlib.blake2b(struct.pack("<Q", counter), digest_size=64, key=key, fanout=1, depth=1, leaf_size=0,
                                       node_offset=4096 << 32, node_depth=0, inner_size=0).digest()
        for i in range(64):
            c = _compress(_init(64, 0, 0, 0, 64, i, 4096, 0, 64), root + bytes(64), 64, True)
            out += struct.pack("<8Q", *c)
    return out


def seal_encrypt_restated(ref, slots):
    """Independent restatement of the documented algorithm on the host (numpy + the reference's own NTT through refshim): used to
    pin the sampler semantics, not as a product path."""
    N, K, L, q, t = ref.N, ref.K, ref.L, [int(v) for v in ref.q], ref.t
    s = blake2xb_stream(ref.prng_seed(), (16 * N + 4095) // 4096 + 1)
    words = np.frombuffer(s, dtype="<u4")
    assert (words[:N] != 0).all()
    r = (words[:N].astype(np.uint64) * 3) >> 32  # 0,1,2 -> -1,0,1
    by = np.frombuffer(s, dtype=np.uint8)[4 * N:4 * N + 12 * N].reshape(2, N, 6).astype(np.int64)
    pc = np.array([bin(v).count("1") for v in range(256)], dtype=np.int64)
    noise = pc[by[..., 0]] + pc[by[..., 1]] + pc[by[..., 2] & 31] - pc[by[..., 3]] - pc[by[..., 4]] - pc[by[..., 5] & 31]
    pk = ref.public_key()
    c = np.zeros((2, K, N), dtype=object)
    for k in range(K):
        u = np.where(r == 0, q[k] - 1, r - 1).astype(np.uint64)
        un = ref.ntt(k, u).astype(object)
        for j in range(2):
            prod = np.array((un * pk[j, k].astype(object)) % q[k], dtype=np.uint64)
            c[j, k] = (ref.ntt(k, prod, inverse=True).astype(object) + noise[j].astype(object)) % q[k]
    qsp, half = q[-1], q[-1] >> 1
    inv = [pow(qsp, -1, q[i]) for i in range(L)]
    out = np.zeros((2, L, N), dtype=np.uint64)
    for j in range(2):
        last = (c[j, K - 1] + half) % qsp
        for i in range(L):
            out[j, i] = np.array(((c[j, i] - (last % q[i]) + half % q[i]) * inv[i]) % q[i], dtype=np.uint64)
    return ref.add_plain(out, ref.encode(slots))


@needs_ref
def test_generator_stream_and_samplers_match_seal():
    """The hashlib restatement of the generator + samplers reproduces the reference's Encryptor bit for bit (this pins the algorithm
    the kernels implement: stream layout, Lemire ternary draw, centred binomial bytes, draw order, rounding division)."""
    q = common.small_params(1024, 3, 48)
    ref = R.Ref(1024, common.T, q, seed=9, steps=(0,), default_gk=False)
    slots = np.arange(1, 41, dtype=np.uint64)
    assert np.array_equal(seal_encrypt_restated(ref, slots), ref.encrypt(slots))
    ref.close()


@needs_ref
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("ring", ["f64_48", "int50"])
def test_encrypt_bit_exact_with_seal(backend, ring):
    q = common.small_params(1024, 3, 48 if ring == "f64_48" else 50)
    ref = R.Ref(1024, common.T, q, seed=12, steps=(0,), default_gk=False)
    ctx = make_ctx(backend, 1024, q)
    rng = np.random.default_rng(2)
    slots = rng.integers(0, common.T, (3, 200), dtype=np.uint64)
    seeds = np.tile(ref.prng_seed(), (3, 1))
    got = ctx.encrypt(ref.public_key(), slots=slots, seeds=seeds)
    for i in range(3):
        assert np.array_equal(got[i], ref.encrypt(slots[i]))
    # plaintext-coefficient entry point, and a different seed gives a different (still valid) ciphertext
    pt = ref.encode(slots[0])
    assert np.array_equal(ctx.encrypt(ref.public_key(), plain=pt, seeds=seeds[:1])[0], got[0])
    other = ctx.encrypt(ref.public_key(), slots=slots[:1], seeds=seeds[:1] + np.uint64(1))[0]
    assert not np.array_equal(other, got[0])
    dec, budget = ref.decrypt(other)
    assert budget > 0 and np.array_equal(dec[:200], slots[0])
    fresh = ctx.encrypt(ref.public_key(), slots=slots[:2])  # seeds from the operating system
    assert not np.array_equal(fresh[0], got[0]) and np.array_equal(ref.decrypt(fresh[1])[0][:200], slots[1])
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.encrypt(ref.public_key(), slots=np.full((1, 4), common.T, dtype=np.uint64))
    ctx.close()
    ref.close()


@needs_ref
@pytest.mark.gpu
def test_encrypt_weights_and_key_at_full_size():
    """BASELINE ring (N = 16384, BFVDefault): a weight row (sealhelper::encrypt_weight_mat) and the packed symmetric key
    (pastahelper::encrypt_symmetric_key) encrypted on the GPU equal the reference's ciphertexts for the same seed."""
    ref = R.Ref(16384, common.T, None, seed=6, steps=(0,), default_gk=False)
    ctx = pkg.Context(16384, common.T, ref.q, device=0)
    rng = np.random.default_rng(4)
    w = np.mod(rng.integers(-128, 128, 784), common.T).astype(np.uint64)
    key = common.pack_key(rng.integers(0, common.T, 256, dtype=np.uint64), 16384)
    seeds = np.tile(ref.prng_seed(), (1, 1))
    assert np.array_equal(ctx.encrypt(ref.public_key(), slots=w[None], seeds=seeds)[0], ref.encrypt(w))
    assert np.array_equal(ctx.encrypt(ref.public_key(), slots=key[None], seeds=seeds)[0], ref.encrypt(key))
    batch = ctx.encrypt(ref.public_key(), slots=np.tile(w, (5, 1)))
    for i in (0, 4):
        dec, budget = ref.decrypt(batch[i])
        assert budget > 100 and np.array_equal(dec[:784], w)
    ctx.close()
    ref.close()
