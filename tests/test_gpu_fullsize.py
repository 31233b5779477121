"""Full-size (BASELINE.json: N=16384, t=65537, BFVDefault) parity on the B200 against the reference itself
(oracle/_ref/libhhe_ref.so: unmodified reference sources + vendored libseal, prebuilt in the container)."""
import numpy as np
import pytest

import common
from oracle import oracle as O
from oracle import refshim as R

pkg = common.package()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libhhe_ref.so not built")]
N = 16384


@pytest.fixture(scope="module")
def world():
    ref = R.Ref(N, common.T, None, seed=3, steps=(0, -1, 128), default_gk=False)
    assert list(ref.q) == common.Q_16384
    ctx = pkg.Context(N, common.T, ref.q, device=0)
    common.load_keys_from_ref(ctx, ref, keysets=(0,))
    rng = np.random.default_rng(5)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    yield dict(ref=ref, ctx=ctx, rng=rng, key=key, enc_key=ref.encrypt(common.pack_key(key, N)))
    ctx.close()
    ref.close()


def test_constants_equal_seal(world):
    ref, ctx = world["ref"], world["ctx"]
    c, rb = ctx.constants(), ref.behz()
    assert np.array_equal(c["psi"], ref.ntt_roots()[0]) and c["psi_t"] == ref.ntt_roots()[1]
    assert c["m_sk"] == rb["m_sk"] and np.array_equal(c["base_B"], rb["base_B"]) and np.array_equal(c["bsk_roots"], rb["bsk_roots"])


def test_ntt_all_limbs(world):
    ref, ctx, rng = world["ref"], world["ctx"], world["rng"]
    for limb in range(ref.K):
        x = rng.integers(0, int(ref.q[limb]), N, dtype=np.uint64)
        f = ctx.ntt(limb, x)
        assert np.array_equal(f, ref.ntt(limb, x))
        assert np.array_equal(ctx.ntt(limb, f, inverse=True), x)


def test_primitives(world):
    ref, ctx, rng = world["ref"], world["ctx"], world["rng"]
    a = ref.encrypt(rng.integers(0, common.T, N, dtype=np.uint64))
    b = ref.encrypt(rng.integers(0, common.T, N, dtype=np.uint64))
    pt = ref.encode(rng.integers(0, common.T, 9000, dtype=np.uint64))
    assert np.array_equal(ctx.add_plain(a, pt), ref.add_plain(a, pt))
    assert np.array_equal(ctx.multiply_plain(a, pt), ref.multiply_plain(a, pt))
    assert np.array_equal(ctx.rotate_rows(a, -1), ref.rotate_rows(a, -1))
    assert np.array_equal(ctx.rotate_rows(a, 128), ref.rotate_rows(a, 128))
    assert np.array_equal(ctx.rotate_columns(a), ref.rotate_columns(a))
    m3 = ref.multiply(a, b)
    assert np.array_equal(ctx.multiply(a, b), m3)
    assert np.array_equal(ctx.relinearize(m3), ref.relinearize(m3))
    assert np.array_equal(ctx.exponentiate3(a), ref.exponentiate3(a))


def test_one_block_bit_exact_with_seal(world):
    """BASELINE.json configs[0]: HE_decrypt of one 128-element block at N=16384."""
    ref, ctx, rng, key = world["ref"], world["ctx"], world["rng"], world["key"]
    pt = rng.integers(0, common.T, 128, dtype=np.uint64)
    sym = O.pasta_plain(key, common.T, pt)
    want = ref.pasta_decompose(world["enc_key"], sym, use_bsgs=False)
    got = ctx.pasta3_decompose(world["enc_key"], sym, use_bsgs=False)
    assert np.array_equal(got, want)
    slots, budget = ref.decrypt(got[0])
    assert budget > 60 and np.array_equal(slots[:128], pt)
    assert not slots[128:8192].any() and not slots[8320:].any()  # SURVEY B.3 slot layout


def test_batch_and_counter_invariance(world):
    """Size-independent property: a block's ciphertext depends only on (counter, words), not on how it was batched."""
    ref, ctx, rng, key = world["ref"], world["ctx"], world["rng"], world["key"]
    pt = rng.integers(0, common.T, 5 * 128 - 30, dtype=np.uint64)
    sym = O.pasta_plain(key, common.T, pt)
    ctx.set_batch(0)
    full = ctx.pasta3_decompose(world["enc_key"], sym)
    ctx.set_batch(2)
    split = ctx.pasta3_decompose(world["enc_key"], sym)
    ctx.set_batch(0)
    assert np.array_equal(full, split)
    one = ctx.pasta3_decompose(world["enc_key"], sym[3 * 128:4 * 128], first_counter=3)
    assert np.array_equal(one[0], full[3])
    for b in (0, 4):
        slots, _ = ref.decrypt(full[b])
        n = min(128, len(pt) - b * 128)
        assert np.array_equal(slots[:n], pt[b * 128:b * 128 + n])


def test_records_sharing_counter_zero(world):
    """BASELINE.json configs[2] shape (ECG: one-block records, every record restarts at counter 0): the batch shares its round
    material and diagonals; each record's ciphertext equals the one from a call of its own, and block 0 is SEAL's."""
    ref, ctx, rng, key = world["ref"], world["ctx"], world["rng"], world["key"]
    pts = rng.integers(0, 256, (5, 128), dtype=np.uint64)
    syms = np.stack([O.pasta_plain(key, common.T, p) for p in pts])
    got = ctx.pasta3_decompose(world["enc_key"], syms.reshape(-1), records=5)
    assert np.array_equal(got[0], ref.pasta_decompose(world["enc_key"], syms[0])[0])
    for r in (1, 4):
        assert np.array_equal(got[r], ctx.pasta3_decompose(world["enc_key"], syms[r])[0])
        assert np.array_equal(ref.decrypt(got[r])[0][:128], pts[r])


def test_siesta_shaped_records_regrouped_by_counter(world):
    """SURVEY.md 8d config 4, real-data variant: records of 300 words = 3 blocks each, counters 0..2 restarting per record
    (src/examples/CSP/CSP.cpp:247-252). The engine regroups the 12 blocks into 3 shared-counter batches; every ciphertext must
    equal the one a single-record call produces, sit at its record-major position, and decrypt (SEAL) to the record."""
    ref, ctx, rng, key = world["ref"], world["ctx"], world["rng"], world["key"]
    recs = rng.integers(0, 32, (4, 300), dtype=np.uint64)
    syms = np.stack([O.pasta_plain(key, common.T, r) for r in recs])
    got = ctx.pasta3_decompose(world["enc_key"], syms.reshape(-1), records=4).reshape(4, 3, 2, ctx.L, N)
    alone = ctx.pasta3_decompose(world["enc_key"], syms[2])
    assert np.array_equal(got[2], alone)
    assert np.array_equal(got[0, 1], ref.pasta_decompose(world["enc_key"], syms[0][:256])[1])  # SEAL's own block 1
    for r in range(4):
        for b in range(3):
            n = min(128, 300 - 128 * b)
            assert np.array_equal(ref.decrypt(got[r, b])[0][:n], recs[r, 128 * b:128 * b + n]), (r, b)


def test_full_bench_batch_every_block_decrypts(world):
    """The bench's batch (296 distinct-counter blocks in lock-step): every output ciphertext decrypts under SEAL to its PASTA
    plaintext with a healthy noise budget, and the batch equals two half-size calls (checksum of checksums)."""
    ref, ctx, rng, key = world["ref"], world["ctx"], world["rng"], world["key"]
    B = 296
    pt = rng.integers(0, common.T, B * 128, dtype=np.uint64)
    sym = O.pasta_plain(key, common.T, pt)
    ctx.set_batch(B)
    got = ctx.pasta3_decompose(world["enc_key"], sym)
    for b in range(B):
        slots, budget = ref.decrypt(got[b])
        assert budget > 60 and np.array_equal(slots[:128], pt[128 * b:128 * (b + 1)]), b
    ctx.set_batch(B // 2)
    halves = ctx.pasta3_decompose(world["enc_key"], sym)
    ctx.set_batch(0)
    dig = lambda a: a.reshape(a.shape[0], -1).sum(axis=1, dtype=np.uint64)  # noqa: E731
    assert np.array_equal(dig(got), dig(halves)) and int(dig(got).sum(dtype=np.uint64)) == int(dig(halves).sum(dtype=np.uint64))


def test_n32768_primitives_bit_exact_with_seal():
    """BASELINE.json configs[4] (primitive sweep) at N=32768, BFVDefault (L=15 + special): NTT, rotate, relinearize, multiply."""
    NN = 32768
    ref = R.Ref(NN, common.T, None, seed=6, steps=(-1,), default_gk=False)
    ctx = pkg.Context(NN, common.T, ref.q, device=0)
    assert ctx.info()["L"] == 15 and ctx.info()["fp64_moduli"] == 0
    c, rb = ctx.constants(), ref.behz()
    assert np.array_equal(c["psi"], ref.ntt_roots()[0]) and np.array_equal(c["base_B"], rb["base_B"])
    common.load_keys_from_ref(ctx, ref, keysets=(0,))
    rng = np.random.default_rng(12)
    for limb in (0, 7, ref.K - 1):
        x = rng.integers(0, int(ref.q[limb]), NN, dtype=np.uint64)
        f = ctx.ntt(limb, x)
        assert np.array_equal(f, ref.ntt(limb, x))
        assert np.array_equal(ctx.ntt(limb, f, inverse=True), x)
    a = ref.encrypt(rng.integers(0, common.T, NN, dtype=np.uint64))
    b = ref.encrypt(rng.integers(0, common.T, NN, dtype=np.uint64))
    assert np.array_equal(ctx.rotate_rows(a, -1), ref.rotate_rows(a, -1))
    m3 = ref.multiply(a, b)
    assert np.array_equal(ctx.multiply(a, b), m3)
    assert np.array_equal(ctx.relinearize(m3), ref.relinearize(m3))
    ctx.close()
    ref.close()


def test_n32768_whole_path_bit_exact_with_seal():
    """N = 32768 beyond the primitive sweep: encode, multiply_plain, mask and one whole PASTA-3 block (split transforms + element-wise
    pieces, integer arithmetic: the 55/56-bit primes are outside the FP64 path) against the reference (~90 s of SEAL for the block)."""
    NN = 32768
    ref = R.Ref(NN, common.T, None, seed=16, steps=(0, -1, 128), default_gk=False)
    ctx = pkg.Context(NN, common.T, ref.q, device=0)
    common.load_keys_from_ref(ctx, ref, keysets=(0,))
    rng = np.random.default_rng(19)
    sl = rng.integers(0, common.T, 9000, dtype=np.uint64)
    assert np.array_equal(ctx.encode(sl), ref.encode(sl))
    a = ref.encrypt(rng.integers(0, common.T, NN, dtype=np.uint64))
    pt = ref.encode(sl)
    assert np.array_equal(ctx.multiply_plain(a, pt), ref.multiply_plain(a, pt))
    assert np.array_equal(ctx.mask(a, np.ones(40, dtype=np.uint64)), ref.mask(a, np.ones(40, dtype=np.uint64)))
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    enc_key = ref.encrypt(common.pack_key(key, NN))
    words = rng.integers(0, common.T, 128, dtype=np.uint64)
    sym = O.pasta_plain(key, common.T, words)
    got = ctx.pasta3_decompose(enc_key, sym)
    assert np.array_equal(got, ref.pasta_decompose(enc_key, sym))
    slots, budget = ref.decrypt(got[0])
    assert budget > 100 and np.array_equal(slots[:128], words)
    ctx.close()
    ref.close()


def test_n8192_primitives_bit_exact_with_seal():
    """BASELINE.json configs[4] at N=8192, BFVDefault (L=4 + special, 43/44-bit primes: FP64-pipe kernels)."""
    NN = 8192
    ref = R.Ref(NN, common.T, None, seed=8, steps=(0, -1, 128), default_gk=False)
    assert list(ref.q) == common.Q_8192
    ctx = pkg.Context(NN, common.T, ref.q, device=0)
    assert ctx.info()["fp64_moduli"] == 5
    common.load_keys_from_ref(ctx, ref, keysets=(0,))
    rng = np.random.default_rng(13)
    for limb in range(ref.K):
        x = rng.integers(0, int(ref.q[limb]), NN, dtype=np.uint64)
        f = ctx.ntt(limb, x)
        assert np.array_equal(f, ref.ntt(limb, x)) and np.array_equal(ctx.ntt(limb, f, inverse=True), x)
    a = ref.encrypt(rng.integers(0, common.T, NN, dtype=np.uint64))
    b = ref.encrypt(rng.integers(0, common.T, NN, dtype=np.uint64))
    pt = ref.encode(rng.integers(0, common.T, 5000, dtype=np.uint64))
    assert np.array_equal(ctx.multiply_plain(a, pt), ref.multiply_plain(a, pt))
    assert np.array_equal(ctx.rotate_rows(a, -1), ref.rotate_rows(a, -1))
    assert np.array_equal(ctx.rotate_columns(a), ref.rotate_columns(a))
    m3 = ref.multiply(a, b)
    assert np.array_equal(ctx.multiply(a, b), m3)
    assert np.array_equal(ctx.relinearize(m3), ref.relinearize(m3))
    # the whole PASTA pipeline runs at this ring too (the noise budget is exhausted, limbs must still match)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    ek = ref.encrypt(common.pack_key(key, NN))
    sym = rng.integers(0, common.T, 128, dtype=np.uint64)
    assert np.array_equal(ctx.pasta3_decompose(ek, sym), ref.pasta_decompose(ek, sym, False))
    ctx.close()
    ref.close()


@pytest.fixture(scope="module")
def world_bsgs():
    """Keys for the baby-step/giant-step affine layer: rotations by -16 k, k = 1..7 (pasta_3_seal.cpp:190-201)."""
    steps = (0, -1, 128) + tuple(-16 * k for k in range(1, 8))
    ref = R.Ref(N, common.T, None, seed=13, steps=steps, default_gk=False)
    ctx = pkg.Context(N, common.T, ref.q, device=0)
    common.load_keys_from_ref(ctx, ref, keysets=(0,))
    rng = np.random.default_rng(17)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    yield dict(ref=ref, ctx=ctx, rng=rng, key=key, enc_key=ref.encrypt(common.pack_key(key, N)))
    ctx.close()
    ref.close()


def test_bsgs_block_bit_exact_with_seal(world_bsgs):
    """PASTA_SEAL::babystep_giantstep (pasta_3_seal.cpp:267-366, N1 = 16, N2 = 8) at the real size: the FP64 inner-sum kernel,
    the pre-rotated diagonals and the -16 k keys, every limb against the reference run with use_bsgs = true (13 s of SEAL)."""
    ref, ctx, rng, key = world_bsgs["ref"], world_bsgs["ctx"], world_bsgs["rng"], world_bsgs["key"]
    pt = rng.integers(0, common.T, 128, dtype=np.uint64)
    sym = O.pasta_plain(key, common.T, pt)
    want = ref.pasta_decompose(world_bsgs["enc_key"], sym, use_bsgs=True)
    got = ctx.pasta3_decompose(world_bsgs["enc_key"], sym, use_bsgs=True)
    assert np.array_equal(got, want)
    slots, budget = ref.decrypt(got[0])
    assert budget > 60 and np.array_equal(slots[:128], pt)
    # the two affine-layer modes are different op sequences: same plaintext, different ciphertexts
    diag = ctx.pasta3_decompose(world_bsgs["enc_key"], sym, use_bsgs=False)
    assert not np.array_equal(diag, got) and np.array_equal(ref.decrypt(diag[0])[0][:128], pt)


def test_bsgs_batch_matches_single_calls(world_bsgs):
    """BSGS mode, batched: distinct counters and records sharing counter 0 give the ciphertexts of one-block calls."""
    ref, ctx, rng, key = world_bsgs["ref"], world_bsgs["ctx"], world_bsgs["rng"], world_bsgs["key"]
    pt = rng.integers(0, common.T, 3 * 128, dtype=np.uint64)
    sym = O.pasta_plain(key, common.T, pt)
    full = ctx.pasta3_decompose(world_bsgs["enc_key"], sym, use_bsgs=True)
    one = ctx.pasta3_decompose(world_bsgs["enc_key"], sym[256:384], use_bsgs=True, first_counter=2)
    assert np.array_equal(one[0], full[2])
    assert np.array_equal(ref.decrypt(full[1])[0][:128], pt[128:256])
    recs = ctx.pasta3_decompose(world_bsgs["enc_key"], np.concatenate([sym[:128], sym[:128]]), use_bsgs=True, records=2)
    assert np.array_equal(recs[0], full[0]) and np.array_equal(recs[1], full[0])
