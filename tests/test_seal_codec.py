"""SEAL 4.0 wire format at the engine's boundary, pinned against SEAL itself.

The checker is the reference's own libseal (oracle/_ref via refshim: Ciphertext/GaloisKeys/RelinKeys save + load). The stateless
codec entry points (hhe_seal_*) are host-side byte work inside libhhe_b200.so and need no device, so this file runs in the
CPU tier; the two device-facing entry points are covered on the emulation harness here and on the B200 (`gpu`) below.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import common
from oracle import refshim as R

pkg = common.package()
seal_io = __import__("importlib").import_module(common.PKG + ".seal_io")
pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
EMUL = os.environ.get("HHE_EMUL_LIB") or os.path.join(common.ROOT, "tests", "emul", "libhhe_emul.so")
N = 1024
Q = common.small_params(N, 3, 48)
STEPS = (0, -1, 128)


@pytest.fixture(scope="module")
def ref():
    return R.Ref(N, common.T, Q, seed=11, steps=STEPS, default_gk=False)


@pytest.fixture(scope="module")
def ring():
    if not os.path.exists(pkg.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return seal_io.Ring(N, common.T, Q)


def test_parms_id_matches_seal(ref, ring):
    assert np.array_equal(ring.parms_id(0), ref.parms_id(0))
    assert np.array_equal(ring.parms_id(1), ref.parms_id(1))


def test_parms_id_bfv_default_16384(ring):
    r = R.Ref(16384, common.T, None, seed=1, steps=(), default_gk=False)
    big = seal_io.Ring(16384, common.T, common.Q_16384)
    assert np.array_equal(big.parms_id(0), r.parms_id(0)) and np.array_equal(big.parms_id(1), r.parms_id(1))
    r.close()


def test_full_size_ciphertexts_bfv_default_16384():
    """BASELINE.json's ring: 2 MiB ciphertexts (and the 3 MiB size-3 form) through every compression mode, both directions."""
    r = R.Ref(16384, common.T, None, seed=2, steps=(), default_gk=False)
    big = seal_io.Ring(16384, common.T, common.Q_16384)
    ct = r.encrypt(np.arange(784, dtype=np.uint64) % 256)
    assert big.save_ciphertext(ct, seal_io.COMPR_NONE) == r.ct_save(ct, 0)
    for compr in (0, 1, 2):
        theirs = r.ct_save(ct, compr)
        got, used = big.load_ciphertext(theirs)
        assert used == len(theirs) and np.array_equal(got, ct)
        back, used = r.ct_load(big.save_ciphertext(ct, compr))
        assert np.array_equal(back, ct)
    ct3 = r.multiply(ct, ct)
    got3, _ = big.load_ciphertext(r.ct_save(ct3, 2))
    assert got3.shape == (3, 8, 16384) and np.array_equal(got3, ct3)
    rk = big.unpack_keys(r.keys_save(2, 2))  # the 18 MiB relinearisation key as SEAL ships it (zstd)
    assert list(rk) == [0] and np.array_equal(rk[0], r.ksk(2))
    r.close()


def test_uncompressed_save_is_byte_identical_to_seal(ref, ring):
    ct = ref.encrypt(np.arange(300, dtype=np.uint64))
    assert ring.save_ciphertext(ct, seal_io.COMPR_NONE) == ref.ct_save(ct, 0)
    ct3 = ref.multiply(ct, ct)  # size-3 ciphertext (between packed_enc_multiply and relinearize)
    assert ring.save_ciphertext(ct3, seal_io.COMPR_NONE) == ref.ct_save(ct3, 0)


@pytest.mark.parametrize("compr", [0, 1, 2])
def test_load_what_seal_saved(ref, ring, compr):
    ct = ref.encrypt(np.arange(128, dtype=np.uint64) * 3)
    data = ref.ct_save(ct, compr)
    got, used = ring.load_ciphertext(data + b"trailing bytes are not consumed")
    assert used == len(data) and np.array_equal(got, ct)


@pytest.mark.parametrize("compr", [1, 2])
def test_seal_loads_what_we_saved(ref, ring, compr):
    ct = ref.encrypt(np.arange(77, dtype=np.uint64))
    data = ring.save_ciphertext(ct, compr)
    assert len(data) < ct.nbytes  # 48-bit residues in 64-bit words do compress
    got, used = ref.ct_load(data)
    assert used == len(data) and np.array_equal(got, ct)
    again, _ = ring.load_ciphertext(data)
    assert np.array_equal(again, ct)


@pytest.mark.parametrize("compr", [0, 2])
def test_key_streams_unpack_to_seal_keys(ref, ring, compr):
    keys = ring.unpack_keys(ref.keys_save(0, compr))
    elts = ref.list_galois(0)
    assert sorted(keys) == sorted((e - 1) // 2 for e in elts)
    for e in elts:
        assert np.array_equal(keys[(e - 1) // 2], ref.ksk(0, e))
    rk = ring.unpack_keys(ref.keys_save(2, compr))
    assert list(rk) == [0] and np.array_equal(rk[0], ref.ksk(2))


def test_checkpoint_vector_round_trip(ref, ring):
    cts = [ref.encrypt(np.full(5, i, dtype=np.uint64)) for i in range(3)]
    blob = ring.save_ciphertext_vector(cts)
    assert struct.unpack_from("<Q", blob)[0] == 3
    back = ring.load_ciphertext_vector(blob)
    assert len(back) == 3 and all(np.array_equal(a, b) for a, b in zip(back, cts))
    # the file the reference writes (CSP.cpp:495-517): size_t count + SEAL's own save of every ciphertext
    theirs = struct.pack("<Q", 3) + b"".join(ref.ct_save(c, 2) for c in cts)
    assert all(np.array_equal(a, b) for a, b in zip(ring.load_ciphertext_vector(theirs), cts))


def test_invalid_streams_are_rejected_like_seal(ref, ring):
    ct = ref.encrypt(np.arange(4, dtype=np.uint64))
    good = ref.ct_save(ct, 0)
    bad_magic = b"\x00\x00" + good[2:]
    with pytest.raises(pkg.HheLogicError):
        ring.load_ciphertext(bad_magic)
    with pytest.raises(R.RefError):
        ref.ct_load(bad_magic)
    with pytest.raises((pkg.HheLogicError, pkg.HheInvalidArgument)):
        ring.load_ciphertext(good[: len(good) // 2])  # header announces more bytes than present
    other = seal_io.Ring(N, common.T, common.small_params(N, 3, 50))  # same shape, other primes: parms_id differs
    with pytest.raises(pkg.HheLogicError):
        other.load_ciphertext(good)
    # a residue >= q_0: is_data_valid_for fails in SEAL, must fail here
    raw = bytearray(good)
    off = len(good) - ct.nbytes
    raw[off : off + 8] = struct.pack("<Q", int(Q[0]))
    with pytest.raises(pkg.HheLogicError):
        ring.load_ciphertext(bytes(raw))
    with pytest.raises(R.RefError):
        ref.ct_load(bytes(raw))
    # zstd stream cut short inside the frame
    z = ref.ct_save(ct, 2)
    cut = bytearray(z[: len(z) - 40])
    cut[8:16] = struct.pack("<Q", len(cut))
    with pytest.raises(pkg.HheLogicError):
        ring.load_ciphertext(bytes(cut))
    # output buffer too small / unsupported mode are argument errors
    with pytest.raises(pkg.HheInvalidArgument):
        ring.save_ciphertext(ct, 7)
    # a ciphertext stream handed to the key loader
    with pytest.raises(pkg.HheLogicError):
        ring.unpack_keys(good)


def _ctx(backend):
    if backend == "emul":
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(EMUL)])
        return pkg.Context(N, common.T, Q, lib_path=EMUL, emulation_harness=True)
    return pkg.Context(N, common.T, Q, device=0)


@pytest.mark.parametrize("backend", [pytest.param("emul"), pytest.param("cuda", marks=pytest.mark.gpu)])
def test_serialized_keys_and_ciphertexts_through_the_engine(ref, backend):
    """Analyst.cpp:273-318 -> CSP: serialized keys in, User.cpp:168 serialized key ciphertext in, serialized results out."""
    ctx = _ctx(backend)
    assert ctx.load_seal_keys(pkg.KEYSET_0, ref.keys_save(0, 2)) == len(STEPS)
    assert ctx.load_seal_keys(pkg.RELIN, ref.keys_save(2, 0)) == 1
    for s in STEPS:
        assert ctx.has_ksk(pkg.KEYSET_0, ref.galois_elt(s))
    ct = ref.encrypt(np.arange(600, dtype=np.uint64))
    assert np.array_equal(ctx.rotate_rows(ct, -1, pkg.KEYSET_0).reshape(2, -1), ref.rotate_rows(ct, -1, 0).reshape(2, -1))
    # a RelinKeys stream is not a GaloisKeys keyset and vice versa: kinds are kept apart
    ctx.clear_keyset(pkg.KEYSET_0)
    assert not ctx.has_ksk(pkg.KEYSET_0, ref.galois_elt(-1))
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.rotate_rows(ct, -1, pkg.KEYSET_0)
    ctx.load_seal_keys(pkg.KEYSET_0, ref.keys_save(0, 0))

    rng = np.random.default_rng(5)
    sym_key = rng.integers(0, common.T, 256, dtype=np.uint64)
    enc_key = ref.encrypt(common.pack_key(sym_key, N))
    sym = rng.integers(0, common.T, 128 + 40, dtype=np.uint64)  # two blocks, the second one ragged
    ring = seal_io.Ring(N, common.T, Q, lib=ctx.lib)
    blobs = ctx.pasta3_decompose_serialized(ref.ct_save(enc_key, 2), sym, compr_mode=2)
    want = ref.pasta_decompose(enc_key, sym)
    assert len(blobs) == 2
    for b, w in zip(blobs, want):
        got, used = ref.ct_load(b)  # SEAL itself accepts the engine's reply
        assert used == len(b) and np.array_equal(got, w)
        mine, _ = ring.load_ciphertext(b)
        assert np.array_equal(mine, w)
    # error classes survive the serialized wrapper: a word >= t is std::invalid_argument in the reference (BatchEncoder::encode),
    # a key ciphertext that is not valid for the parameters is SEAL's std::logic_error
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.pasta3_decompose_serialized(ref.ct_save(enc_key, 2), np.array([common.T], dtype=np.uint64))
    with pytest.raises(pkg.HheLogicError):
        ctx.pasta3_decompose_serialized(b"\x00" * 64, sym)
    ctx.close()
