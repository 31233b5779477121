"""Parity of the engine with the oracle and the reference's golden vectors, through the C ABI.

backend "cuda" = the product (libhhe_b200.so on a B200; marked gpu).
backend "emul" = the SAME kernel bodies and host orchestration compiled with -DHHE_EMULATE (tests/emul, a test harness
that walks each CUDA grid on the host). It exists so that host logic and kernel index arithmetic are checked in the
CPU-only test tier; it is not part of the package and is never used as a fallback.
"""
import os
import subprocess

import numpy as np
import pytest

import common
from oracle import oracle as O

pkg = common.package()
FX = np.load(os.path.join(common.ROOT, "tests", "golden", "pasta_n512.npz"))
EMUL = os.environ.get("HHE_EMUL_LIB") or os.path.join(common.ROOT, "tests", "emul", "libhhe_emul.so")  # tools/asan_emul.sh: the ASan build
BACKENDS = [pytest.param("emul"), pytest.param("cuda", marks=pytest.mark.gpu)]
N = 1024
BSGS_STEPS = (-16, -32, -48, -64, -80, -96, -112)


def make_ctx(backend, NN, q):
    if backend == "emul":
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(EMUL)])
        return pkg.Context(NN, common.T, q, lib_path=EMUL, emulation_harness=True)
    return pkg.Context(NN, common.T, q, device=0)


ToyKeys = common.ToyKeys  # self-contained key generator for the test rings (tests/common.py)


# two rings: 50/51-bit primes -> integer Shoup kernels; 48/49-bit primes -> FP64-pipe kernels (the BFVDefault(16384) case)
@pytest.fixture(scope="module", params=["int50", "f64_48"])
def world(request):
    q = common.small_params(N, 6, 50) if request.param == "int50" else common.small_params(N, 7, 48)
    orc = O.Oracle(N, common.T, q)
    keys = ToyKeys(orc, 42)
    steps0 = (0, -1, 128) + BSGS_STEPS
    loaded = []
    for s in steps0:
        elt = orc.galois_elt(s)
        k = keys.galois_key(elt)
        orc.load_ksk(0, elt, k)
        loaded.append((0, elt, k))
    for s in [1, -1, 2, -2, 4, -4, 8, -8, 16, -16, 32, -32]:  # a power-of-two "default" set for NAF chains
        elt = orc.galois_elt(s)
        k = keys.galois_key(elt)
        orc.load_ksk(1, elt, k)
        loaded.append((1, elt, k))
    rk = keys.relin_key()
    orc.load_ksk(2, 0, rk)
    loaded.append((2, 0, rk))
    rng = np.random.default_rng(7)
    cts = [keys.encrypt_zero_plus(orc, orc.encode(rng.integers(0, common.T, N, dtype=np.uint64))) for _ in range(3)]
    return dict(q=q, orc=orc, keys=keys, loaded=loaded, cts=np.stack(cts), rng=rng)


@pytest.fixture(scope="module", params=BACKENDS)
def eng(request, world):
    ctx = make_ctx(request.param, N, world["q"])
    for kind, elt, k in world["loaded"]:
        ctx.load_ksk(kind, elt, k)
    yield ctx
    ctx.close()


def test_constants(eng, world):
    c, o = eng.constants(), world["orc"]
    assert np.array_equal(c["psi"], o.ntt_roots()[0]) and c["psi_t"] == o.ntt_roots()[1]
    b = o.behz()
    assert c["m_sk"] == b["m_sk"] and c["gamma"] == b["gamma"] and np.array_equal(c["base_B"], b["base_B"])
    assert np.array_equal(c["bsk_roots"], b["bsk_roots"])
    for s in (0, 1, -1, 128, -300, 511):
        assert eng.galois_elt(s) == o.galois_elt(s)


def test_ntt_every_modulus(eng, world):
    o, rng = world["orc"], np.random.default_rng(0)
    for limb in range(2 * o.K):
        mod = int(o.q[limb]) if limb < o.K else int(o.behz()["base_B"][limb - o.K]) if limb - o.K < o.L else o.behz()["m_sk"]
        x = rng.integers(0, mod, (2, N), dtype=np.uint64)
        x[0, :4] = [0, mod - 1, 1, mod - 1]  # edge residues
        f = eng.ntt(limb, x)
        assert np.array_equal(f, np.stack([o.ntt(limb, v) for v in x])), limb
        assert np.array_equal(eng.ntt(limb, f, inverse=True), x), limb


def test_encode(eng, world):
    o, rng = world["orc"], np.random.default_rng(1)
    sl = rng.integers(0, common.T, N, dtype=np.uint64)
    for n in (N, 300, 129, 1):
        assert np.array_equal(eng.encode(sl[:n]), o.encode(sl[:n]))
    sparse = np.zeros(N // 2 + 128, dtype=np.uint64)
    sparse[:128] = sl[:128]
    sparse[N // 2:] = sl[128:256]
    assert np.array_equal(eng.encode(sparse), o.encode(sparse))
    with pytest.raises(pkg.HheInvalidArgument):
        eng.encode(np.array([common.T], dtype=np.uint64))


def test_elementwise(eng, world):
    o, cts = world["orc"], world["cts"]
    pt = o.encode(np.random.default_rng(2).integers(0, common.T, 700, dtype=np.uint64))
    assert np.array_equal(eng.add(cts[0], cts[1]), o.add(cts[0], cts[1]))
    assert np.array_equal(eng.negate(cts[0]), o.negate(cts[0]))
    z = np.zeros_like(cts[0])
    assert np.array_equal(eng.negate(z), z)
    assert np.array_equal(eng.add_plain(cts[0], pt), o.add_plain(cts[0], pt))
    assert np.array_equal(eng.multiply_plain(cts[0], pt), o.multiply_plain(cts[0], pt))
    both = eng.add(cts[:2], cts[1:3])
    assert np.array_equal(both[0], o.add(cts[0], cts[1])) and np.array_equal(both[1], o.add(cts[1], cts[2]))
    with pytest.raises(pkg.HheLogicError):
        eng.multiply_plain(cts[0], np.zeros(N, dtype=np.uint64))
    # SEAL's monomial branch (one nonzero coefficient: multiplied as it is, no centred lift; oracle pinned against the reference in
    # test_oracle_vs_ref.py), mixed with a general plaintext in one batched call
    pts = np.zeros((3, N), dtype=np.uint64)
    pts[0, 0] = common.T - 1
    pts[1, 9] = 5
    pts[2] = pt
    got = eng.multiply_plain(cts, pts)
    for i in range(3):
        assert np.array_equal(got[i], o.multiply_plain(cts[i], pts[i])), i


def test_rotations(eng, world):
    o, cts = world["orc"], world["cts"]
    for s, ks in ((-1, 0), (128, 0), (-16, 0), (-112, 0), (1, 1), (-5, 1), (7, 1), (-27, 1), (31, 1)):
        assert np.array_equal(eng.rotate_rows(cts[0], s, ks), o.rotate_rows(cts[0], s, ks)), s
    assert np.array_equal(eng.rotate_rows(cts[0], 0, 0), cts[0])
    assert np.array_equal(eng.rotate_columns(cts[1]), o.rotate_columns(cts[1]))
    got = eng.rotate_rows(cts, -1, 0)
    assert all(np.array_equal(got[i], o.rotate_rows(cts[i], -1, 0)) for i in range(3))
    with pytest.raises(pkg.HheInvalidArgument):
        eng.rotate_rows(cts[0], -64, 1)  # single-term NAF without a key: "Galois key not present"
    with pytest.raises(pkg.HheInvalidArgument):
        eng.rotate_rows(cts[0], N // 2, 0)  # step count too large


def test_multiply_relinearize(eng, world):
    o, cts = world["orc"], world["cts"]
    m3 = o.multiply(cts[0], cts[1])
    assert np.array_equal(eng.multiply(cts[0], cts[1]), m3)
    assert np.array_equal(eng.square(cts[2]), o.multiply(cts[2], cts[2]))
    assert np.array_equal(eng.relinearize(m3), o.relinearize(m3))
    assert np.array_equal(eng.exponentiate3(cts[0]), o.exponentiate3(cts[0]))
    got = eng.multiply(cts[:2], cts[1:3])
    assert np.array_equal(got[1], o.multiply(cts[1], cts[2]))


def test_vec_sum_mask_flatten_fc(eng, world):
    o, cts = world["orc"], world["cts"]
    for n in (1, 2, 20, 40):
        assert np.array_equal(eng.vec_sum(cts[0], n, 1), o.vec_sum(cts[0], n, 1)), n
    ones = np.ones(44, dtype=np.uint64)
    assert np.array_equal(eng.mask(cts[0], ones), o.mask(cts[0], ones))
    fc = eng.fc_rows(cts[:2], cts[1:3], 24, 1)
    for s in range(2):
        for r in range(2):
            want = o.vec_sum(o.relinearize(o.multiply(cts[s], cts[1 + r])), 24, 1)
            assert np.array_equal(fc[s, r], want)


def test_csp_service_calls(eng, world):
    """hhe_csp_decompose / hhe_csp_evaluate_model against the reference's call sequence (src/examples/CSP/CSP.cpp:235-323):
    per record decomposition -> (mask of the ragged last block) -> flatten, then multiply -> relinearize -> vec_sum."""
    o, keys, cts = world["orc"], world["keys"], world["cts"]
    for s in (-128, -256):
        elt = o.galois_elt(s)
        k = keys.galois_key(elt)
        o.load_ksk(1, elt, k)
        eng.load_ksk(1, elt, k)
    rng = np.random.default_rng(21)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    ek = keys.encrypt_zero_plus(o, o.encode(common.pack_key(key, N)))
    syms = [O.pasta_plain(key, common.T, rng.integers(0, 32, 300, dtype=np.uint64)) for _ in range(2)]  # 3 blocks, 44-word tail
    ones = np.ones(44, dtype=np.uint64)
    flat = None
    for apply_mask in (False, True):
        got = eng.csp_decompose(ek, np.concatenate(syms), records=2, apply_mask=apply_mask, flatten_keys=1)
        for r in range(2):
            blocks = o.pasta_decompose(ek, syms[r])
            if apply_mask:
                blocks[-1] = o.mask(blocks[-1], ones)
            assert np.array_equal(got[r], o.flatten(blocks, 1)), (apply_mask, r)
        flat = got
    res = eng.csp_evaluate_model(flat, cts[1:3], 24, 1)
    for r in range(2):
        for w in range(2):
            assert np.array_equal(res[r, w], o.vec_sum(o.relinearize(o.multiply(flat[r], cts[1 + w])), 24, 1))


def test_round_material(eng):
    for ctr, layer in ((0, 0), (0, 3), (5, 1), (2**40 + 3, 2)):
        got = eng.pasta_layer_material(common.NONCE, ctr, layer)
        want = O.pasta_layer_material(common.T, common.NONCE, ctr, layer)
        assert all(np.array_equal(g, w) for g, w in zip(got, want))


@pytest.mark.parametrize("bsgs", [False, True])
def test_transciphering_vs_oracle(eng, world, bsgs):
    o, keys, rng = world["orc"], world["keys"], np.random.default_rng(11)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    pt = rng.integers(0, common.T, 300, dtype=np.uint64)  # 3 blocks, last one ragged (44 words)
    sym = O.pasta_plain(key, common.T, pt)
    ek = keys.encrypt_zero_plus(o, o.encode(common.pack_key(key, N)))
    got = eng.pasta3_decompose(ek, sym, use_bsgs=bsgs)
    want = o.pasta_decompose(ek, sym, use_bsgs=bsgs)
    assert got.shape == want.shape and np.array_equal(got, want)
    if not bsgs:
        # records API: two records restart the counters (CSP.cpp:247-252) -> same ciphertexts as two separate calls
        two = eng.pasta3_decompose(ek, np.concatenate([sym[:256], sym[:256]]), records=2)
        assert np.array_equal(two[:2], want[:2]) and np.array_equal(two[2:], want[:2])
        # first_counter shifts the SHAKE stream
        assert np.array_equal(eng.pasta3_decompose(ek, sym[128:256], first_counter=1)[0], want[1])


@pytest.mark.parametrize("bsgs,shared_keystream", [(False, True), (True, True), (False, False)])
def test_records_sharing_one_counter(eng, world, bsgs, shared_keystream, monkeypatch):
    """Records restart at counter 0 (CSP.cpp:247-252): the keystream ciphertext of a counter is the same for
    every record and is computed once per call (default); with HHE_NO_SHARED_KEYSTREAM=1 every block is transciphered on its own
    and the batch shares only its round material, diagonals and their transforms. Same ciphertexts either way as transciphering
    each record on its own with the oracle."""
    if not shared_keystream:
        monkeypatch.setenv("HHE_NO_SHARED_KEYSTREAM", "1")
    o, keys, rng = world["orc"], world["keys"], np.random.default_rng(23)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    ek = keys.encrypt_zero_plus(o, o.encode(common.pack_key(key, N)))
    recs = [O.pasta_plain(key, common.T, rng.integers(0, common.T, 100, dtype=np.uint64)) for _ in range(3)]
    got = eng.pasta3_decompose(ek, np.concatenate(recs), use_bsgs=bsgs, records=3)
    for r in range(3):
        assert np.array_equal(got[r], o.pasta_decompose(ek, recs[r], use_bsgs=bsgs)[0]), r
    # two-block records (counters 0, 1 per record, ragged last block): both keystreams shared by both records
    recs2 = [O.pasta_plain(key, common.T, rng.integers(0, common.T, 150, dtype=np.uint64)) for _ in range(2)]
    got2 = eng.pasta3_decompose(ek, np.concatenate(recs2), use_bsgs=bsgs, records=2).reshape(2, 2, 2, o.L, N)
    for r in range(2):
        want = o.pasta_decompose(ek, recs2[r], use_bsgs=bsgs)
        assert np.array_equal(got2[r, 0], want[0]) and np.array_equal(got2[r, 1], want[1]), r


def test_caller_owned_result_buffers(eng, world):
    """pasta3_decompose / fc_rows write into a caller-owned array (pinned memory in bench.py) when one is given"""
    o, keys, rng = world["orc"], world["keys"], np.random.default_rng(29)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    ek = keys.encrypt_zero_plus(o, o.encode(common.pack_key(key, N)))
    sym = O.pasta_plain(key, common.T, rng.integers(0, common.T, 128, dtype=np.uint64))
    buf = np.zeros((1, 2, o.L, N), dtype=np.uint64)
    got = eng.pasta3_decompose(ek, sym, out=buf)
    assert np.shares_memory(got, buf) and np.array_equal(buf, eng.pasta3_decompose(ek, sym))
    with pytest.raises(pkg.HheInvalidArgument):
        eng.pasta3_decompose(ek, sym, out=np.zeros(5, dtype=np.uint64))
    with pytest.raises(pkg.HheInvalidArgument):
        eng.pasta3_decompose(ek, sym, out=np.zeros((1, 2, o.L, N), dtype=np.int64))


# ---- golden vectors generated from the reference itself ---------------------------------------------------------------
@pytest.fixture(scope="module", params=BACKENDS)
def eng512(request):
    ctx = make_ctx(request.param, int(FX["N"]), FX["q"])
    for name, kind in (("gk_m1", 0), ("gk_p128", 0), ("gk_col", 0), ("rk", 2)):
        ctx.load_ksk(kind, int(FX[name + "_elt"]), FX[name])
    yield ctx
    ctx.close()


def test_golden_primitives(eng512):
    e, a, b, pt = eng512, FX["ct_a"], FX["ct_b"], FX["pt"]
    assert np.array_equal(e.ntt(0, a[0, 0]), FX["kat_ntt_fwd"])
    assert np.array_equal(e.ntt(3, a[1, 3], inverse=True), FX["kat_ntt_inv"])
    assert np.array_equal(e.encode(FX["slots_p"]), pt)
    assert np.array_equal(e.add(a, b), FX["kat_add"])
    assert np.array_equal(e.negate(a), FX["kat_negate"])
    assert np.array_equal(e.add_plain(a, pt), FX["kat_add_plain"])
    assert np.array_equal(e.multiply_plain(a, pt), FX["kat_multiply_plain"])
    assert np.array_equal(e.rotate_rows(a, -1), FX["kat_rot_m1"])
    assert np.array_equal(e.rotate_rows(a, 128), FX["kat_rot_p128"])
    assert np.array_equal(e.rotate_columns(a), FX["kat_rot_col"])
    assert np.array_equal(e.multiply(a, b), FX["kat_multiply"])
    assert np.array_equal(e.square(a), FX["kat_square"])
    assert np.array_equal(e.relinearize(FX["kat_multiply"]), FX["kat_relin"])
    assert np.array_equal(e.exponentiate3(a), FX["kat_exp3"])
    assert np.array_equal(e.mask(a, np.ones(44, dtype=np.uint64)), FX["kat_mask"])


def test_golden_material(eng512):
    for tag, ctr, layer in (("c0l0", 0, 0), ("c0l3", 0, 3), ("c5l1", 5, 1)):
        m1, m2, rc = eng512.pasta_layer_material(common.NONCE, ctr, layer)
        assert np.array_equal(m1, FX["mat1_" + tag]) and np.array_equal(m2, FX["mat2_" + tag]) and np.array_equal(rc, FX["rc_" + tag])


def test_golden_transciphering(eng512):
    got = eng512.pasta3_decompose(FX["enc_key"], FX["sym_ct"])
    assert np.array_equal(got, FX["decomposed"])


# ---- N=2048: log2(N/2) = 10 = 1 mod 3, the same pass schedule class as N=16384 (13 = 1 mod 3): covers the key-switch
# prologue that folds two Cooley-Tukey stages into the load, which the N=1024 / N=512 rings do not reach -------------
@pytest.mark.parametrize("backend", BACKENDS)
def test_n2048_fold_schedule(backend):
    NN = 2048
    q = common.small_params(NN, 3, 48)
    orc = O.Oracle(NN, common.T, q)
    keys = ToyKeys(orc, 5)
    ctx = make_ctx(backend, NN, q)
    assert ctx.info()["fp64_moduli"] == len(q)
    e1 = orc.galois_elt(-1)
    for kind, elt, k in ((0, e1, keys.galois_key(e1)), (2, 0, keys.relin_key())):
        orc.load_ksk(kind, elt, k)
        ctx.load_ksk(kind, elt, k)
    rng = np.random.default_rng(3)
    a = keys.encrypt_zero_plus(orc, orc.encode(rng.integers(0, common.T, NN, dtype=np.uint64)))
    b = keys.encrypt_zero_plus(orc, orc.encode(rng.integers(0, common.T, NN, dtype=np.uint64)))
    assert np.array_equal(ctx.rotate_rows(a, -1), orc.rotate_rows(a, -1))
    m3 = orc.multiply(a, b)
    assert np.array_equal(ctx.multiply(a, b), m3)
    assert np.array_equal(ctx.relinearize(m3), orc.relinearize(m3))
    pt = orc.encode(rng.integers(0, common.T, 500, dtype=np.uint64))
    assert np.array_equal(ctx.multiply_plain(a, pt), orc.multiply_plain(a, pt))
    x = rng.integers(0, int(q[1]), NN, dtype=np.uint64)
    assert np.array_equal(ctx.ntt(1, x), orc.ntt(1, x)) and np.array_equal(ctx.ntt(1, x, inverse=True), orc.ntt(1, x, inverse=True))
    ctx.close()
    orc.close()


# ---- the N = 32768 code path (split transforms, quarter-split key switch), forced at N = 1024 so the oracle can check it
@pytest.mark.parametrize("backend", BACKENDS)
def test_split_path_forced(backend, monkeypatch):
    q = common.small_params(N, 3, 50)
    orc = O.Oracle(N, common.T, q)
    keys = ToyKeys(orc, 9)
    monkeypatch.setenv("HHE_FORCE_SPLIT", "1")
    ctx = make_ctx(backend, N, q)
    monkeypatch.delenv("HHE_FORCE_SPLIT")
    e1 = orc.galois_elt(-1)
    for kind, elt, k in ((0, e1, keys.galois_key(e1)), (2, 0, keys.relin_key())):
        orc.load_ksk(kind, elt, k)
        ctx.load_ksk(kind, elt, k)
    rng = np.random.default_rng(4)
    for limb in range(2 * orc.K):
        mod = int(orc.q[limb]) if limb < orc.K else int(orc.behz()["base_B"][limb - orc.K]) if limb - orc.K < orc.L else orc.behz()["m_sk"]
        x = rng.integers(0, mod, (2, N), dtype=np.uint64)
        f = ctx.ntt(limb, x)
        assert np.array_equal(f, np.stack([orc.ntt(limb, v) for v in x])), limb
        assert np.array_equal(ctx.ntt(limb, f, inverse=True), x), limb
    a = keys.encrypt_zero_plus(orc, orc.encode(rng.integers(0, common.T, N, dtype=np.uint64)))
    b = keys.encrypt_zero_plus(orc, orc.encode(rng.integers(0, common.T, N, dtype=np.uint64)))
    assert np.array_equal(ctx.rotate_rows(a, -1), orc.rotate_rows(a, -1))
    m3 = orc.multiply(a, b)
    assert np.array_equal(ctx.multiply(a, b), m3)
    assert np.array_equal(ctx.relinearize(m3), orc.relinearize(m3))
    # the whole-limb operations as split transforms + element-wise pieces (encode, multiply_plain incl. a monomial, mask) and a
    # whole PASTA-3 block on this path (what N = 32768 runs)
    sl = rng.integers(0, common.T, 300, dtype=np.uint64)
    assert np.array_equal(ctx.encode(sl), orc.encode(sl))
    pt = orc.encode(sl)
    assert np.array_equal(ctx.multiply_plain(a, pt), orc.multiply_plain(a, pt))
    mono = np.zeros(N, dtype=np.uint64)
    mono[3] = common.T - 2
    assert np.array_equal(ctx.multiply_plain(a, mono), orc.multiply_plain(a, mono))
    assert np.array_equal(ctx.mask(a, np.ones(20, dtype=np.uint64)), orc.mask(a, np.ones(20, dtype=np.uint64)))
    for s_ in (0, 128):
        e = orc.galois_elt(s_)
        k = keys.galois_key(e)
        orc.load_ksk(0, e, k)
        ctx.load_ksk(0, e, k)
    key256 = rng.integers(0, common.T, 256, dtype=np.uint64)
    enc_key = keys.encrypt_zero_plus(orc, orc.encode(common.pack_key(key256, N)))
    sym = O.pasta_plain(key256, common.T, rng.integers(0, common.T, 128, dtype=np.uint64))
    assert np.array_equal(ctx.pasta3_decompose(enc_key, sym), orc.pasta_decompose(enc_key, sym))
    ctx.close()
    orc.close()


# ---- edge cases and error behaviour (the reference throws std::invalid_argument / std::logic_error here) -------------
def test_edge_cases_and_errors(eng, world):
    o, keys, cts = world["orc"], world["keys"], world["cts"]
    ek = cts[0]
    # empty input: zero blocks, no device work
    assert eng.pasta3_decompose(ek, np.zeros(0, dtype=np.uint64)).shape[0] == 0
    # a single word is one ragged block
    one = eng.pasta3_decompose(ek, np.array([5], dtype=np.uint64))
    assert np.array_equal(one, o.pasta_decompose(ek, np.array([5], dtype=np.uint64)))
    # words must be residues mod t (BatchEncoder::encode would throw)
    with pytest.raises(pkg.HheInvalidArgument):
        eng.pasta3_decompose(ek, np.array([common.T], dtype=np.uint64))
    # flatten of a single ciphertext is the identity; vec_sum(n=1) too
    assert np.array_equal(eng.flatten(cts[:1]), cts[0])
    assert np.array_equal(eng.vec_sum(cts[0], 1, 1), cts[0])
    # wrong buffer sizes are rejected before anything is uploaded
    with pytest.raises(pkg.HheInvalidArgument):
        eng.add(cts[0][:1], cts[1][:1])
    with pytest.raises(pkg.HheInvalidArgument):
        eng.load_ksk(0, 3, np.zeros(10, dtype=np.uint64))
    # all-zero mask: SEAL reports a transparent ciphertext
    with pytest.raises(pkg.HheLogicError):
        eng.mask(cts[0], np.zeros(4, dtype=np.uint64))


@pytest.mark.parametrize("backend", BACKENDS)
def test_missing_keys_raise(backend):
    q = common.small_params(N, 2, 50)
    ctx = make_ctx(backend, N, q)
    ct = np.zeros((2, 2, N), dtype=np.uint64)
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.relinearize(np.zeros((3, 2, N), dtype=np.uint64))  # no relin key loaded
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.rotate_rows(ct, -1)  # "Galois key not present"
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.pasta3_decompose(ct, np.arange(4, dtype=np.uint64))  # PASTA needs steps -1, +128, columns and relin
    ctx.close()


# ---- client side: plain PASTA-3 on the device ------------------------------------------------
@pytest.mark.parametrize("backend", BACKENDS)
def test_plain_pasta3_matches_reference_kat_and_oracle(backend):
    """pasta::PASTA::encrypt / decrypt (src/pasta/pasta_3_plain.cpp:9-47): SURVEY.md Appendix E KAT (generated by the reference,
    committed in tests/golden), the oracle on ragged lengths, and the encrypt -> decrypt round trip."""
    ctx = make_ctx(backend, N, common.small_params(N, 3, 48))
    key = FX["sym_key"]
    # KAT: plaintext 0..255 (two blocks) under get_symmetric_key()
    kat = ctx.pasta3_plain(key, FX["kat_plain"])
    assert np.array_equal(kat, FX["kat_cipher"])
    assert list(kat[:8]) == [30446, 62406, 62716, 38766, 43125, 6036, 63532, 7424] and list(kat[128:132]) == [11021, 61753, 2637, 43237]
    rng = np.random.default_rng(12)
    for n in (1, 127, 128, 129, 300, 5 * 128 + 17):
        words = rng.integers(0, common.T, n, dtype=np.uint64)
        enc = ctx.pasta3_plain(key, words)
        assert np.array_equal(enc, O.pasta_plain(key, common.T, words, decrypt=False))
        assert np.array_equal(ctx.pasta3_plain(key, enc, decrypt=True), words)
    # counters: block b of a long call == block 0 of a call that starts at counter b
    words = rng.integers(0, common.T, 3 * 128, dtype=np.uint64)
    whole = ctx.pasta3_plain(key, words)
    assert np.array_equal(whole[256:], ctx.pasta3_plain(key, words[256:], first_counter=2))
    # unreduced inputs: encrypt reduces, decrypt does not (pasta_3_plain.cpp:20-23 vs :39-41)
    odd = np.array([common.T + 5, 2 * common.T, 70000], dtype=np.uint64)
    assert np.array_equal(ctx.pasta3_plain(key, odd), O.pasta_plain(key, common.T, odd, decrypt=False))
    assert ctx.pasta3_plain(key, np.zeros(0, dtype=np.uint64)).size == 0
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.pasta3_plain(key[:100], words)
    ctx.close()


def test_fullsize_ring_on_the_emulation_harness():
    """N = 16384 with the BFVDefault primes is the only shape where a half-limb CTA has S = 16 x 512 residues: the folded load fused with
    the first register pass (kernels.h fwd_half_fused_f64) and the 512-thread tensor-memory key switch exist only there. The kernel
    index arithmetic of that shape is checked here against the oracle (CPU tier); the B200 tests repeat it against SEAL."""
    NN = 16384
    q = common.Q_16384
    orc = O.Oracle(NN, common.T, q)
    ctx = make_ctx("emul", NN, q)
    rng = np.random.default_rng(21)
    for limb in (0, 4, 8):
        x = rng.integers(0, int(q[limb]), NN, dtype=np.uint64)
        f = ctx.ntt(limb, x)
        assert np.array_equal(f, orc.ntt(limb, x)), limb
        assert np.array_equal(ctx.ntt(limb, f, inverse=True), x), limb
    keys = ToyKeys(orc, 5)
    e1 = orc.galois_elt(-1)
    k1 = keys.galois_key(e1)
    orc.load_ksk(0, e1, k1)
    ctx.load_ksk(0, e1, k1)
    a = keys.encrypt_zero_plus(orc, orc.encode(rng.integers(0, common.T, 300, dtype=np.uint64)))
    assert np.array_equal(ctx.rotate_rows(a, -1), orc.rotate_rows(a, -1))               # cluster-8 key switch (1 item)
    five = np.stack([a] * 5)
    assert np.array_equal(ctx.rotate_rows(five, -1)[4], orc.rotate_rows(a, -1))          # tensor-memory key switch (5 items)
    pt = orc.encode(rng.integers(0, common.T, 500, dtype=np.uint64))
    assert np.array_equal(ctx.multiply_plain(a, pt), orc.multiply_plain(a, pt))          # lift_ntt + ntt_mac
    ctx.close()
    orc.close()
