"""Pins the oracle against the reference itself (oracle/_ref/libhhe_ref.so = unmodified reference sources + libseal).
Skipped where that library was not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

import common
from oracle import oracle as O
from oracle import refshim as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libhhe_ref.so not built")

N = 1024


@pytest.fixture(scope="module")
def pair():
    q = common.small_params(N, 6)
    ref = R.Ref(N, common.T, q, seed=7, steps=(0, -1, 128, -16, -32, -48, -64, -80, -96, -112), default_gk=True)
    orc = O.Oracle(N, common.T, q)
    common.load_keys_from_ref(orc, ref)
    yield ref, orc
    orc.close()
    ref.close()


def test_constants(pair):
    ref, orc = pair
    assert np.array_equal(ref.ntt_roots()[0], orc.ntt_roots()[0]) and ref.ntt_roots()[1] == orc.ntt_roots()[1]
    rb, ob = ref.behz(), orc.behz()
    assert all(np.array_equal(np.asarray(rb[k]), np.asarray(ob[k])) for k in rb)
    for s in (0, 1, -1, 128, -128, -16, 5, -300, 511, -511):
        assert ref.galois_elt(s) == orc.galois_elt(s)


@pytest.mark.parametrize("NN,bits", [(8192, None), (16384, None)])
def test_default_parameter_constants(NN, bits):
    """BFVDefault moduli of the BASELINE.json rings: roots and BEHZ base must equal SEAL's."""
    ref = R.Ref(NN, common.T, None, seed=1, steps=(), default_gk=False)
    orc = O.Oracle(NN, common.T, ref.q)
    assert list(ref.q) == (common.Q_16384 if NN == 16384 else common.Q_8192)
    assert np.array_equal(ref.ntt_roots()[0], orc.ntt_roots()[0]) and ref.ntt_roots()[1] == orc.ntt_roots()[1]
    rb, ob = ref.behz(), orc.behz()
    assert all(np.array_equal(np.asarray(rb[k]), np.asarray(ob[k])) for k in rb)
    x = np.random.default_rng(0).integers(0, int(ref.q[0]), NN, dtype=np.uint64)
    assert np.array_equal(ref.ntt(0, x), orc.ntt(0, x))
    assert np.array_equal(ref.ntt(ref.K - 1, x, True), orc.ntt(ref.K - 1, x, True))
    orc.close()
    ref.close()


def test_ops(pair):
    ref, orc = pair
    rng = np.random.default_rng(1)
    x = rng.integers(0, int(ref.q[0]), N, dtype=np.uint64)
    for limb in (0, ref.K - 1):
        assert np.array_equal(ref.ntt(limb, x), orc.ntt(limb, x))
        assert np.array_equal(ref.ntt(limb, x, True), orc.ntt(limb, x, True))
    assert np.array_equal(ref.ntt_bsk(2, x), orc.ntt(ref.K + 2, x))
    sl = rng.integers(0, common.T, N, dtype=np.uint64)
    for n in (N, 300, 1, 0):
        assert np.array_equal(ref.encode(sl[:n]), orc.encode(sl[:n]))
    a, b = ref.encrypt(sl), ref.encrypt(rng.integers(0, common.T, N, dtype=np.uint64))
    pt = ref.encode(rng.integers(0, common.T, 700, dtype=np.uint64))
    assert np.array_equal(ref.add(a, b), orc.add(a, b))
    assert np.array_equal(ref.negate(a), orc.negate(a))
    assert np.array_equal(ref.add_plain(a, pt), orc.add_plain(a, pt))
    assert np.array_equal(ref.multiply_plain(a, pt), orc.multiply_plain(a, pt))
    for s, ks in ((-1, 0), (128, 0), (-16, 0), (1, 1), (-5, 1), (7, 1), (-300, 1), (511, 1)):
        assert np.array_equal(ref.rotate_rows(a, s, ks), orc.rotate_rows(a, s, ks)), s
    assert np.array_equal(ref.rotate_columns(a), orc.rotate_columns(a))
    m3 = ref.multiply(a, b)
    assert np.array_equal(m3, orc.multiply(a, b))
    assert np.array_equal(ref.square(a), orc.multiply(a, a))
    assert np.array_equal(ref.relinearize(m3), orc.relinearize(m3))
    assert np.array_equal(ref.exponentiate3(a), orc.exponentiate3(a))
    assert np.array_equal(ref.vec_sum(a, 20, 1), orc.vec_sum(a, 20, 1))
    ones = np.ones(44, dtype=np.uint64)
    assert np.array_equal(ref.mask(a, ones), orc.mask(a, ones))


def test_flatten():
    q = common.small_params(N, 6)
    ref = R.Ref(N, common.T, q, seed=7, steps=(-128, -256))
    orc = O.Oracle(N, common.T, q)
    for e in ref.list_galois(0):
        orc.load_ksk(0, e, ref.ksk(0, e))
    rng = np.random.default_rng(3)
    cts = np.stack([ref.encrypt(rng.integers(0, common.T, 128, dtype=np.uint64)) for _ in range(3)])
    assert np.array_equal(ref.flatten(cts), orc.flatten(cts))


def test_pasta_plain_and_material():
    rng = np.random.default_rng(5)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    for n in (1, 128, 129, 300):
        pt = rng.integers(0, common.T, n, dtype=np.uint64)
        ct = R.pasta_plain(key, common.T, pt)
        assert np.array_equal(ct, O.pasta_plain(key, common.T, pt))
        assert np.array_equal(O.pasta_plain(key, common.T, ct, True), pt)
    for ctr in (0, 1, 77, 2**40 + 3):
        for layer in (0, 2):
            a, b = R.pasta_layer_material(common.T, common.NONCE, ctr, layer), O.pasta_layer_material(common.T, common.NONCE, ctr, layer)
            assert all(np.array_equal(u, v) for u, v in zip(a, b))


@pytest.mark.parametrize("bsgs", [False, True])
def test_decomposition(pair, bsgs):
    ref, orc = pair
    rng = np.random.default_rng(9)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    pt = rng.integers(0, common.T, 150, dtype=np.uint64)
    sym = R.pasta_plain(key, common.T, pt)
    ek = ref.encrypt(common.pack_key(key, N))
    want = ref.pasta_decompose(ek, sym, bsgs)
    got = orc.pasta_decompose(ek, sym, bsgs)
    assert np.array_equal(want, got)
    d0, budget = ref.decrypt(got[0])
    d1, _ = ref.decrypt(got[1])
    assert budget > 0 and np.array_equal(d0[:128], pt[:128]) and np.array_equal(d1[:22], pt[128:])


def test_multiply_plain_monomial_branch():
    """Evaluator::multiply_plain with a MONOMIAL plaintext (one nonzero coefficient) takes SEAL's
    negacyclic_multiply_poly_mono_coeffmod branch: with fast plain lift the coefficient is used as it is, without the centred lift,
    also in the upper half (t - 1 multiplies by 65536, not by -1). The constant-plaintext products of the second FC layer
    (host.fc2_plain_rows) hit this branch for every negative weight."""
    N = 1024
    q = common.small_params(N, 3, 48)
    ref = R.Ref(N, common.T, q, seed=2, steps=(0,), default_gk=False)
    orc = O.Oracle(N, common.T, q)
    a = ref.encrypt(np.arange(50, dtype=np.uint64))
    for coeff, pos in ((3, 0), (common.T - 1, 0), (common.T - 2, 5), (40000, N - 1)):
        pt = np.zeros(N, dtype=np.uint64)
        pt[pos] = coeff
        assert np.array_equal(orc.multiply_plain(a, pt), ref.multiply_plain(a, pt)), (coeff, pos)
    pt = np.zeros(N, dtype=np.uint64)
    pt[0], pt[7] = common.T - 1, 2  # two coefficients: the general branch with the centred lift
    assert np.array_equal(orc.multiply_plain(a, pt), ref.multiply_plain(a, pt))
    ref.close()
