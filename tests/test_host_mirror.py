"""The Python mirror of the reference's class/helper interface (privacy-preserving-ml-through-hhe_b200/host.py: PASTA_SEAL,
packed_enc_multiply, encrypted_vec_sum, decompose, evaluate_model) driven the way src/examples/CSP/CSP.cpp:235-323 drives the
reference, on the emulation harness (CPU tier), against the oracle's restatement of the same call sequence. The GPU tier runs the
same mirror at N=16384 against SEAL itself (tests/test_gpu_fc.py)."""
import importlib

import numpy as np
import pytest

import common
from oracle import oracle as O
from test_engine_parity import ToyKeys, make_ctx

pkg = common.package()
host = importlib.import_module(common.PKG + ".host")
N = 1024


@pytest.fixture(scope="module")
def setup():
    q = common.small_params(N, 3, 48)  # FP64-pipe kernels, as the BFVDefault(16384) ring
    orc = O.Oracle(N, common.T, q)
    keys = ToyKeys(orc, 77)
    gk0 = {orc.galois_elt(s): keys.galois_key(orc.galois_elt(s)) for s in (0, -1, 128)}
    gk1 = {orc.galois_elt(s): keys.galois_key(orc.galois_elt(s)) for s in (-128, 1, -1, 2, -2, 4, -4, 8, -8)}
    rk = keys.relin_key()
    for elt, k in gk0.items():
        orc.load_ksk(0, elt, k)
    for elt, k in gk1.items():
        orc.load_ksk(1, elt, k)
    orc.load_ksk(2, 0, rk)
    ctx = make_ctx("emul", N, q)
    yield dict(orc=orc, keys=keys, gk0=gk0, gk1=gk1, rk=rk, ctx=ctx)
    ctx.close()


def test_csp_call_sequence_through_the_mirror(setup):
    o, keys, ctx = setup["orc"], setup["keys"], setup["ctx"]
    hhe = host.PASTA_SEAL(ctx, relin_key=setup["rk"], galois_keys=setup["gk0"])  # the by-value key copies of the reference ctor
    assert hhe.get_plain_size() == 128 and "PASTA-SEAL" in hhe.get_cipher_name()
    assert hhe.add_gk_indices() == [0, -1, 128]
    hhe.activate_bsgs(True)
    assert hhe.add_gk_indices()[-7:] == [-16 * k for k in range(1, 8)]
    hhe.activate_bsgs(False)
    with pytest.raises(pkg.HheInvalidArgument):
        hhe.HE_decrypt(np.zeros(4, dtype=np.uint64))  # secret_key_encrypted not set

    rng = np.random.default_rng(3)
    key = rng.integers(0, common.T, 256, dtype=np.uint64)
    enc_key = keys.encrypt_zero_plus(o, o.encode(common.pack_key(key, N)))
    x = rng.integers(0, 32, 140, dtype=np.uint64)  # 2 blocks, the second holds 12 words
    sym = O.pasta_plain(key, common.T, x)

    # BaseCSP::decompose: decomposition -> mask of the ragged block -> flatten with the dedicated keys (csp_gk: step -128)
    want_blocks = o.pasta_decompose(enc_key, sym)
    want_masked = np.stack([want_blocks[0], o.mask(want_blocks[1], np.ones(12, dtype=np.uint64))])
    e128 = o.galois_elt(-128)
    flat = host.decompose(hhe, [sym], [enc_key], 140, flatten_keys={e128: setup["gk1"][e128]}, mask_in_place=True)
    assert len(flat) == 1 and np.array_equal(flat[0], o.flatten(want_masked, 1))
    # the service's mask-on-a-copy quirk = flatten of the unmasked blocks
    assert np.array_equal(hhe.flatten(want_blocks, {e128: setup["gk1"][e128]}), o.flatten(want_blocks, 1))
    # flatten with the PASTA key set only: the step -128 key is missing there, as SEAL's std::invalid_argument
    with pytest.raises(pkg.HheInvalidArgument):
        hhe.flatten(want_blocks)
    assert np.array_equal(hhe.mask(want_blocks[1], np.ones(12, dtype=np.uint64)), want_masked[1])

    # CSP_hhe_pktnn_1fc::evaluateModel: packed_enc_multiply -> relinearize -> encrypted_vec_sum per weight row
    for elt, k in setup["gk1"].items():
        ctx.load_ksk(pkg.KEYSET_1, elt, k)
    w = np.stack([keys.encrypt_zero_plus(o, o.encode(rng.integers(0, 5, 8, dtype=np.uint64))) for _ in range(2)])
    out = host.evaluate_model(ctx, [flat[0]], w, 8)
    assert out.shape[:2] == (1, 2)
    for r in range(2):
        want = o.vec_sum(o.relinearize(o.multiply(flat[0], w[r])), 8, 1)
        assert np.array_equal(out[0, r], want)
        prod = host.packed_enc_multiply(ctx, flat[0], w[r])
        assert np.array_equal(host.encrypted_vec_sum(ctx, ctx.relinearize(prod), 8), want)


def test_second_fc_layer_composition(setup):
    """fc1 -> encrypted square activation -> fc2 (the reference's TODO, hhe_pktnn_examples.cpp:993-997) on the
    emulation harness against the oracle's restatement of the same SEAL op sequence (square, relinearize, multiply_plain with a
    constant plaintext, add)."""
    o, keys, ctx = setup["orc"], setup["keys"], setup["ctx"]
    for elt, k in setup["gk1"].items():
        ctx.load_ksk(pkg.KEYSET_1, elt, k)
    ctx.load_ksk(pkg.RELIN, 0, setup["rk"])
    rng = np.random.default_rng(11)
    n, H = 8, 3
    x = rng.integers(0, 4, n, dtype=np.uint64)
    W1 = rng.integers(-2, 3, (H, n))
    W2 = np.array([[1, -2, 3], [0, 1, -1]])
    cx = keys.encrypt_zero_plus(o, o.encode(x))
    enc_w1 = np.stack([keys.encrypt_zero_plus(o, o.encode(np.mod(W1[j], common.T).astype(np.uint64))) for j in range(H)])
    got = host.evaluate_model_2fc(ctx, [cx], enc_w1, n, W2)[0]
    # the same sequence on the oracle
    hidden = [o.vec_sum(o.relinearize(o.multiply(cx, enc_w1[j])), n, 1) for j in range(H)]
    sq = [o.relinearize(o.multiply(h, h)) for h in hidden]
    for k in range(2):
        acc = None
        for j in range(H):
            if W2[k, j] == 0:
                continue
            pt = np.zeros(N, dtype=np.uint64)
            pt[0] = abs(int(W2[k, j]))
            term = o.multiply_plain(sq[j], pt)
            if W2[k, j] < 0:
                term = o.negate(term)
            acc = term if acc is None else o.add(acc, term)
        assert np.array_equal(got[k], acc)
    with pytest.raises(pkg.HheInvalidArgument):
        host.fc2_plain_rows(ctx, np.stack(sq), [[0, 0, 0]])
