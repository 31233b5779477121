"""Encrypted FC layer at full size on the B200 (BASELINE.json configs[1]/[2] shapes), written like the reference's
monolithic demos (src/examples/hhe_pktnn_examples.cpp:385-711, 713-1010): transcipher -> mask -> flatten ->
packed_enc_multiply -> relinearize -> encrypted_vec_sum with the analyst's default Galois keys -> decrypt slot n-1 ->
compare with the plaintext dot product / class prediction. Inputs are synthetic (the MNIST / MIT-BIH inputs are missing
from the reference checkout); weights are synthetic integers in the trained range."""
import importlib

import numpy as np
import pytest

import common
from oracle import oracle as O
from oracle import refshim as R

pkg = common.package()
host = importlib.import_module(common.PKG + ".host")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libhhe_ref.so not built")]
N, T = 16384, common.T


def signed(v):
    v = int(v)
    return v - T if v > T // 2 else v


@pytest.fixture(scope="module")
def world():
    steps = (0, -1, 128) + tuple(-128 * i for i in range(1, 7))
    ref = R.Ref(N, T, None, seed=21, steps=steps, default_gk=True)
    ctx = pkg.Context(N, T, ref.q, device=0)
    common.load_keys_from_ref(ctx, ref, keysets=(0, 1))
    rng = np.random.default_rng(8)
    key = rng.integers(0, T, 256, dtype=np.uint64)
    hhe = host.PASTA_SEAL(ctx)
    yield dict(ref=ref, ctx=ctx, hhe=hhe, rng=rng, key=key, enc_key=ref.encrypt(common.pack_key(key, N)))
    ctx.close()
    ref.close()


def test_ecg_row_bit_exact_with_seal(world):
    """configs[2] shape (128 -> 1): one sample, compared limb for limb with the reference (13.7 s of SEAL on one core)."""
    ref, ctx, hhe, rng = world["ref"], world["ctx"], world["hhe"], world["rng"]
    x = rng.integers(0, 256, 128, dtype=np.uint64)
    w = rng.integers(-128, 128, 128)
    sym = O.pasta_plain(world["key"], T, x)
    c = hhe.decomposition(sym, [world["enc_key"]], True)[0]
    enc_w = ref.encrypt(np.mod(w, T).astype(np.uint64))
    prod = host.packed_enc_multiply(ctx, c, enc_w)
    assert np.array_equal(prod, ref.multiply(c, enc_w))
    lin = ctx.relinearize(prod)
    assert np.array_equal(lin, ref.relinearize(prod))
    got = host.encrypted_vec_sum(ctx, lin, 128)
    assert np.array_equal(got, ref.vec_sum(lin, 128, 1))
    slots, budget = ref.decrypt(got)
    assert budget > 0 and signed(slots[127]) == int(np.dot(x.astype(np.int64), w)) % T - (T if int(np.dot(x.astype(np.int64), w)) % T > T // 2 else 0)


def test_mnist_sample_prediction(world):
    """configs[1] shape (784 -> 10): 7 blocks, mask(16 ones) on the last, flatten, 10 encrypted weight rows."""
    ref, ctx, hhe, rng = world["ref"], world["ctx"], world["hhe"], world["rng"]
    x = rng.integers(0, 256, 784, dtype=np.uint64)
    W = rng.integers(-8, 9, (10, 784))
    sym = O.pasta_plain(world["key"], T, x)
    flat = host.decompose(hhe, [sym], [world["enc_key"]], 784, mask_in_place=True)[0]
    slots, _ = ref.decrypt(flat)
    assert np.array_equal(slots[:784], x) and not slots[784:8192].any()
    enc_w = np.stack([ref.encrypt(np.mod(W[r], T).astype(np.uint64)) for r in range(10)])
    out = host.evaluate_model(ctx, [flat], enc_w, 784)[0]
    logits = [signed(ref.decrypt(out[r])[0][783]) for r in range(10)]
    want = [int(v) for v in W @ x.astype(np.int64)]
    assert [v % T for v in logits] == [v % T for v in want]
    if max(abs(v) for v in want) < T // 2:
        assert int(np.argmax(logits)) == int(np.argmax(want))
    # one output neuron limb-exact against the reference's own op sequence would take ~100 s of SEAL: check the
    # first 40 rotations' partial sum instead (same NAF chains, same keys)
    lin = ctx.relinearize(ctx.multiply(flat, enc_w[0]))
    assert np.array_equal(ctx.vec_sum(lin, 40), ref.vec_sum(lin, 40, 1))


def test_missing_default_key_raises_like_seal(world):
    ctx = world["ctx"]
    c = world["enc_key"]
    with pytest.raises(pkg.HheInvalidArgument):
        ctx.vec_sum(c, 5, keys=0)  # keyset 0 has no power-of-two keys: NAF(-2) = single term without a key


def test_mnist_output_neuron_limb_exact(world):
    """ONE output neuron of the 784 -> 10 layer, every limb of the result ciphertext equal to
    the reference's own op sequence -- multiply, relinearize, encrypted_vec_sum(784): 783 rotations from the same input,
    2,875 key switches of SEAL's NAF chains (sealhelper.cpp:379-392), about 100 s of SEAL on one host core."""
    ref, ctx, rng = world["ref"], world["ctx"], world["rng"]
    x = np.zeros(N // 2, dtype=np.uint64)
    x[:784] = rng.integers(0, 256, 784, dtype=np.uint64)
    w = rng.integers(-8, 9, 784)
    cx, cw = ref.encrypt(x), ref.encrypt(np.mod(w, T).astype(np.uint64))
    lin = ctx.relinearize(ctx.multiply(cx, cw))
    assert np.array_equal(lin, ref.relinearize(ref.multiply(cx, cw)))
    got = ctx.vec_sum(lin, 784)
    assert np.array_equal(got, ref.vec_sum(lin, 784, 1))
    slots, budget = ref.decrypt(got)
    assert budget > 0 and signed(slots[783]) == int(np.dot(x[:784].astype(np.int64), w))


def test_siesta_record_flattened_limb_exact(world):
    """A 300-word record (the service's inputLen, CSPRPC.cpp:196): 3 blocks -> mask of the last block (44 ones) -> flatten with the
    dedicated keys for -128 and -256, both the demo behaviour (mask in place) and the service's (mask on a copy), every limb
    against the reference's PASTA_SEAL::decomposition / mask / flatten (SEAL_Cipher.cpp:161-181)."""
    ref, ctx, hhe, rng = world["ref"], world["ctx"], world["hhe"], world["rng"]
    rec = rng.integers(0, 32, 300, dtype=np.uint64)
    sym = O.pasta_plain(world["key"], T, rec)
    want_blocks = ref.pasta_decompose(world["enc_key"], sym)
    got_blocks = hhe.decomposition(sym, [world["enc_key"]], True)
    assert np.array_equal(got_blocks, want_blocks)
    ones = np.ones(300 % 128, dtype=np.uint64)
    masked = [want_blocks[0], want_blocks[1], ref.mask(want_blocks[2], ones)]
    want_demo, want_service = ref.flatten(np.stack(masked), 0), ref.flatten(want_blocks, 0)
    assert np.array_equal(host.decompose(hhe, [sym], [world["enc_key"]], 300, mask_in_place=True)[0], want_demo)
    assert np.array_equal(host.decompose(hhe, [sym], [world["enc_key"]], 300)[0], want_service)
    # the one-call service entry point, both mask modes (keyset 0 holds the -128 i keys in this fixture)
    assert np.array_equal(ctx.csp_decompose(world["enc_key"], sym, records=1, apply_mask=True, flatten_keys=0)[0], want_demo)
    assert np.array_equal(ctx.csp_decompose(world["enc_key"], sym, records=1, apply_mask=False, flatten_keys=0)[0], want_service)
    slots, _ = ref.decrypt(want_demo)
    assert np.array_equal(slots[:300], rec) and not slots[300:8192].any()


def test_two_fc_layers_limb_exact_and_decrypt(world):
    """SURVEY.md 8 f.3 at full size: transciphered input -> fc1 (4 hidden neurons, n = 128) -> encrypted square activation ->
    fc2 (2 outputs), every ciphertext of the second layer limb-exact against SEAL running the same op sequence (square,
    relinearize, multiply_plain by a constant plaintext, negate, add), the decrypted outputs equal to the plaintext network
    (notebooks/mnist_hhe_plain.ipynb: fc1 -> x^2 -> fc2) and noise budget left. Weights are small so that every value fits t."""
    ref, ctx, hhe, rng = world["ref"], world["ctx"], world["hhe"], world["rng"]
    n, H = 128, 4
    x = rng.integers(0, 3, n, dtype=np.uint64)
    W1 = rng.integers(-1, 2, (H, n))
    W2 = np.array([[1, -1, 2, 1], [-2, 1, 0, 3]])
    h_plain = W1 @ x.astype(np.int64)
    out_plain = W2 @ (h_plain ** 2)
    assert np.abs(h_plain ** 2).max() < T // 2 and np.abs(out_plain).max() < T // 2, "the synthetic network must fit the plain modulus"
    sym = O.pasta_plain(world["key"], T, x)
    c = hhe.decomposition(sym, [world["enc_key"]], True)[0]
    enc_w1 = np.stack([ref.encrypt(np.mod(W1[j], T).astype(np.uint64)) for j in range(H)])
    hidden = host.evaluate_model(ctx, [c], enc_w1, n)[0]
    sq = host.square_activation(ctx, hidden)
    got = host.fc2_plain_rows(ctx, sq, W2)
    for j in range(H):
        assert np.array_equal(sq[j], ref.relinearize(ref.square(hidden[j])))
    for k in range(2):
        acc = None
        for j in range(H):
            if W2[k, j] == 0:
                continue
            pt = np.zeros(N, dtype=np.uint64)
            pt[0] = abs(int(W2[k, j]))
            term = ref.multiply_plain(sq[j], pt)
            if W2[k, j] < 0:
                term = ref.negate(term)
            acc = term if acc is None else ref.add(acc, term)
        assert np.array_equal(got[k], acc)
        slots, budget = ref.decrypt(got[k])
        assert budget > 0 and signed(slots[n - 1]) == int(out_plain[k])
    assert np.array_equal(host.evaluate_model_2fc(ctx, [c], enc_w1, n, W2)[0], got)
