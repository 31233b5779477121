"""Host-side mirror of the reference's interface for the hot path (Python flavour; the C++ flavour is
host/hhe_seal_shim.h). Same names, argument meaning and error behaviour as

    pasta::PASTA_SEAL            src/pasta/pasta_3_seal.h:9-54      (decomposition, HE_decrypt, activate_bsgs)
    pasta::SEALZpCipher          src/pasta/SEAL_Cipher.h:79,87-88   (mask, flatten)
    sealhelper::*                src/util/sealhelper.h:84-87,125-129 (packed_enc_multiply, encrypted_vec_sum)
    BaseCSP::decompose / CSP_hhe_pktnn_1fc::evaluateModel  src/examples/CSP/CSP.cpp:235-323

Ciphertexts are numpy uint64 arrays in SEAL's layout ([size][L][N]); keys are raw KSwitchKeys arrays.
"""
import numpy as np

from . import KEYSET_0, KEYSET_1, RELIN, Context, HheInvalidArgument  # noqa: F401

PASTA_T = 128
NONCE = 123456789


class PASTA_SEAL:
    """pasta::PASTA_SEAL. `relin_key` and `galois_keys` ({galois_elt: ksk}) play the role of the rk/gk constructor copies."""

    def __init__(self, context: Context, relin_key=None, galois_keys=None):
        self.ctx = context
        self.use_bsgs = False
        self.secret_key_encrypted = None
        if relin_key is not None:
            context.load_ksk(RELIN, 0, relin_key)
        for elt, k in (galois_keys or {}).items():
            context.load_ksk(KEYSET_0, elt, k)

    def get_cipher_name(self):
        return "PASTA-SEAL (n=128,r=3) [B200]"

    def get_plain_size(self):
        return PASTA_T

    def activate_bsgs(self, activate):
        self.use_bsgs = bool(activate)

    def add_gk_indices(self):
        """pasta_3_seal.cpp:190-201"""
        idx = [0, -1]
        if 2 * PASTA_T != self.ctx.N:
            idx.append(PASTA_T)
        if self.use_bsgs:
            idx += [-16 * k for k in range(1, 8)]
        return idx

    def decomposition(self, ciphertext, enc_ssk, batch_encoder=False):
        """pasta_3_seal.cpp:106-172: counters restart at 0 for every call; `batch_encoder` is ignored as in the reference."""
        key = enc_ssk[0] if isinstance(enc_ssk, (list, tuple)) else enc_ssk
        return self.ctx.pasta3_decompose(key, ciphertext, use_bsgs=self.use_bsgs, nonce=NONCE, first_counter=0)

    def HE_decrypt(self, ciphertext, batch_encoder=False):
        """pasta_3_seal.cpp:42-104 (the member key must have been set, as encrypt_key does in the reference)."""
        if self.secret_key_encrypted is None:
            raise HheInvalidArgument(-1, "secret_key_encrypted is not set")
        return self.decomposition(ciphertext, self.secret_key_encrypted)

    def mask(self, cipher, mask):
        return self.ctx.mask(cipher, mask)

    def flatten(self, cts, galois_keys=None):
        """SEAL_Cipher.cpp:170-181. `galois_keys` ({elt: ksk}) are uploaded as keyset 1, like the dedicated csp_gk."""
        if galois_keys:
            # the dictionary replaces keyset 1 as a whole, like a seal::GaloisKeys object: a rotation it lacks must not find
            # a key left over from an earlier object (SEAL would fall back to the NAF terms or throw)
            self.ctx.clear_keyset(KEYSET_1)
            for elt, k in galois_keys.items():
                self.ctx.load_ksk(KEYSET_1, elt, k)
            return self.ctx.flatten(cts, keys=KEYSET_1)
        return self.ctx.flatten(cts, keys=KEYSET_0)


def packed_enc_multiply(ctx: Context, encrypted1, encrypted2):
    """sealhelper::packed_enc_multiply (src/util/sealhelper.cpp:268-274) -> size-3 ciphertext"""
    return ctx.multiply(encrypted1, encrypted2)


def encrypted_vec_sum(ctx: Context, encrypted_inp, vec_size, keyset=KEYSET_1):
    """sealhelper::encrypted_vec_sum (src/util/sealhelper.cpp:379-392); the total lands in slot vec_size-1"""
    return ctx.vec_sum(encrypted_inp, vec_size, keys=keyset)


def decompose(hhe: PASTA_SEAL, records, enc_sym_key, input_len, flatten_keys=None, mask_in_place=False):
    """BaseCSP::decompose (CSP.cpp:235-283) for a list of symmetric-ciphertext records. The default, `mask_in_place=False`,
    is the reference SERVICE's behaviour -- its mask is applied to a copy (CSP.cpp:262-269) -- and the same
    default as hhe_csp_decompose / Context.csp_decompose (`apply_mask=0`); `True` masks in place like the monolithic demos
    (hhe_pktnn_examples.cpp:620-624)."""
    out = []
    rem = input_len % PASTA_T
    for rec in records:
        blocks = hhe.decomposition(np.asarray(rec, dtype=np.uint64), enc_sym_key, True)
        if rem and mask_in_place:
            blocks[-1] = hhe.mask(blocks[-1], np.ones(rem, dtype=np.uint64))
        out.append(hhe.flatten(blocks, flatten_keys) if len(blocks) > 1 else blocks[0])
    return out


def evaluate_model(ctx: Context, decomposed, enc_weights, input_len, keyset=KEYSET_1):
    """CSP_hhe_pktnn_1fc::evaluateModel (CSP.cpp:288-323) generalised to several weight rows
    (hhe_pktnn_examples.cpp:957-992): out[sample][row] = vec_sum(relin(x * w_row), input_len)."""
    return ctx.fc_rows(np.stack(decomposed), enc_weights, input_len, keys=keyset)


# ---- second layer of the 2-FC network ------------------------------------------------------------
# The reference stops after fc1 ("TODO: CSP does the encrypted square activation ... TODO: CSP evaluates the encrypted fc2
# layer", src/examples/hhe_pktnn_examples.cpp:993-997); the plaintext network is notebooks/mnist_hhe_plain.ipynb (fc1 -> x^2 ->
# fc2). fc1 leaves hidden neuron j in slot n-1 of its own ciphertext (encrypted_vec_sum), so
#   square activation  s_j = relinearize(square(h_j))                                   Evaluator::square + relinearize_inplace
#   fc2 row k          out_k = sum_j sign(w) * multiply_plain(s_j, const(|w|)), w = W2[k][j]
#                                                                    Evaluator::multiply_plain + negate_inplace + add_inplace
# where const(m) is the CONSTANT plaintext polynomial m: it multiplies every slot by m and costs only log2(m) bits of noise
# instead of the ~30 bits of a batch-encoded plaintext, which is what lets a second layer fit the budget left after fc1 (59-62
# bits at N = 16384). Negative weights are taken as a product with |w| followed by a negation, NOT as the plaintext t - |w|: a
# constant is a monomial, and SEAL multiplies monomials by the coefficient as it is (no centred lift), so t - |w| would cost 16
# bits of noise (the engine reproduces that branch bit for bit, tests/test_oracle_vs_ref.py::test_multiply_plain_monomial_branch).
# Zero weights are skipped (SEAL refuses a zero plaintext: "result ciphertext is transparent"). Row k's result sits in slot n-1.
# Everything runs through the engine's existing entry points (hhe_square, hhe_relinearize, hhe_multiply_plain, hhe_negate, hhe_add).
def square_activation(ctx: Context, hidden):
    """relinearize(square(h)) for a batch of hidden-neuron ciphertexts [H][2][L][N]"""
    return ctx.relinearize(ctx.square(np.asarray(hidden, dtype=np.uint64)))


def fc2_plain_rows(ctx: Context, squared, W2):
    """out[k] = sum_j W2[k][j] * squared[j] with small integer weights (any sign); all nonzero products of a row in one
    batched multiply_plain, then summed."""
    squared = np.asarray(squared, dtype=np.uint64)
    W2 = np.asarray(W2, dtype=np.int64)
    if W2.ndim != 2 or W2.shape[1] != squared.shape[0]:
        raise HheInvalidArgument(-1, "fc2 weights must be [rows][hidden]")
    out = []
    for k in range(W2.shape[0]):
        js = [j for j in range(W2.shape[1]) if abs(int(W2[k, j])) % ctx.t]
        if not js:
            raise HheInvalidArgument(-1, "fc2 row has no nonzero weight: the result would be a transparent ciphertext")
        pts = np.zeros((len(js), ctx.N), dtype=np.uint64)
        pts[:, 0] = [abs(int(W2[k, j])) % ctx.t for j in js]
        terms = ctx.multiply_plain(squared[js], pts)
        acc = None
        for j, term in zip(js, terms):
            if W2[k, j] < 0:
                term = ctx.negate(term)
            acc = term if acc is None else ctx.add(acc, term)
        out.append(acc)
    return np.stack(out)


def evaluate_model_2fc(ctx: Context, decomposed, enc_w1, input_len, W2, keyset=KEYSET_1):
    """fc1 (evaluate_model) -> square activation -> fc2 for a list of decomposed records: out[sample][row2]."""
    h = evaluate_model(ctx, decomposed, enc_w1, input_len, keyset)  # [samples][H]
    return np.stack([fc2_plain_rows(ctx, square_activation(ctx, h[s]), W2) for s in range(h.shape[0])])
