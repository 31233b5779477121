/* include/hhe_b200.h -- C ABI of libhhe_b200.so, the B200-native (sm_100a) BFV engine for the reference's
 * PASTA-3 transciphering + encrypted-FC hot path.
 *
 * Every entry point replaces one reference interface (paths relative to the reference repository):
 *   hhe_ctx_create            pasta::SEALZpCipher::create_context          src/pasta/SEAL_Cipher.cpp:38-68
 *                             sealhelper::get_seal_context                 src/util/sealhelper.cpp:8-41
 *   hhe_load_ksk              PASTA_SEAL ctor's rk/gk copies               src/pasta/pasta_3_seal.h:11-15, src/examples/CSP/CSP.cpp:238-242
 *   hhe_pasta3_decompose      pasta::PASTA_SEAL::decomposition/HE_decrypt  src/pasta/pasta_3_seal.cpp:42-172
 *   hhe_mask                  pasta::SEALZpCipher::mask                    src/pasta/SEAL_Cipher.cpp:161-166
 *   hhe_flatten               pasta::SEALZpCipher::flatten                 src/pasta/SEAL_Cipher.cpp:170-181
 *   hhe_multiply              sealhelper::packed_enc_multiply              src/util/sealhelper.cpp:268-274
 *   hhe_relinearize           seal::Evaluator::relinearize_inplace         src/examples/CSP/CSP.cpp:306
 *   hhe_vec_sum               sealhelper::encrypted_vec_sum                src/util/sealhelper.cpp:379-392
 *   hhe_fc_rows               the per-output-neuron loop                   src/examples/hhe_pktnn_examples.cpp:957-992, CSP.cpp:288-323
 *   hhe_rotate_rows/columns, hhe_add, hhe_negate, hhe_add_plain, hhe_multiply_plain, hhe_square, hhe_encode,
 *   hhe_exponentiate3         seal::Evaluator / seal::BatchEncoder members reached by the above
 *                             libs/seal/include/SEAL-4.0/seal/evaluator.h:92,118,234,261,301,621,665,729,955,1060; batchencoder.h:80
 *   hhe_ntt                   seal::util::ntt_negacyclic_harvey / inverse  libs/seal/include/SEAL-4.0/seal/util/ntt.h:195-340
 *
 * Data layouts are SEAL 4.0's, host memory, caller-owned, little-endian uint64:
 *   ciphertext   u64[size][L][N]   coefficient (non-NTT) form, residues canonical in [0, q_i), size 2 (3 after multiply)
 *   plaintext    u64[N]            coefficients in [0, t)
 *   ksk          u64[L][2][K][N]   = KSwitchKeys::data()[(elt-1)/2][J].data(), NTT form, K = L+1 key-level limbs
 * Batched calls take `count` items laid out back to back.
 *
 * Errors: every function returns HHE_OK or a negative status; hhe_last_error() gives the message for the calling
 * thread. HHE_ERR_INVALID maps to the std::invalid_argument SEAL/PASTA would throw (missing Galois key, bad sizes),
 * HHE_ERR_RUNTIME to std::runtime_error / CUDA failures. There is NO CPU fallback: without a usable CUDA device
 * hhe_ctx_create fails with HHE_ERR_NO_DEVICE.
 *
 * Threading: one host thread at a time per context (each context owns its CUDA streams); contexts are independent, also
 * across devices: every entry point makes its context's device current for the call and restores the caller's afterwards.
 * Host-buffer calls are synchronous with respect to the host (results are in `out` on return); the hhe_dev_* calls only enqueue
 * work on the context's stream (hhe_sync waits for it).
 */
#ifndef HHE_B200_H
#define HHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hhe_ctx hhe_ctx;

enum {
  HHE_OK = 0,
  HHE_ERR_INVALID = -1,   /* std::invalid_argument in the reference */
  HHE_ERR_RUNTIME = -2,   /* std::runtime_error / CUDA error */
  HHE_ERR_NO_DEVICE = -3, /* no CUDA device / wrong architecture: there is no CPU path */
  HHE_ERR_LOGIC = -4      /* std::logic_error in the reference */
};

/* Key kinds for hhe_load_ksk / keyset arguments. Galois keysets mirror the two seal::GaloisKeys objects the
 * reference juggles (CSP-specific set for PASTA/flatten, analyst default power-of-two set for encrypted_vec_sum). */
enum { HHE_KEYSET_0 = 0, HHE_KEYSET_1 = 1, HHE_RELIN = 2 };

const char *hhe_last_error(void);
const char *hhe_version(void);
/* 1 for the product (compiled by nvcc for sm_100a). 0 only for the host emulation of the kernel bodies that the repository's
 * CPU-tier tests build (tests/emul): bindings must refuse such a library outside those tests -- there is no CPU path. */
int hhe_build_is_cuda(void);

/* q[0..nq): coefficient-modulus primes, last = special prime (as CoeffModulus::BFVDefault returns them).
 * device: CUDA ordinal. stream: a cudaStream_t to run on (NULL: the context creates its own). */
int hhe_ctx_create(hhe_ctx **out, uint64_t N, uint64_t t, const uint64_t *q, int nq, int device, void *stream);
void hhe_ctx_destroy(hhe_ctx *ctx);
/* info[0..7) = N, L, K, t, max batch (blocks processed per lock-step wave), SM count,
 * number of moduli served by the FP64-pipe kernels (q <= 2^49; the others use the integer Shoup kernels) */
int hhe_ctx_info(const hhe_ctx *ctx, uint64_t *info);
void *hhe_ctx_stream(const hhe_ctx *ctx);
/* Blocks per lock-step batch inside decompose. 0 = automatic: two blocks per SM, reduced if the batch's working set (about
 * 60 MiB per block at N = 16384) would not fit 80 % of the free HBM (cudaMemGetInfo at the time of the call). */
int hhe_set_batch(hhe_ctx *ctx, int blocks);
uint32_t hhe_galois_elt(const hhe_ctx *ctx, int step);
/* Constants the engine derived (for parity checks against SEAL): psi[K] then psi_t; then m_sk, gamma, m_tilde,
 * base_B[L], psi_Bsk[L+1].  out has K + 1 + 3 + 2L + 1 words. */
int hhe_ctx_constants(const hhe_ctx *ctx, uint64_t *out);

int hhe_load_ksk(hhe_ctx *ctx, int kind, uint32_t galois_elt, const uint64_t *ksk);
int hhe_has_ksk(const hhe_ctx *ctx, int kind, uint32_t galois_elt);
/* Drops every key of one kind: a different seal::GaloisKeys object replaces the previous one as a whole, so a rotation
 * whose key the new object lacks fails with HHE_ERR_INVALID ("Galois key not present") as it does in SEAL. */
int hhe_clear_keyset(hhe_ctx *ctx, int kind);

/* ---- SEAL-level primitives, batched over `count` ciphertexts ---- */
int hhe_ntt(hhe_ctx *ctx, int limb, int inverse, uint64_t *data, size_t count); /* limb<K: q_limb; >=K: Bsk[limb-K] */
int hhe_encode(hhe_ctx *ctx, const uint64_t *slots, size_t n_slots, uint64_t *pt, size_t count);
int hhe_add(hhe_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count);
int hhe_negate(hhe_ctx *ctx, const uint64_t *a, uint64_t *out, size_t count);
int hhe_add_plain(hhe_ctx *ctx, const uint64_t *a, const uint64_t *pt, uint64_t *out, size_t count);
int hhe_multiply_plain(hhe_ctx *ctx, const uint64_t *a, const uint64_t *pt, uint64_t *out, size_t count);
int hhe_rotate_rows(hhe_ctx *ctx, const uint64_t *a, int steps, int keyset, uint64_t *out, size_t count);
int hhe_rotate_columns(hhe_ctx *ctx, const uint64_t *a, int keyset, uint64_t *out, size_t count);
int hhe_multiply(hhe_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out3, size_t count);
int hhe_square(hhe_ctx *ctx, const uint64_t *a, uint64_t *out3, size_t count);
int hhe_relinearize(hhe_ctx *ctx, const uint64_t *a3, uint64_t *out, size_t count);
int hhe_exponentiate3(hhe_ctx *ctx, const uint64_t *a, uint64_t *out, size_t count);

/* ---- the hot path ---- */
/* Transcipher n_words symmetric-ciphertext words (ceil(n/128) blocks). Block b uses SHAKE counter first_counter+b
 * (the reference: nonce 123456789, first_counter 0). enc_key: the HE-encrypted symmetric key (one size-2 ct).
 * out: ceil(n/128) size-2 ciphertexts. Galois keys are taken from HHE_KEYSET_0, relin key from HHE_RELIN. */
int hhe_pasta3_decompose(hhe_ctx *ctx, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n_words, uint64_t nonce,
                         uint64_t first_counter, int use_bsgs, uint64_t *out);
/* Same, for `records` independent records of n_words each that all restart at first_counter (CSP.cpp:247-252). The keystream
 * ciphertext of a counter does not depend on the data and the evaluation is deterministic, so it is evaluated once per distinct
 * counter per call and shared by the records (bit-identical to per-record evaluation; HHE_NO_SHARED_KEYSTREAM=1 turns it off). */
int hhe_pasta3_decompose_records(hhe_ctx *ctx, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n_words,
                                 size_t records, uint64_t nonce, uint64_t first_counter, int use_bsgs, uint64_t *out);
int hhe_mask(hhe_ctx *ctx, const uint64_t *a, const uint64_t *mask, size_t n_mask, uint64_t *out, size_t count);
/* in: `count` groups of `per` ciphertexts -> `count` ciphertexts */
int hhe_flatten(hhe_ctx *ctx, const uint64_t *in, size_t per, int keyset, uint64_t *out, size_t count);
int hhe_vec_sum(hhe_ctx *ctx, const uint64_t *a, size_t n, int keyset, uint64_t *out, size_t count);
/* Encrypted FC: for each of `samples` inputs x and `rows` encrypted weight rows w: vec_sum(relin(x*w), n).
 * out[samples][rows] ciphertexts; slot n-1 holds the dot product. */
int hhe_fc_rows(hhe_ctx *ctx, const uint64_t *x, size_t samples, const uint64_t *w, size_t rows, size_t n, int keyset,
                uint64_t *out);

/* ---- service-level calls (the CSP's request handlers) ---- */
/* BaseCSP::decompose (src/examples/CSP/CSP.cpp:235-283): per record decomposition (counters restart at 0) -> optional
 * mask of the last block -> flatten (rotations by -128*i with `flatten_keyset`). out: one ciphertext per record.
 * apply_mask = 0: the reference service's behaviour (it masks a copy, CSP.cpp:262-269); 1: masks in place like
 * src/examples/hhe_pktnn_examples.cpp:620-624. Intermediate ciphertexts stay in HBM. */
int hhe_csp_decompose(hhe_ctx *ctx, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n_words, size_t records, uint64_t nonce,
                      int use_bsgs, int apply_mask, int flatten_keyset, uint64_t *out);
/* CSP_hhe_pktnn_1fc::evaluateModel (src/examples/CSP/CSP.cpp:288-323): out[records][rows] =
 * vec_sum(relinearize(record * weight_row), input_len); the dot product lands in slot input_len-1. */
int hhe_csp_evaluate_model(hhe_ctx *ctx, const uint64_t *records_ct, size_t records, const uint64_t *enc_weights, size_t rows,
                           size_t input_len, int sum_keyset, uint64_t *out);

/* ---- SEAL 4.0 wire format at the boundary ----
 * The reference moves ciphertexts and keys as seal::Ciphertext / GaloisKeys / RelinKeys ::save byte streams: gRPC payloads
 * (src/examples/CSP/CSP.cpp:131,201; src/examples/Analyst/Analyst.cpp:258-341; src/examples/User/User.cpp:83,168) and the
 * decomposition checkpoint file (CSP.cpp:495-547,575-595: size_t count, then `count` saved ciphertexts). These entry points
 * read/write that format (libs/seal/include/SEAL-4.0/seal/serialization.h:49-91; ciphertext.h:466-642; kswitchkeys.h:186-300)
 * straight from/to the engine's buffers. Host-side byte work: the stateless ones need no device.
 * compr_mode: 0 none, 1 zlib, 2 zstd (SEAL's default). Errors: HHE_ERR_LOGIC for invalid data (SEAL: std::logic_error),
 * HHE_ERR_INVALID for bad arguments, HHE_ERR_RUNTIME when libzstd.so.1 is missing for mode 2. */
typedef struct hhe_seal_ring {
  uint64_t N, t;
  const uint64_t *q; /* key-level primes, last = special prime */
  int nq;
} hhe_seal_ring;
/* EncryptionParameters::parms_id(): level 0 = SEALContext::first_parms_id(), 1 = key_parms_id() */
int hhe_seal_parms_id(const hhe_seal_ring *ring, int level, uint64_t out[4]);
size_t hhe_seal_ct_save_bound(const hhe_seal_ring *ring, int size);
/* Ciphertext::save: ct = u64[size][L][N] (coefficient form, first_parms_id) -> bytes */
int hhe_seal_ct_save(const hhe_seal_ring *ring, const uint64_t *ct, int size, int compr_mode, uint8_t *out, size_t cap,
                     size_t *written);
/* Ciphertext::load(context, ...): validates like SEAL (header, parms_id, dimensions, residues < q_i) */
int hhe_seal_ct_load(const hhe_seal_ring *ring, const uint8_t *in, size_t len, uint64_t *ct, size_t cap_words, int *size,
                     size_t *consumed);
/* GaloisKeys / RelinKeys::load: unpacks every key present into ksk[n][L][2][K][N] (the hhe_load_ksk layout) and its index
 * ((galois_elt - 1) / 2; 0 for the relinearisation key). ksk == NULL only counts. */
int hhe_seal_keys_unpack(const hhe_seal_ring *ring, const uint8_t *in, size_t len, uint64_t *index, uint64_t *ksk, size_t cap_keys,
                         size_t *n_keys, size_t *consumed);
/* Same walk, every key uploaded to the device as it is parsed: kind = HHE_KEYSET_0/1 for a GaloisKeys stream, HHE_RELIN for
 * a RelinKeys stream (replaces the by-value key copies of the PASTA_SEAL ctor, CSP.cpp:238-242). */
int hhe_load_seal_keys(hhe_ctx *ctx, int kind, const uint8_t *in, size_t len, size_t *n_keys, size_t *consumed);
/* hhe_pasta3_decompose with the serialized forms on both sides: enc_key_bytes = User.cpp:168 payload, out = `nblocks`
 * Ciphertext::save objects back to back (sizes in out_sizes[nblocks]); what BaseCSP::decompose + the gRPC reply do. */
int hhe_pasta3_decompose_serialized(hhe_ctx *ctx, const uint8_t *enc_key_bytes, size_t enc_key_len, const uint64_t *sym_ct,
                                    size_t n_words, uint64_t nonce, uint64_t first_counter, int use_bsgs, int compr_mode, uint8_t *out,
                                    size_t cap, size_t *out_sizes, size_t *written);

/* ---- device-resident variants (inputs/outputs already in HBM; used for kernel-only timing and pipelines) ---- */
int hhe_dev_alloc(hhe_ctx *ctx, size_t bytes, void **dptr);
int hhe_dev_free(hhe_ctx *ctx, void *dptr);
int hhe_dev_upload(hhe_ctx *ctx, void *dptr, const void *host, size_t bytes);
int hhe_dev_download(hhe_ctx *ctx, void *host, const void *dptr, size_t bytes);
int hhe_sync(hhe_ctx *ctx);
int hhe_dev_ntt(hhe_ctx *ctx, int limb, int inverse, uint64_t *d_data, size_t count);
int hhe_dev_rotate_rows(hhe_ctx *ctx, const uint64_t *d_a, int steps, int keyset, uint64_t *d_out, size_t count);
int hhe_dev_relinearize(hhe_ctx *ctx, const uint64_t *d_a3, uint64_t *d_out, size_t count);
int hhe_dev_multiply(hhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out3, size_t count);
/* counters[nblocks]: SHAKE counter per block; d_sym: u64[nblocks][128] (short blocks zero-padded, lens[] words each) */
int hhe_dev_pasta3_decompose(hhe_ctx *ctx, const uint64_t *d_enc_key, const uint64_t *d_sym, const uint32_t *lens,
                             const uint64_t *counters, size_t nblocks, uint64_t nonce, int use_bsgs, uint64_t *d_out);
/* Kernel launches issued by this context since creation (for bench.py's gpu_launches). */
uint64_t hhe_launch_count(const hhe_ctx *ctx);
/* Per-kernel device time: when enabled every launch is bracketed by CUDA events on the context's stream.
 * hhe_profile_report writes a JSON object {"kernel": {"launches": n, "ms": t}, ...} into buf (NUL-terminated). */
int hhe_profile_enable(hhe_ctx *ctx, int on);
int hhe_profile_reset(hhe_ctx *ctx);
int hhe_profile_report(hhe_ctx *ctx, char *buf, size_t cap);

/* ---- client side: plain (non-homomorphic) PASTA-3 for bulk data owners ----
 * pasta::PASTA::encrypt / decrypt (src/pasta/pasta_3_plain.cpp:9-47; keystream :156-171, round functions :198-282):
 * out[i] = (in[i] + ks) mod t, or in[i] - ks (+ t if negative) for decrypt, exactly as the reference (inputs are not
 * reduced on decrypt). key256: the 256-word symmetric key. Block b uses SHAKE counter first_counter + b (reference:
 * nonce 123456789, counters restart at 0 for every call). The modulus is the context's plain modulus t. */
int hhe_pasta3_plain(hhe_ctx *ctx, const uint64_t *key256, const uint64_t *in, size_t n_words, uint64_t nonce, uint64_t first_counter,
                     int decrypt, uint64_t *out);

/* seal::Encryptor::encrypt with a public key (BFV; libs/seal/include/SEAL-4.0/seal/encryptor.h:132, seal/util/rlwe.h:52-105) for
 * bulk data owners and analysts: the operation behind sealhelper::encrypt_weight_mat (src/util/sealhelper.cpp:123-142) and
 * pastahelper::encrypt_symmetric_key (src/util/pastahelper.cpp:355-377). pk = seal::PublicKey::data(): u64[2][K][N], NTT form.
 * seeds: u64[count][8], one seal::prng_seed_type per ciphertext for SEAL's Blake2xb generator (NULL: 64 bytes per ciphertext
 * from the operating system, as SEAL does). For the same seed the ciphertext is bit-identical to SEAL's: the Blake2xb stream, the
 * ternary and centred-binomial samplers and their draw order are reproduced on the GPU. plain: u64[count][N] coefficients < t.
 * hhe_encrypt_slots = BatchEncoder::encode + encrypt (slots: u64[count][n_slots], values < t). */
int hhe_encrypt(hhe_ctx *ctx, const uint64_t *pk, const uint64_t *seeds, const uint64_t *plain, size_t count, uint64_t *out);
int hhe_encrypt_slots(hhe_ctx *ctx, const uint64_t *pk, const uint64_t *seeds, const uint64_t *slots, size_t n_slots, size_t count,
                      uint64_t *out);

/* ---- plain PASTA-3 material (device SHAKE128 + matrix generation, for parity tests of that kernel) ---- */
/* mat1[128*128], mat2[128*128], rc[256] as u32 for (nonce, counter, layer 0..3) */
int hhe_pasta_layer_material(hhe_ctx *ctx, uint64_t nonce, uint64_t counter, int layer, uint32_t *mat1, uint32_t *mat2,
                             uint32_t *rc);

#ifdef __cplusplus
}
#endif
#endif
