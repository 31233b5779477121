/* oracle/hhe_oracle.h -- CPU restatement of the reference hot path. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain C restatement of (a) the SEAL 4.0 BFV evaluation primitives the reference reaches through
 * seal::Evaluator / seal::BatchEncoder (SEAL's .cpp sources are NOT in /root/reference: only headers
 * under libs/seal/include/SEAL-4.0 and the prebuilt libs/seal/lib/libseal-4.0.a, pinned version 4.0.0;
 * the algorithms below restate its published behaviour) and (b) the reference's
 * own PASTA-3 code (src/pasta/pasta_3_plain.cpp, src/pasta/pasta_3_seal.cpp, src/pasta/SEAL_Cipher.cpp,
 * src/util/sealhelper.cpp) and XKCP SHAKE128 (libs/keccak).
 *
 * Pinning: the reference holds no golden vectors for this path, so this restatement
 * is pinned against the reference itself: oracle/_ref/libhhe_ref.so (the unmodified reference sources +
 * vendored libseal, built by oracle/Makefile) in tests/test_oracle_vs_ref.py, and against fixtures that
 * library generated, committed under tests/golden/ with tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * Layouts (SEAL's): ciphertext = u64[size][L][N], coefficient form, canonical residues;
 * plaintext = u64[N] coefficients in [0,t); key-switching key = u64[L digits][2][K][N], NTT form.
 */
#ifndef HHE_ORACLE_H
#define HHE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hor_ctx hor_ctx;

/* q[0..K) = coefficient-modulus primes, the last one being the special (key-switching) prime. */
hor_ctx *hor_create(uint64_t N, uint64_t t, const uint64_t *q, int K);
void hor_destroy(hor_ctx *c);

/* out[0..K) = psi of every q, out[K] = psi of t */
void hor_ntt_roots(const hor_ctx *c, uint64_t *out);
/* out = m_sk, gamma, m_tilde, base_B[L], psi of Bsk[L+1]   (3 + 2L + 1 words) */
void hor_behz(const hor_ctx *c, uint64_t *out);
uint32_t hor_galois_elt(const hor_ctx *c, int step);
/* NAF of a rotation step as SEAL emits it (util/numth.h:22-42); returns term count. */
int hor_naf(int value, int *terms);

/* kind 0/1: galois keyset 0/1; kind 2: relinearisation key. Data is copied. */
int hor_load_ksk(hor_ctx *c, int kind, uint32_t galois_elt, const uint64_t *data);

/* limb < K: coefficient primes; limb >= K: Bsk prime (limb-K). In place. */
void hor_ntt(const hor_ctx *c, int limb, int inverse, uint64_t *data);

void hor_encode(const hor_ctx *c, const uint64_t *slots, size_t n, uint64_t *pt);
void hor_add(const hor_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out);
void hor_negate(const hor_ctx *c, const uint64_t *a, uint64_t *out);
void hor_add_plain(const hor_ctx *c, const uint64_t *a, const uint64_t *pt, uint64_t *out);
void hor_multiply_plain(const hor_ctx *c, const uint64_t *a, const uint64_t *pt, uint64_t *out);
/* one Galois automorphism + key switch with keyset `ks`; -1 if key missing */
int hor_apply_galois(const hor_ctx *c, const uint64_t *a, uint32_t elt, int ks, uint64_t *out);
/* Evaluator::rotate_rows semantics (direct key if present, else NAF chain); -1 if impossible */
int hor_rotate_rows(const hor_ctx *c, const uint64_t *a, int steps, int ks, uint64_t *out);
int hor_rotate_columns(const hor_ctx *c, const uint64_t *a, int ks, uint64_t *out);
void hor_multiply(const hor_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out3);
int hor_relinearize(const hor_ctx *c, const uint64_t *a3, uint64_t *out);
int hor_exponentiate3(const hor_ctx *c, const uint64_t *a, uint64_t *out);

/* sealhelper::encrypted_vec_sum (src/util/sealhelper.cpp:379-392) */
int hor_vec_sum(const hor_ctx *c, const uint64_t *a, size_t n, int ks, uint64_t *out);
/* SEALZpCipher::mask / flatten (src/pasta/SEAL_Cipher.cpp:161-181) */
void hor_mask(const hor_ctx *c, const uint64_t *a, const uint64_t *mask, size_t n, uint64_t *out);
int hor_flatten(const hor_ctx *c, const uint64_t *cts, size_t count, int ks, uint64_t *out);

/* PASTA_SEAL::decomposition (src/pasta/pasta_3_seal.cpp:106-172) with galois keyset 0.
 * Block b uses SHAKE counter first_counter + b (the reference always starts at 0). */
int hor_pasta_decompose(const hor_ctx *c, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n, uint64_t nonce,
                        uint64_t first_counter, int use_bsgs, uint64_t *out);

/* ---- plain PASTA-3 / SHAKE128 (src/pasta/pasta_3_plain.cpp) ---- */
/* mat1[128*128], mat2[128*128], rc[256] of affine layer `layer` (0..3) of block (nonce, counter) */
void hor_pasta_layer_material(uint64_t p, uint64_t nonce, uint64_t counter, int layer, uint64_t *mat1, uint64_t *mat2,
                              uint64_t *rc);
void hor_pasta_keystream(const uint64_t *key256, uint64_t p, uint64_t nonce, uint64_t counter, uint64_t *ks128);
void hor_pasta_plain(const uint64_t *key256, uint64_t p, const uint64_t *in, size_t n, int decrypt, uint64_t *out);
void hor_shake128(const uint8_t *in, size_t inlen, uint8_t *out, size_t outlen);

#ifdef __cplusplus
}
#endif
#endif
