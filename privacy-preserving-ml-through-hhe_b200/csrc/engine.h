// Engine: device-resident BFV evaluator for the PASTA-3 transciphering + encrypted-FC path.
// All pointer arguments of the batched methods are DEVICE pointers; `items` ciphertexts are laid out back to back
// in SEAL's layout. Work is enqueued on the context's stream; nothing here synchronises with the host.
#pragma once
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "devconsts.h"
#include "kernels.h"
#include "launch.h"
#include "params.h"

namespace hhe {

class Engine {
 public:
  Engine(const Params &p, int device, void *stream);
  ~Engine();
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;

  const Params &params() const { return P_; }
  Device &dev() { return dev_; }
  size_t ct_words(int size = 2) const { return static_cast<size_t>(size) * P_.L * P_.N; }
  int batch_limit() const { return batch_; }
  void set_batch(int b) { batch_ = b; }
  int device() const { return device_; }
  // Default lock-step batch: two blocks per SM (whole waves of the half-limb kernels), reduced so that the decomposition's
  // working set (about 45 MiB per block at N = 16384 in BSGS mode, the larger of the two) fits the free HBM with a margin.
  int auto_batch();

  // ---- scratch arena (stack discipline, stream ordered) ----
  struct Scope {
    Engine &e;
    size_t chunk, used;
    explicit Scope(Engine &eng) : e(eng), chunk(eng.cur_chunk_), used(eng.chunks_.empty() ? 0 : eng.chunks_[eng.cur_chunk_].used) {}
    ~Scope() { e.arena_restore(chunk, used); }
  };
  u64 *scratch(size_t words);

  // ---- keys ----
  void load_ksk(int kind, u32 elt, const u64 *host_ksk);
  const W2 *find_key(int kind, u32 elt) const;
  void clear_keyset(int kind);  // drops every key of one kind (a new seal::GaloisKeys object replaces the previous one)

  // ---- primitives (device pointers) ----
  void ntt(const u64 *in, u64 *out, size_t items, int limbs, const TabMap &map, bool inverse, size_t item_stride = 0,
           size_t limb_stride = 0);
  void add(const u64 *a, const u64 *b, u64 *out, size_t items, int size = 2);
  void negate(const u64 *a, u64 *out, size_t items);
  void add_plain(const u64 *a, const u64 *pt, size_t pstride, u64 *out, size_t items, bool negate_first, const u32 *ptidx = nullptr,
                 const u32 *aidx = nullptr);
  void broadcast(const u64 *src, u64 *out, size_t words, size_t items);
  void encode_slots(const u64 *slots, size_t sstride, const u32 *lens, u32 n, u64 *pt, size_t items);
  // ndiag > 1: `ndiag` consecutive diagonals diag, diag + 1, ... in one launch, pt laid out [ndiag][items][N]
  void encode_material(const u32 *material, const u32 *mat_index, int mode, int layer, int diag, u64 *pt, size_t items, int ndiag = 1);
  // nolift (device, per plaintext, optional): 1 = monomial plaintext, multiplied without the centred lift as SEAL does
  void lift_ntt(const u64 *pt, u64 *D, size_t items, const u32 *nolift = nullptr);
  void ntt_mac(const u64 *ct, const u64 *D, size_t dstride, u64 *sum, size_t items, bool first, int comps = 2, size_t sum_off = 0,
               u64 *ntt_out = nullptr, const u32 *didx = nullptr);
  void ct_intt(u64 *ct, size_t items, int size = 2);
  void multiply_plain(const u64 *a, const u64 *pt, size_t pstride, u64 *out, size_t items, const u32 *nolift = nullptr);
  void galois(const u64 *a, u32 elt, u64 *out, size_t items);
  // out[c] = ModDown(sum_J NTT(target_J) * key) + base_c ; target/base given with item strides (words)
  // accum (optional, [items][2][L][N], may alias out): out = accum + result
  void key_switch(const u64 *target, size_t tstride, const W2 *key, const u64 *base0, const u64 *base1, size_t bstride,
                  u64 *out, size_t items, const u64 *accum = nullptr);
  void apply_galois(const u64 *a, u32 elt, const W2 *key, u64 *out, size_t items, const u64 *accum = nullptr);
  void rotate_rows_add(const u64 *a, int steps, int keyset, u64 *acc, size_t items);  // acc += rotate_rows(a, steps)
  void rotate_rows(const u64 *a, int steps, int keyset, u64 *out, size_t items);
  void rotate_columns(const u64 *a, int keyset, u64 *out, size_t items);
  void relinearize(const u64 *a3, u64 *out, size_t items);
  void multiply(const u64 *a, const u64 *b, u64 *out3, size_t items);
  void exponentiate3(const u64 *a, u64 *out, size_t items);

  // ---- hot path ----
  void material(const u64 *d_counters, size_t nblocks, u64 nonce, u32 *d_out);
  void pasta_decompose(const u64 *d_enc_key, const u64 *d_sym, const u32 *d_lens, const std::vector<u64> &counters,
                       u64 nonce, bool use_bsgs, u64 *d_out);
  // repeated counters: the keystream ciphertext once per distinct counter, then what is each block's own (engine_pasta.cu)
  bool share_keystreams() const;
  void pasta_keystreams(const u64 *d_enc_key, const std::vector<u64> &counters, u64 nonce, bool use_bsgs, u64 *d_ks);
  void pasta_finish(const u64 *d_ks, const u32 *d_idx, const u64 *d_sym, const u32 *d_lens, size_t nblocks, u64 *d_out);
  void pasta_check_keys();
  // plain PASTA-3 (PASTA::encrypt / decrypt): n_words words, block b uses counter first_counter + b
  void pasta_plain(const u64 *d_key256, const u64 *d_in, size_t n_words, u64 nonce, u64 first_counter, bool decrypt, u64 *d_out);
  void mask(const u64 *a, const u64 *d_mask_slots, u32 n, u64 *out, size_t items);
  void flatten(const u64 *in, size_t per, int keyset, u64 *out, size_t items);
  void vec_sum(const u64 *a, size_t n, int keyset, u64 *out, size_t items);
  void strided_copy(const u64 *src, size_t sstride, u64 *dst, size_t dstride, size_t words, size_t rows);
  // seal::Encryptor::encrypt (public key, BFV) with SEAL's Blake2xb generator seeded per ciphertext (engine_enc.cu)
  void encrypt(const u64 *d_pk, const u64 *d_seeds, const u64 *d_pt, size_t count, u64 *d_out);

  const DevConsts *dconsts() const { return dC_; }
  TwRef twref() const { return TwRef{dTw_, P_.N, f64_gmin_}; }
  const u32 *index_map() const { return dIndex_; }

 private:
  friend struct Scope;
  void arena_restore(size_t chunk, size_t used);
  // d_counters: the `nd` distinct SHAKE counters of the batch; didx (device, nb entries, null when nd == nb): block -> index
  // of its counter. Blocks with equal counters share round material, encoded diagonals and their lifted transforms.
  void pasta_batch(const u64 *d_enc_key, const u64 *d_sym, const u32 *d_lens, const u64 *d_counters, size_t nb, size_t nd,
                   const u32 *didx, u64 nonce, bool use_bsgs, u64 *d_out);
  void affine_diagonal(u64 *state, const u32 *mat, int layer, size_t nb, size_t nd, const u32 *didx);
  void affine_diagonal_resident(u64 *state, const u32 *mat, int layer, size_t nb, size_t nd, const u32 *didx);
  const u32 *ntt_perm(u32 elt);                                      // NTT-slot permutation of a Galois element (gather form)
  const u32 *ntt_perm_inv(u32 elt) { return ntt_perm(elt) + P_.N; }  // its inverse (scatter form)
  void affine_bsgs(u64 *state, const u32 *mat, int layer, size_t nb, size_t nd, const u32 *didx);
  void feistel(u64 *state, size_t nb);
  const W2 *need_key(int kind, u32 elt) const;
  const u64 *feistel_mask_ntt();

  Params P_;
  Device dev_;
  int device_ = 0;
  int batch_ = 0;
  int f64_gmin_ = 1;
  bool split_ = false;         // N = 32768 code path (split transforms)
  bool compact_keys_ = false;
  bool tmem_ks_ = false;  // FP64 key switch with accumulators in tensor memory (keys stored group-major)
  size_t half_smem(int logh) const { return ntt_smem_words(1 << logh) * 8; }  // dynamic shared memory of a half-limb kernel
  int ks_split_max_ = 3;   // up to this many items (measured: 1-3 blocks faster, 4 slower than the standard kernel) a key switch runs as eight-CTA clusters, one digit per CTA (HHE_KS_SPLIT_MAX)
  int ks_threads_ = 512;   // CTA size of the tensor-memory key-switch kernel (512 x 64 registers or 256 x 128 registers)
  bool cluster_inv_ = false;  // FP64 inverse transforms as two-CTA clusters (half-limb CTAs, last stage over distributed shared memory)
  int pf_ntt_ = 0, pf_limbs_ = 0, pf_items_ = 0;  // L2 prefetch distances (limbs: plain transforms / other half-limb kernels; items: ks_digits); 0 = off
  bool half_fwd_ = false;  // FP64 forward transforms of lift_ntt / ntt_mac / corr0_mac as half-limb CTAs (two per SM)
  // reuse / perm: the NTT form of the source polynomial and the slot permutation of the Galois element (digit J on key limb J
  // needs no transform); perm == nullptr with reuse != nullptr: `reuse` is already stored permuted (read linearly)
  void launch_ks_digits(const u64 *target, size_t tstride, const W2 *key, u64 *acc, size_t items, const u64 *reuse, size_t reuse_stride,
                        const u32 *perm);  // every key limb is on the FP64 path: keys are stored as doubles (8 bytes per residue)
  DevConsts *dC_ = nullptr;
  W2 *dTw_ = nullptr;
  u32 *dIndex_ = nullptr;
  u64 *dFeistel_ = nullptr;  // cached NTT of the lifted Feistel mask, [L][N]
  std::map<std::pair<int, u32>, W2 *> keys_;
  std::map<u32, u32 *> perms_;  // Galois element -> permutation of NTT slots (device)
  struct Chunk {
    u64 *ptr;
    size_t words, used;
  };
  std::vector<Chunk> chunks_;
  size_t cur_chunk_ = 0;
};

}  // namespace hhe
