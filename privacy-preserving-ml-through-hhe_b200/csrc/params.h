// Host-side derivation of every constant table the engine needs for a BFV parameter set
// (N, t, q_0..q_{K-1}). Everything SEAL computes when it builds a SEALContext for the same parameters is
// re-derived here from first principles so that results are limb-exact:
//   * NTT roots: the numerically smallest primitive 2N-th root per modulus (seal/util/ntt.h:69-93)
//   * BatchEncoder slot -> coefficient index map (seal/batchencoder.h:80-134)
//   * add_plain scaling constants floor(Q/t) mod q_i, Q mod t (seal/util/scalingvariant.h)
//   * key-switch ModDown constants (seal/evaluator.h:1260)
//   * BEHZ auxiliary base and conversion matrices (seal/util/rns.h:190-400)
#pragma once
#include <cstdint>
#include <vector>

namespace hhe {

using u64 = uint64_t;
using u128 = unsigned __int128;

constexpr int kMaxLimbs = 18;  // K <= 17 (N = 32768 has K = 16), Bsk has L + 1

struct Twiddle {  // (w, floor(w * 2^64 / q)) -- one 128-bit load on the device
  u64 w, ws;
};

struct NttTable {
  u64 q = 0, psi = 0;
  std::vector<Twiddle> fwd;  // fwd[k] = psi^bitrev(k)
  std::vector<Twiddle> inv;  // inv[k] = psi^-bitrev(k)
  Twiddle n_inv{};           // N^-1
};

struct Params {
  u64 N = 0, t = 0;
  int logn = 0, K = 0, L = 0;
  std::vector<u64> q;         // K primes, special last
  std::vector<NttTable> tab;  // [0,K): q_i ; [K, K+L+1): Bsk ; [2K]: t   (index = table id)
  int tab_bsk(int p) const { return K + p; }
  int tab_plain() const { return 2 * K; }
  std::vector<uint32_t> index_map;  // N entries

  // add_plain / multiply_plain
  u64 q_mod_t = 0, half_t = 0;         // Q mod t, (t+1)/2
  u64 q_div_t_mod_q[kMaxLimbs] = {};   // floor(Q/t) mod q_i
  // key switching
  u64 half_sp = 0;                      // floor(q_sp / 2)
  u64 half_sp_mod_q[kMaxLimbs] = {};
  Twiddle inv_sp_mod_q[kMaxLimbs] = {};  // q_sp^-1 mod q_i (Shoup pair)
  // BEHZ
  u64 m_sk = 0, gamma = 0, m_tilde = 0;
  u64 bsk[kMaxLimbs] = {};              // base_B (L) then m_sk
  Twiddle mtilde_mod_q[kMaxLimbs] = {};      // m_tilde * (Q/q_i)^-1 mod q_i   (fused first two scalings)
  Twiddle inv_punct_q[kMaxLimbs] = {};       // (Q/q_i)^-1 mod q_i
  Twiddle t_inv_punct_q[kMaxLimbs] = {};     // t * (Q/q_i)^-1 mod q_i
  u64 q2bsk[kMaxLimbs][kMaxLimbs] = {};      // [p][i] = (Q/q_i) mod bsk_p
  uint32_t q2mt[kMaxLimbs] = {};             // (Q/q_i) mod 2^32
  uint32_t neg_inv_q_mt = 0;                 // -Q^-1 mod 2^32
  u64 q_mod_bsk[kMaxLimbs] = {};
  Twiddle inv_mt_bsk[kMaxLimbs] = {};        // m_tilde^-1 mod bsk_p
  Twiddle t_mod_bsk[kMaxLimbs] = {};         // t mod bsk_p
  Twiddle inv_q_bsk[kMaxLimbs] = {};         // Q^-1 mod bsk_p
  Twiddle inv_punct_b[kMaxLimbs] = {};       // (P_B/b_i)^-1 mod b_i
  u64 b2q[kMaxLimbs][kMaxLimbs] = {};        // [j][i] = (P_B/b_i) mod q_j
  u64 b2msk[kMaxLimbs] = {};                 // (P_B/b_i) mod m_sk
  Twiddle inv_pb_msk{};                      // P_B^-1 mod m_sk
  u64 pb_mod_q[kMaxLimbs] = {};              // P_B mod q_j

  // Throws std::invalid_argument on unusable parameters.
  static Params derive(u64 N, u64 t, const u64 *q, int nq);
  uint32_t galois_elt_from_step(int step) const;  // 0 if the step is out of range
};

// SEAL's non-adjacent-form decomposition of a rotation step (seal/util/numth.h:22-42), emission order kept.
std::vector<int> naf_steps(int value);

u64 mul_mod(u64 a, u64 b, u64 q);
u64 pow_mod(u64 a, u64 e, u64 q);
u64 inv_mod_prime(u64 a, u64 q);
Twiddle shoup_pair(u64 w, u64 q);

}  // namespace hhe
