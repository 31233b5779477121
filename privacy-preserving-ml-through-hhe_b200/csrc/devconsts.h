// Per-context constant block, copied once to HBM and read (broadcast, L1/L2-resident) by every kernel.
#pragma once
#include "modarith.h"
#include "modarith_f64.h"
#include "params.h"

namespace hhe {

constexpr int kMaxTab = 2 * kMaxLimbs + 1;
constexpr int kPastaT = 128;

struct DevConsts {
  u64 N;
  int logn, K, L;
  u64 t, half_t, q_mod_t;
  DevMod mod[kMaxTab];  // by NTT table id: [0,K) q_i, [K,2K) Bsk, [2K] t
  W2 n_inv[kMaxTab];    // N^-1 per table
  // FP64-pipe path (modarith_f64.h): tables whose modulus is <= 2^49 keep twiddles / keys as (w, w/q) doubles
  unsigned char f64[kMaxTab];
  double qf[kMaxTab], qinvf[kMaxTab];
  D2 n_inv_f[kMaxTab];
  u64 q_div_t_mod_q[kMaxLimbs];
  u64 half_sp;
  u64 half_sp_mod_q[kMaxLimbs];
  W2 inv_sp_mod_q[kMaxLimbs];
  D2 inv_sp_f[kMaxLimbs];  // the same constant for the FP64 path
  // BEHZ
  W2 mtilde_ipq[kMaxLimbs];  // m_tilde * (Q/q_i)^-1 mod q_i
  W2 t_ipq[kMaxLimbs];       // t * (Q/q_i)^-1 mod q_i
  u64 q2bsk[kMaxLimbs][kMaxLimbs];
  u32 q2mt[kMaxLimbs];
  u32 neg_inv_q_mt;
  u64 q_mod_bsk[kMaxLimbs];
  W2 inv_mt_bsk[kMaxLimbs];
  W2 t_mod_bsk[kMaxLimbs];
  W2 inv_q_bsk[kMaxLimbs];
  W2 inv_punct_b[kMaxLimbs];
  u64 b2q[kMaxLimbs][kMaxLimbs];
  u64 b2msk[kMaxLimbs];
  W2 inv_pb_msk;
  u64 pb_mod_q[kMaxLimbs];
};

inline DevMod make_devmod(u64 q) {
  DevMod m;
  m.q = q;
  // floor(2^128 / q) = (floor(2^64/q) << 64) + floor(((2^64 mod q) << 64) / q), assembled with 128-bit division
  u128 top = (static_cast<u128>(1) << 64);
  u64 hi = static_cast<u64>(top / q);
  u64 r = static_cast<u64>(top % q);
  u64 lo = static_cast<u64>((static_cast<u128>(r) << 64) / q);
  m.cr0 = lo;
  m.cr1 = hi;
  return m;
}

inline W2 w2(const Twiddle &t) { return W2{t.w, t.ws}; }

// The FP64 path needs 8q <= 2^52 for EVERY coefficient prime (key-switching keys are then stored as doubles) and at
// most 8 key digits (the key inner product sums L terms of magnitude <= 1.5q in a double). The BEHZ auxiliary primes
// (61 bit) and the plaintext table (q = t, whose values feed integer-only kernels) stay on the integer path.
inline bool table_is_f64(const Params &p, int tab) {
#ifdef HHE_NO_F64
  (void)p; (void)tab;
  return false;
#else
  // the coefficient primes, and (round 2) the plain modulus: BatchEncoder::encode's inverse transform mod t then runs on the FP64
  // half-limb cluster kernels as well (t = 65537 is far below the 2^49 limit)
  if ((tab >= p.K && tab != p.tab_plain()) || p.L > 8) return false;
  for (int i = 0; i < p.K; ++i)
    if (p.q[i] >= kF64ModLimit) return false;
  return tab < p.K || p.t < kF64ModLimit;
#endif
}

inline DevConsts make_devconsts(const Params &p) {
  DevConsts c{};
  c.N = p.N;
  c.logn = p.logn;
  c.K = p.K;
  c.L = p.L;
  c.t = p.t;
  c.half_t = p.half_t;
  c.q_mod_t = p.q_mod_t;
  for (size_t i = 0; i < p.tab.size(); ++i) {
    if (!p.tab[i].q) continue;
    c.mod[i] = make_devmod(p.tab[i].q);
    c.n_inv[i] = w2(p.tab[i].n_inv);
    c.f64[i] = table_is_f64(p, static_cast<int>(i)) ? 1 : 0;
    c.qf[i] = static_cast<double>(p.tab[i].q);
    c.qinvf[i] = 1.0 / static_cast<double>(p.tab[i].q);
    c.n_inv_f[i] = D2{static_cast<double>(p.tab[i].n_inv.w), static_cast<double>(p.tab[i].n_inv.w) / static_cast<double>(p.tab[i].q)};
  }
  c.half_sp = p.half_sp;
  for (int i = 0; i < kMaxLimbs; ++i) {
    c.q_div_t_mod_q[i] = p.q_div_t_mod_q[i];
    c.half_sp_mod_q[i] = p.half_sp_mod_q[i];
    c.inv_sp_mod_q[i] = w2(p.inv_sp_mod_q[i]);
    if (i < p.L && p.q[i])
      c.inv_sp_f[i] = D2{static_cast<double>(p.inv_sp_mod_q[i].w), static_cast<double>(p.inv_sp_mod_q[i].w) / static_cast<double>(p.q[i])};
    c.mtilde_ipq[i] = w2(p.mtilde_mod_q[i]);
    c.t_ipq[i] = w2(p.t_inv_punct_q[i]);
    c.q2mt[i] = p.q2mt[i];
    c.q_mod_bsk[i] = p.q_mod_bsk[i];
    c.inv_mt_bsk[i] = w2(p.inv_mt_bsk[i]);
    c.t_mod_bsk[i] = w2(p.t_mod_bsk[i]);
    c.inv_q_bsk[i] = w2(p.inv_q_bsk[i]);
    c.inv_punct_b[i] = w2(p.inv_punct_b[i]);
    c.b2msk[i] = p.b2msk[i];
    c.pb_mod_q[i] = p.pb_mod_q[i];
    for (int j = 0; j < kMaxLimbs; ++j) {
      c.q2bsk[i][j] = p.q2bsk[i][j];
      c.b2q[i][j] = p.b2q[i][j];
    }
  }
  c.neg_inv_q_mt = p.neg_inv_q_mt;
  c.inv_pb_msk = w2(p.inv_pb_msk);
  return c;
}

}  // namespace hhe
