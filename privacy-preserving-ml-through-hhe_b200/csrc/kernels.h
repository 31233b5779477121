// Kernel bodies (functors, host+device). Each is launched through launch.h as one CUDA kernel on sm_100a.
// Layouts: ciphertext batch u64[count][size][L][N]; key-switch accumulators u64[count][2][K][N];
// key-switching keys W2[L][2][K][N] (value + Shoup quotient, NTT form); twiddles W2[table][fwd|inv][N].
#pragma once
#include <type_traits>
#include "devconsts.h"
#include "ntt_core.h"
#include "tmem.h"

namespace hhe {

// global-memory IO functors for the fused first / last register pass of the FP64 transforms (ntt_core.h: SmemIO)
struct LoadU64 {  // canonical residues -> doubles
  static constexpr bool kLoad = true, kStore = false, kGroupOut = false;
  const u64 *src;
  HD double load(int i) const { return u_to_f(src[i]); }
  HD void store(int, double) const {}
  HD void group_out(int, const double *) const {}
};
struct LoadLift {  // centred lift of a plaintext coefficient into the limb (Evaluator::multiply_plain)
  static constexpr bool kLoad = true, kStore = false, kGroupOut = false;
  const u64 *src;
  u64 thr, inc;
  HD double load(int i) const {
    const u64 m = src[i];
    return u_to_f(m >= thr ? m + inc : m);
  }
  HD void store(int, double) const {}
  HD void group_out(int, const double *) const {}
};
struct StoreScaled {  // inverse transform output: multiply by N^-1, canonicalise, store
  static constexpr bool kLoad = false, kStore = true, kGroupOut = false;
  u64 *dst;
  D2 ninv;
  double q, qinv;
  HD double load(int) const { return 0.0; }
  HD void store(int i, double v) const { dst[i] = f_canonical(f_mulmod_const(v, ninv, q), q, qinv); }
  HD void group_out(int, const double *) const {}
  // two-step form used by the cluster kernels: aux(j) issues the global loads store(j, v, aux) needs, so a caller can have
  // several of them in flight before the first dependent instruction
  struct Aux {};
  HD Aux aux(int) const { return Aux{}; }
  HD void store(int i, double v, const Aux &) const { store(i, v); }
};

struct LoadCorr {  // corr[j] = (r0[j] mod q_i) - (half mod q_i), r0 = acc0[special] + half mod q_sp  (Corr0MacBody)
  static constexpr bool kLoad = true, kStore = false, kGroupOut = false;
  const u64 *sp;
  u64 half_sp, half_i, q;
  DevMod mi, msp;
  HD double load(int j) const {
    const u64 r = csub(sp[j] + half_sp, msp.q);
    const u64 ri = msp.q > q ? (msp.q < 2 * q ? csub(r, q) : barrett64(r, mi)) : r;
    return u_to_f(sub_mod(ri, half_i, q));
  }
  HD void store(int, double) const {}
  HD void group_out(int, const double *) const {}
};

struct TwRef {
  const W2 *base;  // all tables
  u64 N;
  HD const W2 *fwd(int tab) const { return base + (static_cast<size_t>(tab) * 2) * N; }
  HD const W2 *inv(int tab) const { return base + (static_cast<size_t>(tab) * 2 + 1) * N; }
  // FP64 tables reuse the same 16 bytes per entry as two planes of N doubles: index-major, then component-major
  int gmin;
  HD F64Tw fwd_f(int tab) const {
    const double *p = reinterpret_cast<const double *>(fwd(tab));
    return F64Tw{p, p + N, gmin};
  }
  HD F64Tw inv_f(int tab) const {
    const double *p = reinterpret_cast<const double *>(inv(tab));
    return F64Tw{p, p + N, gmin};
  }
};

// Key-switch digit transforms (FP64 path): digits are residues mod q_J <= 2 q_k. Bounds in units of q_k / 2 after one /
// two folded Cooley-Tukey stages (2 + 0.94 and 2.94 + 1.13), and the bound (units of q_k / 16) the inner product accepts
// from the transform: 8 products of magnitude <= 1.25 q_k keep the accumulators below 10 q_k.
constexpr int kKsFold1 = 6, kKsFold2 = 9, kKsOut16 = 64;

// ------------------------------------------------------------------------------------------------------------
// Half-limb forward transforms (FP64 path). The first Cooley-Tukey stage of an N-point transform pairs residue i with
// i + N/2; after it the two halves are independent N/2-point transforms. A CTA that owns half h applies that stage while
// loading (it reads the whole limb, L2 serves the second reader) and then needs only N/2 doubles of shared memory
// (68 KiB at N = 16384): two CTAs per SM, whose load, barrier and store phases overlap. When the sub-transform's
// odd-sized first register pass is a single stage it is folded into the load too.
// LD: raw(i) fetches input residue i of the limb (global memory), cvt() turns it into a double of magnitude <= 2q.
struct RawU64 {
  const u64 *src;
  HD u64 raw(int i) const { return src[i]; }
  HD const u64 *ptr(int i) const { return src + i; }
  HD double cvt(u64 v) const { return u_to_f(v); }
};
struct RawLift {  // centred lift of a plaintext coefficient into the limb (Evaluator::multiply_plain)
  const u64 *src;
  u64 thr, inc;
  HD u64 raw(int i) const { return src[i]; }
  HD const u64 *ptr(int i) const { return src + i; }
  HD double cvt(u64 m) const { return u_to_f(m >= thr ? m + inc : m); }
};
struct RawCorr {  // corr[j] = (r0[j] mod q_i) - (half mod q_i), r0 = acc0[special] + half mod q_sp  (Corr0Mac)
  const u64 *sp;
  u64 half_sp, half_i, q;
  DevMod mi, msp;
  HD u64 raw(int j) const { return sp[j]; }
  HD const u64 *ptr(int j) const { return sp + j; }
  HD double cvt(u64 v) const {
    const u64 r = csub(v + half_sp, msp.q);
    // r < q_sp: no reduction when q_sp <= q_i, one conditional subtraction when q_sp < 2 q_i (every BFVDefault set: the primes of a
    // set differ by at most one bit), the 128-bit Barrett reduction (integer multiplies, which stall the FP64 pipe) only otherwise
    const u64 ri = msp.q > q ? (msp.q < 2 * q ? csub(r, q) : barrett64(r, mi)) : r;
    return u_to_f(sub_mod(ri, half_i, q));
  }
};

// With WARP_LOCAL (CTA of NT = S/16 threads, two folded stages) warp w folds the residues i, i + S/2 with
// i in [256 w, 256 w + 256): exactly the residues its own threads read in the LAST register pass of the previous
// transform that used the same buffer, so back-to-back transforms need only a warp-level sync in between.
HD constexpr bool half_warp_local(int logh, int nt) { return NttSchedule_first(logh) == 1 && nt * 16 == (1 << logh); }

template <int LOGH, class LD, bool WARP_LOCAL = false>
HD void fwd_half_load_f64(double *fm, F64Tw twk, double q, double qi, int h, int nt, const LD &ld) {
  constexpr int S = 1 << LOGH;
  constexpr bool kFold = NttSchedule<LOGH>::kFirst == 1;
  const D2 w1{twk.idx[1], f_mul(twk.idx[1], qi)};
  if (kFold) {
    const D2 w2{twk.idx[2 + h], f_mul(twk.idx[2 + h], qi)};
    FOR_THREADS(tid, nt) {
      // thread's residues: first + k * step; software pipeline: the four loads of iteration n+1 are in flight while
      // iteration n is computed
      const int first = WARP_LOCAL ? ((tid >> 5) << 8) + (tid & 31) : tid;
      const int step = WARP_LOCAL ? 32 : nt;
      const int count = WARP_LOCAL ? 8 : (S / 2 - tid + nt - 1) / nt;
      if constexpr (WARP_LOCAL) {
      // ks_digits: loads two iterations ahead (8 more registers, still 64 without spills: ks_digits -1.4 %; the other users of this
      // loader spill with it and lose 2 %, so they keep the one-iteration pipeline below; profiles/r2_ab_twiddle_pipeline.txt)
      u64 v[4] = {0, 0, 0, 0}, nv[4] = {0, 0, 0, 0}, nn[4] = {0, 0, 0, 0};
      if (count > 0) {
        v[0] = ld.raw(first);
        v[1] = ld.raw(first + S);
        v[2] = ld.raw(first + S / 2);
        v[3] = ld.raw(first + S / 2 + S);
      }
      if (count > 1) {
        nv[0] = ld.raw(first + step);
        nv[1] = ld.raw(first + step + S);
        nv[2] = ld.raw(first + step + S / 2);
        nv[3] = ld.raw(first + step + S / 2 + S);
      }
      for (int k = 0; k < count; ++k) {
        const int i = first + k * step, in = i + 2 * step;
        if (k + 2 < count) {
          nn[0] = ld.raw(in);
          nn[1] = ld.raw(in + S);
          nn[2] = ld.raw(in + S / 2);
          nn[3] = ld.raw(in + S / 2 + S);
        }
        const double t0 = f_mulmod_const(ld.cvt(v[1]), w1, q), t1 = f_mulmod_const(ld.cvt(v[3]), w1, q);
        const double a0 = h ? f_add(ld.cvt(v[0]), -t0) : f_add(ld.cvt(v[0]), t0);  // |.| <= 2.94q
        const double a1 = h ? f_add(ld.cvt(v[2]), -t1) : f_add(ld.cvt(v[2]), t1);
        const double tt = f_mulmod_const(a1, w2, q);
        fm[pidx(i)] = f_add(a0, tt);  // |.| <= 4.1q
        fm[pidx(i + S / 2)] = f_add(a0, -tt);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = nv[e], nv[e] = nn[e];
      }
      } else {
      u64 v[4] = {0, 0, 0, 0}, nv[4] = {0, 0, 0, 0};
      if (count > 0) {
        v[0] = ld.raw(first);
        v[1] = ld.raw(first + S);
        v[2] = ld.raw(first + S / 2);
        v[3] = ld.raw(first + S / 2 + S);
      }
      for (int k = 0; k < count; ++k) {
        const int i = first + k * step, in = i + step;
        if (k + 1 < count) {
          nv[0] = ld.raw(in);
          nv[1] = ld.raw(in + S);
          nv[2] = ld.raw(in + S / 2);
          nv[3] = ld.raw(in + S / 2 + S);
        }
        const double t0 = f_mulmod_const(ld.cvt(v[1]), w1, q), t1 = f_mulmod_const(ld.cvt(v[3]), w1, q);
        const double a0 = h ? f_add(ld.cvt(v[0]), -t0) : f_add(ld.cvt(v[0]), t0);  // |.| <= 2.94q
        const double a1 = h ? f_add(ld.cvt(v[2]), -t1) : f_add(ld.cvt(v[2]), t1);
        const double tt = f_mulmod_const(a1, w2, q);
        fm[pidx(i)] = f_add(a0, tt);  // |.| <= 4.1q
        fm[pidx(i + S / 2)] = f_add(a0, -tt);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = nv[e];
      }
      }
    }
  } else {
    FOR_THREADS(tid, nt) {
      constexpr int U = 4;
      for (int i0 = tid; i0 < S; i0 += nt * U) {
        u64 xs[U], ys[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * nt;
          if (i < S) {
            xs[u] = ld.raw(i);
            ys[u] = ld.raw(i + S);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * nt;
          if (i < S) {
            const double t = f_mulmod_const(ld.cvt(ys[u]), w1, q);
            fm[pidx(i)] = h ? f_add(ld.cvt(xs[u]), -t) : f_add(ld.cvt(xs[u]), t);  // |.| <= 2.94q
          }
        }
      }
    }
  }
  SYNC();
}

// CTA size of the whole-limb kernels (Engine: ntt_threads)
HD constexpr int whole_threads(int logs) { return (1 << logs) / 8 < 32 ? 32 : ((1 << logs) / 8 > HHE_MAX_THREADS ? HHE_MAX_THREADS : (1 << logs) / 8); }

// CTA size of the half-limb kernels (one radix-8 group per thread and pass at least, 512 threads at most)
HD constexpr int half_threads(int logh, int maxt = 512) { return (1 << logh) / 8 < 32 ? 32 : ((1 << logh) / 8 > maxt ? maxt : (1 << logh) / 8); }

// the register passes that follow fwd_half_load_f64; results in shared memory (or handed to IO::group_out), |.| <= MAXOUT16/16 q
template <int LOGH, class IO = SmemIO, int MAXOUT16 = kF64AnyOut16, int MAXT = 512>
HD void fwd_half_passes_f64(double *fm, F64Tw twk, double q, double qi, int h, int nt, const IO &io = IO()) {
  constexpr bool kFold = NttSchedule<LOGH>::kFirst == 1;
  constexpr int NT = half_threads(LOGH, MAXT);  // every launch of a half-limb kernel uses exactly this CTA size
  if (kFold)
    ntt_fwd_core_f64_from<LOGH, 1, (kFold ? 1 : 0), kKsFold2, IO, MAXOUT16, NT>(fm, twk, q, qi, h, nt, io);
  else
    ntt_fwd_core_f64<LOGH, 1, kKsFold1, IO, MAXOUT16, NT>(fm, twk, q, qi, h, nt, io);
}

// Folded load + FIRST REGISTER PASS in one step (S = 16 nt, i.e. N = 16384 with 512 threads). Thread t's eight fold iterations
// i = t + e S/16 (e = 0..7) produce the residues t + e S/16 and S/2 + t + e S/16: exactly the two radix-8 groups (element stride
// S/16) that the same thread owns in the first register pass. So the fold's outputs never visit shared memory: the thread runs
// the pass's three stages on them in registers and stores the pass's outputs. One shared-memory round trip (S doubles written and
// read back, 14 % of the LSU wavefronts of a key-switch digit) and one barrier less per transform; the FP64 work is unchanged.
// `before_store`: executed by every thread between its last global load and its first shared-memory store (the key-switch kernel
// places the barrier there that protects the previous digit's last pass, which is still reading the buffer).
HD constexpr bool half_fused_first_pass(int logh, int nt) {
#if defined(HHE_NO_FUSED_FIRST_PASS)
  return false && logh && nt;
#else
  return NttSchedule_first(logh) == 1 && nt * 16 == (1 << logh) && logh >= 7;
#endif
}

template <int LOGH, class LD, class IO, int MAXOUT16, int MAXT, class PRE>
HD void fwd_half_fused_f64(double *fm, F64Tw twk, double q, double qi, int h, int nt, const LD &ld, const IO &io, const PRE &before_store) {
  constexpr int S = 1 << LOGH;
  constexpr int NT = half_threads(LOGH, MAXT);
  static_assert(NttSchedule<LOGH>::kFirst == 1 && NT * 16 == S, "fused first pass: S = 16 nt and a single-stage odd pass");
  using Pass1 = FwdChainF64<LOGH, 1, 1, kKsFold2 * 8, MAXOUT16, NT>;  // the pass being fused (local stages 1..3)
  static_assert(Pass1::R == 3 && !Pass1::kLast, "fused first pass is a full radix-8 pass followed by others");
  using Rest = FwdChainF64<LOGH, 1, 4, Pass1::kOut, MAXOUT16, NT>;
  constexpr int LG = LOGH - 4;  // element stride of the pass: S/16 = NT
  const D2 w1{twk.idx[1], f_mul(twk.idx[1], qi)};
  const D2 w2{twk.idx[2 + h], f_mul(twk.idx[2 + h], qi)};
  FOR_THREADS(tid, nt) {
    double xa[8], xb[8];
    u64 v[4], nv[4] = {0, 0, 0, 0};
    v[0] = ld.raw(tid);
    v[1] = ld.raw(tid + S);
    v[2] = ld.raw(tid + S / 2);
    v[3] = ld.raw(tid + S / 2 + S);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e + 1 < 8) {
        const int in = tid + (e + 1) * NT;
        nv[0] = ld.raw(in);
        nv[1] = ld.raw(in + S);
        nv[2] = ld.raw(in + S / 2);
        nv[3] = ld.raw(in + S / 2 + S);
      }
      const double t0 = f_mulmod_const(ld.cvt(v[1]), w1, q), t1 = f_mulmod_const(ld.cvt(v[3]), w1, q);
      const double a0 = h ? f_add(ld.cvt(v[0]), -t0) : f_add(ld.cvt(v[0]), t0);  // |.| <= 2.94q
      const double a1 = h ? f_add(ld.cvt(v[2]), -t1) : f_add(ld.cvt(v[2]), t1);
      const double tt = f_mulmod_const(a1, w2, q);
      xa[e] = f_add(a0, tt);  // residue tid + e NT          |.| <= 4.1q
      xb[e] = f_add(a0, -tt);  // residue S/2 + tid + e NT
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = nv[c];
    }
    // the two groups of the first register pass: g = tid (block 0) and g = NT + tid (block 1), element stride NT
    double wv[8];
    group_tw_f64<3, LOGH, 1, 1>(twk, h, tid, wv);
    group_math_f64<3, false, Pass1::kHalfMode ? kHalf : kNone>(xa, wv, q, qi);
    group_tw_f64<3, LOGH, 1, 1>(twk, h, NT + tid, wv);
    group_math_f64<3, false, Pass1::kHalfMode ? kHalf : kNone>(xb, wv, q, qi);
    before_store();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      fm[pidx(tid + (e << LG))] = xa[e];
      fm[pidx(S / 2 + tid + (e << LG))] = xb[e];
    }
  }
  sync_domain<LG, NT>();  // as after the unfused pass: its element stride is the larger one
  Rest::run(fm, twk, q, qi, h, nt, io);
}

struct NoPre {
  HD void operator()() const {}
};
struct CtaBarrier {  // every thread of the CTA calls it (uniform control flow)
  HD void operator()() const { SYNC(); }
};

// folded load + all register passes of a half-limb forward transform (fused first pass where the CTA shape allows it; FUSE = false
// for loaders whose conversion needs too many registers next to the 16 residues: corr_mac's RawCorr spills and loses 2 %)
template <int LOGH, class LD, class IO = SmemIO, int MAXOUT16 = kF64AnyOut16, int MAXT = 512, bool FUSE = true>
HD void fwd_half_transform_f64(double *fm, F64Tw twk, double q, double qi, int h, int nt, const LD &ld, const IO &io = IO()) {
  if constexpr (FUSE && half_fused_first_pass(LOGH, half_threads(LOGH, MAXT))) {
    fwd_half_fused_f64<LOGH, LD, IO, MAXOUT16, MAXT>(fm, twk, q, qi, h, nt, ld, io, NoPre{});
  } else {
    fwd_half_load_f64<LOGH>(fm, twk, q, qi, h, nt, ld);
    fwd_half_passes_f64<LOGH, IO, MAXOUT16, MAXT>(fm, twk, q, qi, h, nt, io);
  }
}

constexpr int kMaxMapLimbs = 3 * kMaxLimbs;  // size-3 ciphertext in the Bsk base
struct TabMap {  // limb index inside an item -> NTT table id
  unsigned char id[kMaxMapLimbs];
};

// ------------------------------------------------------------------------------------------------------------
// Plain batched NTT / inverse NTT of whole limbs: grid = count * limbs, CTA = one limb in shared memory.
// Replaces seal::util::ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey (seal/util/ntt.h:195-340).
template <int LOGS>
struct NttBody {
  static constexpr const char *kName = "ntt";
  const u64 *in;
  u64 *out;
  const DevConsts *C;
  TwRef tw;
  TabMap map;
  int limbs;
  int inverse;
  size_t istride;  // words between consecutive items (limbs * N when dense)
  size_t lstride;  // words between consecutive limbs of an item (N when dense)
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    u64 *sm = reinterpret_cast<u64 *>(smem);
    const int tab = map.id[bid % limbs];
    const u64 q = C->mod[tab].q;
    const size_t at = static_cast<size_t>(bid / limbs) * istride + static_cast<size_t>(bid % limbs) * lstride;
    const u64 *src = in + at;
    u64 *dst = out + at;
    if (C->f64[tab]) {  // FP64-pipe transform (modarith_f64.h)
      double *fm = reinterpret_cast<double *>(smem);
      const double qd = C->qf[tab], qi = C->qinvf[tab];
      if (!inverse) {
        ntt_fwd_core_f64<LOGS, 0, 2, LoadU64, kF64AnyOut16, whole_threads(LOGS)>(fm, tw.fwd_f(tab), qd, qi, 0, nt, LoadU64{src});
        FOR_THREADS(tid, nt) {
          for (int i = tid; i < S; i += nt) dst[i] = f_canonical(fm[pidx(i)], qd, qi);
        }
      } else {
        FOR_THREADS(tid, nt) {
          for (int i = tid; i < S; i += nt) fm[pidx(i)] = u_to_f(src[i]);
        }
        SYNC();
        ntt_inv_core_f64<LOGS, 0, StoreScaled, whole_threads(LOGS)>(fm, tw.inv_f(tab), qd, qi, 0, nt, StoreScaled{dst, C->n_inv_f[tab], qd, qi});
      }
      return;
    }
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) sm[pidx(i)] = src[i];
    }
    SYNC();
    if (!inverse) {
      ntt_fwd_core<LOGS>(sm, tw.fwd(tab), q, 1, nt);
      const DevMod mm = C->mod[tab];
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) dst[i] = barrett64(sm[pidx(i)], mm);
      }
    } else {
      ntt_inv_core<LOGS>(sm, tw.inv(tab), q, 1, nt);
      const W2 ninv = C->n_inv[tab];
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) dst[i] = mul_shoup(sm[pidx(i)], ninv, q);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Key-switch digit kernel (Evaluator::switch_key_inplace inner loops, seal/evaluator.h:1260):
//   acc[b][c][k] = sum_J NTT_k( target[b][J] mod q_k ) (.) key[J][c][k]
// One CTA per (b, key limb k, half h): the first Cooley-Tukey stage is folded into the load so each CTA owns an
// independent N/2-point sub-transform (64 KiB at N=16384) and keeps both components' accumulators for its half in
// shared memory (128 KiB) across the digit loop -- the accumulators never touch HBM until the inner product is done.
// Grid order puts b fastest so co-resident CTAs read the same key tile from L2.
template <int LOGH>
struct KsDigitsBody {
  static constexpr const char *kName = "ks_digits";
  const u64 *target;  // [count][..] : item b's target polynomial (L limbs) at target + b*stride
  size_t stride;
  const W2 *key;  // [L][2][K][N]
  u64 *acc;       // [count][2][K][N], NTT form, canonical
  const DevConsts *C;
  TwRef tw;
  int count;
  // Optional (FP64 path): reuse[b][J] = NTT_J(c_J) of the polynomial whose Galois image is the target, and the NTT-domain
  // permutation of that Galois element. Then NTT_J(target_J) = reuse[b][J][perm[.]] and digit J needs no transform on
  // key limb k = J (8 of the 72 digit transforms of a key switch).
  const u64 *reuse;
  size_t reuse_stride;
  const u32 *perm;
  // FP64-pipe version of the same computation (q_k <= 2^49): digits, twiddles, key and accumulators are doubles.
  HD void run_f64(int b, int k, int h, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGH;
    const int N = 2 * S;
    const int K = C->K, L = C->L;
    double *fm = reinterpret_cast<double *>(smem);
    double *acc0 = fm + ntt_smem_words(S);
    double *acc1 = acc0 + S;
    const double q = C->qf[k], qi = C->qinvf[k];
    const F64Tw twk = tw.fwd_f(k);
    const D2 w1{twk.idx[1], f_mul(twk.idx[1], qi)};
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) acc0[i] = acc1[i] = 0.0;
    }
    // When the sub-transform's odd-sized first pass is a single stage it is folded into the load as well: the thread
    // that owns residues i and i + S/2 applies global stages 0 and 1 before anything is written to shared memory.
    constexpr bool kFold = NttSchedule<LOGH>::kFirst == 1;
    const D2 w2{twk.idx[2 + h], f_mul(twk.idx[2 + h], qi)};
    for (int J = 0; J < L; ++J) {
      const u64 *dig = target + static_cast<size_t>(b) * stride + static_cast<size_t>(J) * N;
      // digits are residues mod q_J < 2 q_k (all primes of a parameter set have the same size): as doubles they are
      // valid inputs (< 8 q_k) without re-reduction
      if (reuse && J == k) {
        const u64 *rn = reuse + static_cast<size_t>(b) * reuse_stride + static_cast<size_t>(J) * N;
        const u32 *pm = perm + static_cast<size_t>(h) * S;
        FOR_THREADS(tid, nt) {
          constexpr int U = 4;
          for (int i0 = tid; i0 < S; i0 += nt * U) {
            u32 pj[U];
            u64 rv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) pj[u] = pm[i0 + u * nt < S ? i0 + u * nt : i0];
#pragma unroll
            for (int u = 0; u < U; ++u) rv[u] = rn[pj[u]];
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (i0 + u * nt < S) fm[pidx(i0 + u * nt)] = u_to_f(rv[u]);
          }
        }
        SYNC();
      } else if (kFold) {
        FOR_THREADS(tid, nt) {
          constexpr int U = 2;
          for (int i0 = tid; i0 < S / 2; i0 += nt * U) {
            u64 v[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int i = i0 + u * nt;
              if (i < S / 2) {
                v[u][0] = dig[i];
                v[u][1] = dig[i + S];
                v[u][2] = dig[i + S / 2];
                v[u][3] = dig[i + S / 2 + S];
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int i = i0 + u * nt;
              if (i < S / 2) {
                const double t0 = f_mulmod_const(u_to_f(v[u][1]), w1, q), t1 = f_mulmod_const(u_to_f(v[u][3]), w1, q);
                const double a0 = h ? f_add(u_to_f(v[u][0]), -t0) : f_add(u_to_f(v[u][0]), t0);  // |.| <= 2.94q
                const double a1 = h ? f_add(u_to_f(v[u][2]), -t1) : f_add(u_to_f(v[u][2]), t1);
                const double tt = f_mulmod_const(a1, w2, q);
                fm[pidx(i)] = f_add(a0, tt);  // |.| <= 4.1q
                fm[pidx(i + S / 2)] = f_add(a0, -tt);
              }
            }
          }
        }
        SYNC();
        ntt_fwd_core_f64_from<LOGH, 1, (kFold ? 1 : 0), kKsFold2, SmemIO, kKsOut16>(fm, twk, q, qi, h, nt);
      } else {
        FOR_THREADS(tid, nt) {
          constexpr int U = 4;
          for (int i0 = tid; i0 < S; i0 += nt * U) {
            u64 xs[U], ys[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int i = i0 + u * nt;
              xs[u] = i < S ? dig[i] : 0;
              ys[u] = i < S ? dig[i + S] : 0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int i = i0 + u * nt;
              if (i < S) {
                const double t = f_mulmod_const(u_to_f(ys[u]), w1, q);
                fm[pidx(i)] = h ? f_add(u_to_f(xs[u]), -t) : f_add(u_to_f(xs[u]), t);  // |.| <= 2.94q
              }
            }
          }
        }
        SYNC();
        ntt_fwd_core_f64<LOGH, 1, kKsFold1, SmemIO, kKsOut16>(fm, twk, q, qi, h, nt);
      }
      // compact FP64 key: double[L][2][K][N] (8 bytes per residue; k/q is formed as k * (1/q))
      const double *k0 = reinterpret_cast<const double *>(key) + ((static_cast<size_t>(J) * 2 + 0) * K + k) * N + static_cast<size_t>(h) * S;
      const double *k1 = k0 + static_cast<size_t>(K) * N;
      FOR_THREADS(tid, nt) {
#pragma unroll 4
        for (int i = tid; i < S; i += nt) {
          const double v = fm[pidx(i)];
          const double a = k0[i], c = k1[i];
          // L <= 8 terms of magnitude <= 1.25q (|v| <= 4q, kKsOut16): |acc| <= 10q < 2^53, every partial sum is an exact integer
          acc0[i] = f_add(acc0[i], f_mulmod_var(v, a, q, qi));
          acc1[i] = f_add(acc1[i], f_mulmod_var(v, c, q, qi));
        }
      }
      SYNC();
    }
    u64 *o0 = acc + ((static_cast<size_t>(b) * 2 + 0) * K + k) * N + static_cast<size_t>(h) * S;
    u64 *o1 = o0 + static_cast<size_t>(K) * N;
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) {
        o0[i] = f_canonical(acc0[i], q, qi);
        o1[i] = f_canonical(acc1[i], q, qi);
      }
    }
  }
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGH;
    const int N = 2 * S;
    const int K = C->K, L = C->L;
    // the 2K CTAs of one item are adjacent in the grid: they share the item's digits through L2, and the ~8 items
    // in flight across the chip sweep the whole (L2-resident) key
    const int b = bid / (2 * K);
    const int kh = bid % (2 * K);
    const int k = kh >> 1, h = kh & 1;
    u64 *sm = reinterpret_cast<u64 *>(smem);
    u64 *acc0 = sm + ntt_smem_words(S);
    u64 *acc1 = acc0 + S;
    if (C->f64[k]) {
      run_f64(b, k, h, nt, smem);
      return;
    }
    const DevMod mk = C->mod[k];
    const u64 q = mk.q, two_q = q << 1, nq = 0 - q, four_q = q << 2;
    const bool wide = q < kWideSlackLimit;
    const W2 *twk = tw.fwd(k);
    const W2 w1 = twk[1];
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) acc0[i] = acc1[i] = 0;
    }
    for (int J = 0; J < L; ++J) {
      const u64 *dig = target + static_cast<size_t>(b) * stride + static_cast<size_t>(J) * N;
      const bool reduce = C->mod[J].q > q;
      FOR_THREADS(tid, nt) {
        constexpr int U = 8;  // loads in flight per thread (2*U 64-bit requests)
        for (int i0 = tid; i0 < S; i0 += nt * U) {
          u64 xs[U], ys[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * nt;
            xs[u] = i < S ? dig[i] : 0;
            ys[u] = i < S ? dig[i + S] : 0;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * nt;
            if (i < S) {
              u64 x = xs[u], y = ys[u];
              if (reduce) {
                x = barrett64(x, mk);
                y = barrett64(y, mk);
              }
              if (wide) {
                const u64 t = mul_shoup_wide(y, w1.w, w1.ws, nq);
                sm[pidx(i)] = h ? x + four_q - t : x + t;
              } else {
                const u64 t = mul_shoup_lazy(y, w1.w, w1.ws, q);
                sm[pidx(i)] = h ? x + two_q - t : x + t;
              }
            }
          }
        }
      }
      SYNC();
      ntt_fwd_core<LOGH>(sm, twk, q, 2 + h, nt);
      const W2 *k0 = key + ((static_cast<size_t>(J) * 2 + 0) * K + k) * N + static_cast<size_t>(h) * S;
      const W2 *k1 = k0 + static_cast<size_t>(K) * N;
      FOR_THREADS(tid, nt) {
#pragma unroll 4
        for (int i = tid; i < S; i += nt) {
          const u64 v = sm[pidx(i)];
          const W2 a = k0[i], c = k1[i];
          if (wide) {  // each term < 4q: L <= 16 digits stay below 64q < 2^63 without reduction
            acc0[i] += mul_shoup_wide(v, a.w, a.ws, nq);
            acc1[i] += mul_shoup_wide(v, c.w, c.ws, nq);
          } else {
            u64 s0 = acc0[i] + mul_shoup_lazy(v, a.w, a.ws, q);
            u64 s1 = acc1[i] + mul_shoup_lazy(v, c.w, c.ws, q);
            acc0[i] = s0 >= two_q ? s0 - two_q : s0;
            acc1[i] = s1 >= two_q ? s1 - two_q : s1;
          }
        }
      }
      SYNC();
    }
    u64 *o0 = acc + ((static_cast<size_t>(b) * 2 + 0) * K + k) * N + static_cast<size_t>(h) * S;
    u64 *o1 = o0 + static_cast<size_t>(K) * N;
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) {
        o0[i] = barrett64(acc0[i], mk);
        o1[i] = barrett64(acc1[i], mk);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Key-switch digit kernel, tensor-memory version (FP64 path). Same computation as KsDigitsBody::run_f64, but
//   * the 2 x N/2 inner-product accumulators of the CTA live in tensor memory (tmem.h) as a per-thread register-file
//     extension instead of 128 KiB of shared memory  ->  68 KiB of shared memory per CTA, two CTAs per SM (512 threads each),
//     whose phases overlap;
//   * the multiply-accumulate with the key is fused into the last register pass of the digit transform: each thread keeps
//     the 8 outputs of a group, multiplies them by the key and adds them to its own TMEM slots (no shared-memory round trip,
//     one barrier less per digit). For that the key is stored "group-major": residue 8g+e of a half at e*(S/8)+g, so the
//     key loads of a warp are contiguous (Engine::load_ksk permutes once).
struct KsMacOut {
  static constexpr bool kLoad = false, kStore = false, kGroupOut = true;
  const double *k0, *k1;  // group-major key planes of (digit, component 0/1, key limb, half)
  int G, nt;
  double q, qi;
  TmemAcc tm;
  HD double load(int) const { return 0.0; }
  HD void store(int, double) const {}
  HD void group_out(int g, const double *x) const {
    const int gi = g / nt;  // which of this thread's groups
    // The tensor-memory accesses are ordered asm statements the compiler will not move loads across: component 1's key loads are
    // therefore issued by hand before component 0's store, so their latency overlaps the TMEM round trip.
    double kv[8], kn[8], a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) kv[e] = k0[g + static_cast<size_t>(e) * G];
    tm.ld8(gi * 2, a);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = f_add(a[e], f_mulmod_var(x[e], kv[e], q, qi));
#pragma unroll
    for (int e = 0; e < 8; ++e) kn[e] = k1[g + static_cast<size_t>(e) * G];
    tm.st8(gi * 2, a);
    tm.ld8(gi * 2 + 1, a);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = f_add(a[e], f_mulmod_var(x[e], kn[e], q, qi));
    tm.st8(gi * 2 + 1, a);
  }
};

template <int LOGH, int MAXT = 512>
struct KsDigitsTmemBody {
  static constexpr const char *kName = "ks_digits";
  static constexpr int kMaxThreads = MAXT, kMinBlocks = 2;
  const u64 *target;
  size_t stride;
  const double *key;  // group-major FP64 key [L][2][K][N]
  u64 *acc;           // [count][2][K][N], NTT form, canonical
  const DevConsts *C;
  TwRef tw;
  int count;
  const u64 *reuse;  // see KsDigitsBody
  size_t reuse_stride;
  const u32 *perm;
  int pf;  // L2 prefetch distance in items (0: off)
  static constexpr size_t smem_bytes(int nt, bool emulate) {
    return (ntt_smem_words(1 << LOGH) + 2) * 8 + (emulate ? static_cast<size_t>(2) * (1 << LOGH) * 8 : 0) + 0 * nt;
  }
  HD void operator()(int bid, int, unsigned char *smem) const {
    constexpr int nt = half_threads(LOGH, MAXT);  // the launch uses exactly this CTA size (Engine::launch_ks_digits)
    constexpr int S = 1 << LOGH, G = S / 8;
    const int N = 2 * S;
    const int K = C->K, L = C->L;
    const int b = bid / (2 * K), kh = bid % (2 * K), k = kh >> 1, h = kh & 1;
    if (pf && kh == 0 && b + pf < count) {  // digits (and the reused transforms) of the item that starts two waves later
      cta_prefetch_l2(target + static_cast<size_t>(b + pf) * stride, sizeof(u64) * L * N);
      if (reuse) cta_prefetch_l2(reuse + static_cast<size_t>(b + pf) * reuse_stride, sizeof(u64) * L * N);
    }
    const int gpt = G / nt;  // groups (of 8 residues) per thread; 2 slots (components) each
    double *fm = reinterpret_cast<double *>(smem);
    u32 *tslot = reinterpret_cast<u32 *>(fm + ntt_smem_words(S));
    double *emu = fm + ntt_smem_words(S) + 2;
    (void)emu;

    const double q = C->qf[k], qi = C->qinvf[k];
    const F64Tw twk = tw.fwd_f(k);
    constexpr bool kFold = NttSchedule<LOGH>::kFirst == 1;
    constexpr bool kWarpLocal = half_warp_local(LOGH, nt);
    u32 tbase = 0;
#if defined(__CUDA_ARCH__)
    const int ncols = tmem_columns(nt, gpt * 2);
    tbase = tmem_alloc_cta(tslot, ncols);
#else
    (void)tslot;
#endif
    FOR_THREADS(tid, nt) {
      const TmemAcc tm = TmemAcc::make(tbase, tid, nt, gpt * 2, emu);
      const double z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int s = 0; s < gpt * 2; ++s) tm.st8(s, z);
    }
    for (int J = 0; J < L; ++J) {
      const double *k0 = key + ((static_cast<size_t>(J) * 2 + 0) * K + k) * N + static_cast<size_t>(h) * S;
      const double *k1 = k0 + static_cast<size_t>(K) * N;
      if (reuse && J == k) {
        // NTT_J(target_J) is the permuted NTT of the source polynomial: no transform, straight to the accumulation
        const u64 *rn = reuse + static_cast<size_t>(b) * reuse_stride + static_cast<size_t>(J) * N;
        const u32 *pm = perm ? perm + static_cast<size_t>(h) * S : nullptr;
        FOR_THREADS(tid, nt) {
          const KsMacOut mac{k0, k1, G, nt, q, qi, TmemAcc::make(tbase, tid, nt, gpt * 2, emu)};
          for (int g = tid; g < G; g += nt) {
            u32 pj[8];
            double x[8];
            // perm == nullptr: the producer stored the transform already permuted (corr_mac's scatter store): 8 consecutive words
#pragma unroll
            for (int e = 0; e < 8; ++e) pj[e] = pm ? pm[8 * g + e] : static_cast<u32>(h * S + 8 * g + e);
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = u_to_f(rn[pj[e]]);
            mac.group_out(g, x);
          }
        }
        continue;
      }
      const u64 *dig = target + static_cast<size_t>(b) * stride + static_cast<size_t>(J) * N;
      if constexpr (half_fused_first_pass(LOGH, nt)) {
        // folded load + first register pass in registers; the CTA barrier sits between a thread's last digit load and its first
        // store, where it also covers the previous digit's last pass (other warps may still be reading the buffer)
        const MacIO io{k0, k1, G, nt, gpt * 2, q, qi, tbase, emu};
        fwd_half_fused_f64<LOGH, RawU64, MacIO, kKsOut16, MAXT>(fm, twk, q, qi, h, nt, RawU64{dig}, io, CtaBarrier{});
        continue;
      }
      fwd_half_load_f64<LOGH, RawU64, kWarpLocal>(fm, twk, q, qi, h, nt, RawU64{dig});  // ends with a barrier
      // register passes; the last one hands its outputs to KsMacOut::group_out. The functor is rebuilt per thread inside
      // the chain's FOR_THREADS through TmemAcc::make, so pass the ingredients.
      run_passes<kFold>(fm, twk, q, qi, h, nt, k0, k1, tbase, gpt, emu);
      // the next digit's load (or the write-out) overwrites fm: with the warp-local fold it touches only residues this warp
      // itself read in the last pass
      if (kWarpLocal)
        SYNCWARP();
      else
        SYNC();
    }
    // write-out: TMEM -> canonical residues -> global (through shared memory so that stores are coalesced)
    u64 *fo = reinterpret_cast<u64 *>(fm);
    for (int comp = 0; comp < 2; ++comp) {
      FOR_THREADS(tid, nt) {
        const TmemAcc tm = TmemAcc::make(tbase, tid, nt, gpt * 2, emu);
        for (int g = tid; g < G; g += nt) {
          double a[8];
          tm.ld8((g / nt) * 2 + comp, a);
#pragma unroll
          for (int e = 0; e < 8; ++e) fo[pidx(8 * g + e)] = f_canonical(a[e], q, qi);
        }
      }
      SYNC();
      u64 *o = acc + ((static_cast<size_t>(b) * 2 + comp) * K + k) * N + static_cast<size_t>(h) * S;
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) o[i] = fo[pidx(i)];
      }
      SYNC();
    }
#if defined(__CUDA_ARCH__)
    tmem_free_cta(tbase, ncols);
#endif
  }

  // IO functor that rebuilds the per-thread TMEM view from the thread index of the group it is called for
  struct MacIO {
    static constexpr bool kLoad = false, kStore = false, kGroupOut = true;
    const double *k0, *k1;
    int G, nt, slots;
    double q, qi;
    u32 tbase;
    double *emu;
    HD double load(int) const { return 0.0; }
    HD void store(int, double) const {}
    HD void group_out(int g, const double *x) const {
      const KsMacOut mac{k0, k1, G, nt, q, qi, TmemAcc::make(tbase, g % nt, nt, slots, emu)};
      mac.group_out(g, x);
    }
  };
  template <bool FOLD>
  HD void run_passes(double *fm, F64Tw twk, double q, double qi, int h, int nt, const double *k0, const double *k1, u32 tbase, int gpt,
                     double *emu) const {
    constexpr int G = (1 << LOGH) / 8;
    const MacIO io{k0, k1, G, nt, gpt * 2, q, qi, tbase, emu};
    fwd_half_passes_f64<LOGH, MacIO, kKsOut16, MAXT>(fm, twk, q, qi, h, nt, io);
  }
};

// ------------------------------------------------------------------------------------------------------------
// Key-switch digit kernel for SERVICE-SIZED requests (a handful of ciphertexts): the digit loop spread over a thread-block
// cluster. With one ciphertext the kernel above keeps only 2K = 18 CTAs busy, each walking its L = 8 digits one after the other
// (80 us per key switch at N = 16384, 518 of them per PASTA block). Here a cluster of 8 CTAs serves one (item, key limb, half):
// CTA r transforms digit r and multiplies it with its key slice into shared-memory partial products; after a cluster barrier
// every CTA sums one eighth of the residues over the 8 partial products through distributed shared memory and writes it out.
// Same arithmetic, same result (exact integer sums in doubles), one digit's latency instead of eight. FP64 path, L <= 8.
constexpr int kKsSplitCluster = 8;
template <int LOGH>
struct KsDigitsSplitBody {
  static constexpr const char *kName = "ks_digits_split";
  static constexpr int kMaxThreads = 512, kMinBlocks = 1;
  const u64 *target;
  size_t stride;
  const double *key;  // group-major FP64 key [L][2][K][N] (as KsDigitsTmemBody)
  u64 *acc;           // [count][2][K][N], NTT form, canonical
  const DevConsts *C;
  TwRef tw;
  const u64 *reuse;
  size_t reuse_stride;
  const u32 *perm;
  static constexpr size_t smem_bytes() { return (ntt_smem_words(1 << LOGH) + 2 * (static_cast<size_t>(1) << LOGH)) * 8; }
  HD void phase1(int bid, int, unsigned char *smem) const {
    constexpr int nt = half_threads(LOGH);
    constexpr int S = 1 << LOGH, G = S / 8;
    const int N = 2 * S, K = C->K, L = C->L;
    const int J = bid % kKsSplitCluster, u = bid / kKsSplitCluster;
    const int b = u / (2 * K), kh = u % (2 * K), k = kh >> 1, h = kh & 1;
    double *fm = reinterpret_cast<double *>(smem);
    double *part = fm + ntt_smem_words(S);  // [2][S], group-major like the key: residue 8g+e at e*G+g
    const double q = C->qf[k], qi = C->qinvf[k];
    if (J >= L) {
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < 2 * S; i += nt) part[i] = 0.0;
      }
      return;
    }
    if (reuse && J == k) {  // NTT_J(target_J) is the permuted NTT of the source polynomial
      const u64 *rn = reuse + static_cast<size_t>(b) * reuse_stride + static_cast<size_t>(J) * N;
      const u32 *pm = perm ? perm + static_cast<size_t>(h) * S : nullptr;
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) fm[pidx(i)] = u_to_f(rn[pm ? pm[i] : static_cast<u32>(h * S + i)]);
      }
      SYNC();
    } else {
      const F64Tw twk = tw.fwd_f(k);
      fwd_half_transform_f64<LOGH, RawU64, SmemIO, kKsOut16>(fm, twk, q, qi, h, nt,
                                                             RawU64{target + static_cast<size_t>(b) * stride + static_cast<size_t>(J) * N});
    }
    const double *k0 = key + ((static_cast<size_t>(J) * 2 + 0) * K + k) * N + static_cast<size_t>(h) * S;
    const double *k1 = k0 + static_cast<size_t>(K) * N;
    FOR_THREADS(tid, nt) {
      // one CTA per SM here (128 registers per thread): the key values of U residues are requested before the first product, so the
      // multiply-accumulate costs S / (nt U) L2 round trips instead of S / nt
      constexpr int U = 8;
      for (int i0 = tid; i0 < S; i0 += nt * U) {
        double ka[U], kb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = i0 + u * nt < S ? i0 + u * nt : i0;
          ka[u] = k0[idx];
          kb[u] = k1[idx];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = i0 + u * nt;
          if (idx < S) {
            const int e = idx / G, g = idx % G;
            const double v = fm[pidx(8 * g + e)];
            part[idx] = f_mulmod_var(v, ka[u], q, qi);  // |.| <= 1.25 q
            part[S + idx] = f_mulmod_var(v, kb[u], q, qi);
          }
        }
      }
    }
  }
  // peers[r]: shared memory of cluster rank r (own memory included)
  HD void phase2(int bid, int, unsigned char *const *peers) const {
    constexpr int nt = half_threads(LOGH);
    constexpr int S = 1 << LOGH, G = S / 8, W = S / kKsSplitCluster;
    const int N = 2 * S, K = C->K;
    const int r = bid % kKsSplitCluster, u = bid / kKsSplitCluster;
    const int b = u / (2 * K), kh = u % (2 * K), k = kh >> 1, h = kh & 1;
    const double q = C->qf[k], qi = C->qinvf[k];
    FOR_THREADS(tid, nt) {
      for (int t = tid; t < 2 * W; t += nt) {
        const int comp = t / W, idx = r * W + t % W;  // group-major index of this CTA's slice
        double term[kKsSplitCluster], sum = 0.0;
#pragma unroll
        for (int p = 0; p < kKsSplitCluster; ++p)  // the eight distributed-shared-memory loads first, then the sum
          term[p] = (reinterpret_cast<const double *>(peers[p]) + ntt_smem_words(S))[comp * S + idx];
#pragma unroll
        for (int p = 0; p < kKsSplitCluster; ++p) sum = f_add(sum, term[p]);  // <= 10 q: exact
        const int e = idx / G, g = idx % G;
        acc[((static_cast<size_t>(b) * 2 + comp) * K + k) * N + static_cast<size_t>(h) * S + 8 * g + e] = f_canonical(sum, q, qi);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// N = 32768: a limb (256 KiB) no longer fits one SM's shared memory. Transforms are split in two N/2-point
// sub-transforms (one CTA each): forward folds the first Cooley-Tukey stage into the load; inverse leaves the last
// Gentleman-Sande stage (and the 1/N scaling) to InvFinalBody. Integer (wide-slack / Harvey) arithmetic only: the
// BFVDefault(32768) primes are 55-56 bits.
template <int LOGS>  // S = 2^LOGS = N/2
struct NttSplitBody {
  static constexpr const char *kName = "ntt_split";
  const u64 *in;
  u64 *out;
  const DevConsts *C;
  TwRef tw;
  TabMap map;
  int limbs;
  int inverse;
  size_t istride;
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    u64 *sm = reinterpret_cast<u64 *>(smem);
    const int h = bid & 1, lb = bid >> 1;
    const int tab = map.id[lb % limbs];
    const DevMod mm = C->mod[tab];
    const u64 q = mm.q;
    const size_t at = static_cast<size_t>(lb / limbs) * istride + static_cast<size_t>(lb % limbs) * (2 * S);
    const u64 *src = in + at;
    u64 *dst = out + at + static_cast<size_t>(h) * S;
    if (!inverse) {
      const W2 w1 = tw.fwd(tab)[1];
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) {
          const u64 x = src[i], t = mul_shoup_lazy(src[i + S], w1.w, w1.ws, q);  // x < q, t < 2q
          sm[pidx(i)] = h ? x + (q << 1) - t : x + t;                            // < 3q
        }
      }
      SYNC();
      ntt_fwd_core<LOGS>(sm, tw.fwd(tab), q, 2 + h, nt);
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) dst[i] = barrett64(sm[pidx(i)], mm);
      }
    } else {
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) sm[pidx(i)] = src[static_cast<size_t>(h) * S + i];
      }
      SYNC();
      ntt_inv_core<LOGS>(sm, tw.inv(tab), q, 2 + h, nt);
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) dst[i] = sm[pidx(i)];  // lazy, < 2q; finished by InvFinalBody
      }
    }
  }
};

// last Gentleman-Sande stage of the split inverse transform + 1/N: (a, b) = (d[j], d[j + N/2]) -> ((a+b)/N, (a-b) w^-1 / N)
struct InvFinalBody {
  static constexpr const char *kName = "ntt_inv_final";
  u64 *data;
  const DevConsts *C;
  TwRef tw;
  TabMap map;
  int limbs;
  size_t istride;
  size_t total;  // items * limbs * N/2
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t half = C->N >> 1;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t lb = g / half, j = g % half;
        const int tab = map.id[lb % limbs];
        const u64 q = C->mod[tab].q, two_q = q << 1;
        u64 *d = data + (lb / limbs) * istride + (lb % limbs) * (2 * half);
        const W2 w = tw.inv(tab)[1], ninv = C->n_inv[tab];
        const u64 a = d[j], b = d[j + half];  // both < 2q
        const u64 s = csub(a + b, two_q);
        const u64 t = mul_shoup_lazy(a + two_q - b, w.w, w.ws, q);
        d[j] = mul_shoup(s, ninv, q);
        d[j + half] = mul_shoup(t, ninv, q);
      }
    }
  }
};

// Key-switch digit kernel for N = 32768: quarter split (two Cooley-Tukey stages folded into the digit load: 3 Shoup
// products per residue instead of 1, but work buffer + accumulators fit: 68 + 128 KiB). CTA per (item, key limb, quarter).
template <int LOGQ>  // S = 2^LOGQ = N/4
struct KsDigitsQuadBody {
  static constexpr const char *kName = "ks_digits_quad";
  const u64 *target;
  size_t stride;
  const W2 *key;  // [L][2][K][N]
  u64 *acc;       // [count][2][K][N]
  const DevConsts *C;
  TwRef tw;
  int count;
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGQ;
    const int N = 4 * S;
    const int K = C->K, L = C->L;
    const int b = bid / (4 * K);
    const int kc = bid % (4 * K);
    const int k = kc >> 2, chunk = kc & 3, c1 = chunk >> 1, c0 = chunk & 1;
    u64 *sm = reinterpret_cast<u64 *>(smem);
    u64 *acc0 = sm + ntt_smem_words(S);
    u64 *acc1 = acc0 + S;
    const DevMod mk = C->mod[k];
    const u64 q = mk.q, two_q = q << 1;
    const W2 *twk = tw.fwd(k);
    const W2 w1 = twk[1], w2 = twk[2 + c1];
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) acc0[i] = acc1[i] = 0;
    }
    for (int J = 0; J < L; ++J) {
      const u64 *dig = target + static_cast<size_t>(b) * stride + static_cast<size_t>(J) * N;
      const bool reduce = C->mod[J].q > q;
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) {
          // residues i, i + N/4 (this half of stage 0 pairs them with i + N/2, i + 3N/4)
          u64 x0 = dig[i], x1 = dig[i + S], x2 = dig[i + 2 * S], x3 = dig[i + 3 * S];
          if (reduce) {
            x0 = barrett64(x0, mk);
            x1 = barrett64(x1, mk);
            x2 = barrett64(x2, mk);
            x3 = barrett64(x3, mk);
          }
          const u64 t0 = mul_shoup_lazy(x2, w1.w, w1.ws, q), t1 = mul_shoup_lazy(x3, w1.w, w1.ws, q);
          const u64 u0 = c1 ? x0 + two_q - t0 : x0 + t0;  // < 3q
          const u64 u1 = c1 ? x1 + two_q - t1 : x1 + t1;
          const u64 tt = mul_shoup_lazy(u1, w2.w, w2.ws, q);
          const u64 u0r = csub(u0, two_q);                // < 2q
          sm[pidx(i)] = c0 ? u0r + two_q - tt : u0r + tt;  // < 4q
        }
      }
      SYNC();
      ntt_fwd_core<LOGQ>(sm, twk, q, 4 + chunk, nt);
      const W2 *k0 = key + ((static_cast<size_t>(J) * 2 + 0) * K + k) * N + static_cast<size_t>(chunk) * S;
      const W2 *k1 = k0 + static_cast<size_t>(K) * N;
      FOR_THREADS(tid, nt) {
        for (int i = tid; i < S; i += nt) {
          const u64 v = sm[pidx(i)];
          const W2 a = k0[i], c = k1[i];
          u64 s0 = acc0[i] + mul_shoup_lazy(v, a.w, a.ws, q);
          u64 s1 = acc1[i] + mul_shoup_lazy(v, c.w, c.ws, q);
          acc0[i] = s0 >= two_q ? s0 - two_q : s0;
          acc1[i] = s1 >= two_q ? s1 - two_q : s1;
        }
      }
      SYNC();
    }
    u64 *o0 = acc + ((static_cast<size_t>(b) * 2 + 0) * K + k) * N + static_cast<size_t>(chunk) * S;
    u64 *o1 = o0 + static_cast<size_t>(K) * N;
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) {
        o0[i] = barrett64(acc0[i], mk);
        o1[i] = barrett64(acc1[i], mk);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// ModDown with rounding (second half of switch_key_inplace) fused with the final additions:
//   out[c][i] = ( acc[c][i] - (r_c mod q_i) + (half mod q_i) ) * q_sp^-1  +  base_c[i],  r_c = acc[c][sp] + half mod q_sp
// acc is in coefficient form (after the inverse NTT). base1 may be NULL (rotations: component 1 starts from zero).
struct ModDownBody {
  static constexpr const char *kName = "moddown";
  const u64 *acc;    // [count][2][K][N]
  const u64 *base0;  // item b at base0 + b*bstride : [L][N]
  const u64 *base1;
  size_t bstride;
  u64 *out;  // [count][2][L][N]
  const DevConsts *C;
  size_t total;  // count * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const int K = C->K, L = C->L;
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t b = g >> C->logn, j = g & (N - 1);
        const DevMod msp = C->mod[K - 1];
        for (int c = 0; c < 2; ++c) {
          const u64 *a = acc + ((b * 2 + c) * K) * N + j;
          u64 r = a[static_cast<size_t>(K - 1) * N] + C->half_sp;
          r = csub(r, msp.q);
          const u64 *base = c ? base1 : base0;
          for (int i = 0; i < L; ++i) {
            const DevMod mi = C->mod[i];
            u64 ri = msp.q > mi.q ? barrett64(r, mi) : r;
            u64 v = sub_mod(a[static_cast<size_t>(i) * N], ri, mi.q);
            v = add_mod(v, C->half_sp_mod_q[i], mi.q);
            v = mul_shoup(v, C->inv_sp_mod_q[i], mi.q);
            if (base) v = add_mod(v, base[b * bstride + static_cast<size_t>(i) * N + j], mi.q);
            out[((b * 2 + c) * L + i) * N + j] = v;
          }
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// GaloisTool::apply_galois on coefficient-form ciphertexts (seal/util/galois.h:32-33), gather form:
//   out[idx] = +-in[i],  i*elt = idx or idx+N (mod 2N)  <=>  i' = idx*elt^-1 mod 2N, i = i' mod N, sign = i' >= N
struct GaloisBody {
  static constexpr const char *kName = "galois";
  const u64 *in;  // [count][2][L][N]
  u64 *out;
  const DevConsts *C;
  u32 elt_inv;
  size_t total;  // count*2*L*N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t limb = g >> C->logn, idx = g & (N - 1);
        const u64 q = C->mod[static_cast<u32>(limb) % static_cast<u32>(C->L)].q;
        const u64 ip = (idx * elt_inv) & (2 * N - 1);
        u64 v = in[limb * N + (ip & (N - 1))];
        if (ip >= N) v = neg_mod(v, q);
        out[g] = v;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Element-wise ciphertext ops (Evaluator::add / negate / add_plain, seal/evaluator.h:92,118,665)
struct AddBody {  // out = a + b
  static constexpr const char *kName = "add";
  const u64 *a, *b;
  u64 *out;
  const DevConsts *C;
  int limbs;  // limbs per item (size * L)
  size_t total;
  HD void operator()(int bid, int nt, unsigned char *) const {
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const u64 q = C->mod[static_cast<u32>(g >> C->logn) % static_cast<u32>(limbs) % static_cast<u32>(C->L)].q;
        out[g] = add_mod(a[g], b[g], q);
      }
    }
  }
};

struct NegateBody {
  static constexpr const char *kName = "negate";
  const u64 *a;
  u64 *out;
  const DevConsts *C;
  size_t total;
  HD void operator()(int bid, int nt, unsigned char *) const {
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) out[g] = neg_mod(a[g], C->mod[static_cast<u32>(g >> C->logn) % static_cast<u32>(C->L)].q);
    }
  }
};

// out = (negate ? -a : a) + Delta-scaled plaintext on component 0 (multiply_add_plain_with_scaling_variant):
//   c0[j] += m_j*floor(Q/t) + floor((m_j*(Q mod t) + (t+1)/2) / t)
struct AddPlainBody {
  static constexpr const char *kName = "add_plain";
  const u64 *a;   // [count][2][L][N]
  const u64 *pt;  // [count][N], item stride pstride (0 = shared)
  size_t pstride;
  u64 *out;
  const DevConsts *C;
  int negate;
  size_t total;  // count*2*L*N
  const u32 *ptidx;  // optional: item -> plaintext index (blocks sharing a SHAKE counter share round constants)
  const u32 *aidx;   // optional: item -> ciphertext index in `a` (blocks sharing a SHAKE counter share the keystream ciphertext)
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    const int L = C->L;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const u32 limb = static_cast<u32>(g >> C->logn);
        const size_t j = g & (N - 1);
        const int i = static_cast<int>(limb % static_cast<u32>(L));
        const size_t item = limb / static_cast<u32>(2 * L);
        const int comp = static_cast<int>((limb / static_cast<u32>(L)) & 1);
        const DevMod mi = C->mod[i];
        u64 v = aidx ? a[(static_cast<size_t>(aidx[item]) * 2 * L + (limb % static_cast<u32>(2 * L))) * N + j] : a[g];
        if (negate) v = neg_mod(v, mi.q);
        if (comp == 0) {
          const u64 m = pt[(ptidx ? ptidx[item] : item) * pstride + j];
          const u64 fix = (m * C->q_mod_t + C->half_t) / C->t;  // m, Q mod t < 2^32
          v = add_mod(v, mul_add_mod(m, C->q_div_t_mod_q[i], fix, mi), mi.q);
        }
        out[g] = v;
      }
    }
  }
};

struct BroadcastBody {  // out[item] = src for every item
  static constexpr const char *kName = "broadcast";
  const u64 *src;
  u64 *out;
  size_t words, total;
  HD void operator()(int bid, int nt, unsigned char *) const {
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) out[g] = src[g % words];
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// BatchEncoder::encode (seal/batchencoder.h:80): scatter slot values through the index map, inverse NTT mod t.
// One CTA per plaintext. The slot source depends on `mode`:
//   kSlots     : u64 slots[item][sstride], first lens[item] (or n) valid          (mask, symmetric ciphertext block)
//   kDiag      : PASTA diagonal `diag` of layer `layer` from the generated matrices (pasta_3_seal.cpp:388-401)
//   kDiagBsgs  : same, pre-rotated for baby-step/giant-step                         (pasta_3_seal.cpp:281-326)
//   kRc        : round constants of layer `layer`                                   (pasta_3_plain.cpp:286-295)
//   kFeistel   : the constant Feistel mask                                          (pasta_3_seal.cpp:229-235)
enum EncodeMode { kSlots = 0, kDiag = 1, kDiagBsgs = 2, kRc = 3, kFeistel = 4 };

// material layout per block: u32 [4 layers][2 matrices][128][128], then u32 rc[4][256]
constexpr size_t kMatWords = static_cast<size_t>(4) * 2 * kPastaT * kPastaT;
constexpr size_t kMaterialWords = kMatWords + 4 * 2 * kPastaT;

template <int LOGS>
struct EncodeBody {
  static constexpr const char *kName = "encode";
  const u64 *slots;
  size_t sstride;
  const u32 *lens;  // per item valid count (NULL: n)
  u32 n;
  const u32 *material;  // [items][kMaterialWords]
  const u32 *mat_index; // item -> material row (NULL: identity)
  const u32 *index_map;
  u64 *pt;  // [items][N]
  const DevConsts *C;
  TwRef tw;
  int mode, layer, diag0;
  int per;  // kDiag / kDiagBsgs: plaintexts per diagonal (item = g * per + m encodes diagonal diag0 + g from material row m); 0: one diagonal
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    constexpr int T = kPastaT;
    const int diag = diag0 + (per ? bid / per : 0), mrow = per ? bid % per : bid;
    u64 *sm = reinterpret_cast<u64 *>(smem);
    const int tab = 2 * C->K;
    const u64 t = C->t;
    const int half = S / 2;
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) sm[pidx(i)] = 0;
    }
    SYNC();
    const u32 *mat = material ? material + static_cast<size_t>(mat_index ? mat_index[mrow] : mrow) * kMaterialWords : nullptr;
    FOR_THREADS(tid, nt) {
      if (mode == kSlots) {
        const u32 cnt = lens ? lens[bid] : n;
        for (u32 i = tid; i < cnt; i += nt) sm[pidx(index_map[i])] = slots[static_cast<size_t>(bid) * sstride + i];
      } else if (mode == kRc) {
        for (int i = tid; i < 2 * T; i += nt) {
          const int slot = i < T ? i : half + (i - T);
          sm[pidx(index_map[slot])] = mat[kMatWords + layer * 2 * T + i];
        }
      } else if (mode == kFeistel) {
        for (int i = tid; i < 2 * T; i += nt) {
          const int j = i & (T - 1);
          if (j) sm[pidx(index_map[(i < T ? 0 : half) + j])] = 1;
        }
      } else {
        // diagonal `diag`: entry j of matrix M is M[j][(j - diag) mod T]
        for (int i = tid; i < 2 * T; i += nt) {
          const int which = i >= T, p = i & (T - 1);
          const u32 *M = mat + (static_cast<size_t>(layer) * 2 + which) * T * T;
          int slot = p, src = p;
          if (mode == kDiagBsgs) {
            const int shift = (diag / 16) * 16;  // giant-step pre-rotation k*N1
            src = (p + shift) & (T - 1);
            if (half != T && p >= T - shift) slot = p + (half - T);
          }
          const u32 v = M[src * T + ((src + T - diag) & (T - 1))];
          sm[pidx(index_map[which * half + slot])] = v;
        }
      }
    }
    SYNC();
    ntt_inv_core<LOGS>(sm, tw.inv(tab), t, 1, nt);
    const W2 ninv = C->n_inv[tab];
    u64 *dst = pt + static_cast<size_t>(bid) * S;
    FOR_THREADS(tid, nt) {
      for (int i = tid; i < S; i += nt) dst[i] = mul_shoup(sm[pidx(i)], ninv, t);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// N = 32768 (split transforms): the whole-limb kernels encode / lift_ntt / ntt_mac as element-wise pieces around Engine::ntt.
// EncodeScatterBody builds the slot vector of EncodeBody in global memory (the inverse NTT mod t follows as a split transform);
// LiftBody is the centred lift of multiply_plain into every limb; DyadicMacBody the product with the lifted diagonal and the sum.
struct EncodeScatterBody {
  static constexpr const char *kName = "encode";
  const u64 *slots;
  size_t sstride;
  const u32 *lens;
  u32 n;
  const u32 *material;
  const u32 *mat_index;
  const u32 *index_map;
  u64 *pt;  // [items][N]: slot values at their index_map positions, zero elsewhere
  const DevConsts *C;
  int mode, layer, diag0, per;
  HD void operator()(int bid, int nt, unsigned char *) const {
    constexpr int T = kPastaT;
    const size_t N = C->N;
    const int half = static_cast<int>(N / 2);
    const int diag = diag0 + (per ? bid / per : 0), mrow = per ? bid % per : bid;
    u64 *dst = pt + static_cast<size_t>(bid) * N;
    FOR_THREADS(tid, nt) {
      for (size_t i = tid; i < N; i += nt) dst[i] = 0;
    }
    SYNC();
    const u32 *mat = material ? material + static_cast<size_t>(mat_index ? mat_index[mrow] : mrow) * kMaterialWords : nullptr;
    FOR_THREADS(tid, nt) {
      if (mode == kSlots) {
        const u32 cnt = lens ? lens[bid] : n;
        for (u32 i = tid; i < cnt; i += nt) dst[index_map[i]] = slots[static_cast<size_t>(bid) * sstride + i];
      } else if (mode == kRc) {
        for (int i = tid; i < 2 * T; i += nt) dst[index_map[i < T ? i : half + (i - T)]] = mat[kMatWords + layer * 2 * T + i];
      } else if (mode == kFeistel) {
        for (int i = tid; i < 2 * T; i += nt) {
          const int j = i & (T - 1);
          if (j) dst[index_map[(i < T ? 0 : half) + j]] = 1;
        }
      } else {
        for (int i = tid; i < 2 * T; i += nt) {
          const int which = i >= T, p = i & (T - 1);
          const u32 *M = mat + (static_cast<size_t>(layer) * 2 + which) * T * T;
          int slot = p, src = p;
          if (mode == kDiagBsgs) {
            const int shift = (diag / 16) * 16;
            src = (p + shift) & (T - 1);
            if (half != T && p >= T - shift) slot = p + (half - T);
          }
          dst[index_map[which * half + slot]] = M[src * T + ((src + T - diag) & (T - 1))];
        }
      }
    }
  }
};

struct LiftBody {  // out[item][i][j] = centred lift of pt[item][j] into limb i (no lift for monomial plaintexts, see LiftNttBody)
  static constexpr const char *kName = "lift";
  const u64 *pt;  // [items][N]
  u64 *out;       // [items][L][N]
  const DevConsts *C;
  const u32 *nolift;
  size_t total;  // items * L * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    const u32 L = static_cast<u32>(C->L);
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const u32 limb = static_cast<u32>(g >> C->logn), i = limb % L, item = limb / L;
        const u64 m = pt[static_cast<size_t>(item) * N + (g & (N - 1))];
        const bool lift = m >= C->half_t && !(nolift && nolift[item]);
        out[g] = lift ? m + (C->mod[i].q - C->t) : m;
      }
    }
  }
};

struct DyadicMacBody {  // sum (+)= x (.) D   (x: NTT form [items][comps][L][N])
  static constexpr const char *kName = "dyadic_mac1";
  const u64 *x;
  const u64 *D;
  size_t dstride;
  u64 *sum;
  const DevConsts *C;
  int first, comps;
  size_t sum_stride, sum_off;
  const u32 *didx;
  size_t total;  // items * comps * L * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    const u32 L = static_cast<u32>(C->L);
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const u32 limb = static_cast<u32>(g >> C->logn), i = limb % L, cl = limb % (comps * L), item = limb / (comps * L);
        const size_t j = g & (N - 1);
        const DevMod mi = C->mod[i];
        u64 v = mul_mod(x[g], D[(didx ? didx[item] : item) * dstride + static_cast<size_t>(i) * N + j], mi);
        u64 *dst = sum + static_cast<size_t>(item) * sum_stride + sum_off + static_cast<size_t>(cl) * N + j;
        if (!first) v = add_mod(v, *dst, mi.q);
        *dst = v;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Plaintext half of Evaluator::multiply_plain (seal/evaluator.h:729): centred lift of pt into limb i, forward NTT.
// grid = items * L
template <int LOGS>
struct LiftNttBody {
  static constexpr const char *kName = "lift_ntt";
  const u64 *pt;  // [items][N]
  u64 *out;       // [items][L][N]
  const DevConsts *C;
  TwRef tw;
  const u32 *nolift;  // optional, per item: 1 = monomial plaintext, SEAL multiplies by the coefficient without the centred lift
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    u64 *sm = reinterpret_cast<u64 *>(smem);
    const int L = C->L, i = bid % L;
    const size_t item = bid / L;
    const u64 q = C->mod[i].q, inc = q - C->t, thr = (nolift && nolift[item]) ? ~static_cast<u64>(0) : C->half_t;
    const u64 *src = pt + item * S;
    if (C->f64[i]) {
      double *fm = reinterpret_cast<double *>(smem);
      const double qd = C->qf[i], qi = C->qinvf[i];
      ntt_fwd_core_f64<LOGS, 0, 2>(fm, tw.fwd_f(i), qd, qi, 0, nt, LoadLift{src, thr, inc});
      u64 *dstf = out + static_cast<size_t>(bid) * S;
      FOR_THREADS(tid, nt) {
        for (int j = tid; j < S; j += nt) dstf[j] = f_canonical(fm[pidx(j)], qd, qi);
      }
      return;
    }
    FOR_THREADS(tid, nt) {
      for (int j = tid; j < S; j += nt) {
        const u64 m = src[j];
        sm[pidx(j)] = m >= thr ? m + inc : m;
      }
    }
    SYNC();
    ntt_fwd_core<LOGS>(sm, tw.fwd(i), q, 1, nt);
    u64 *dst = out + static_cast<size_t>(bid) * S;
    const DevMod mm = C->mod[i];
    FOR_THREADS(tid, nt) {
      for (int j = tid; j < S; j += nt) dst[j] = barrett64(sm[pidx(j)], mm);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Ciphertext half of multiply_plain, accumulated in the NTT domain:
//   sum[b][c][i] (+)= NTT_i(ct[b][c][i]) (.) D[b][i]          grid = items * 2 * L
// (the PASTA diagonal loop sums 128 such products; accumulation before the single inverse NTT is exact)
template <int LOGS>
struct NttMacBody {
  static constexpr const char *kName = "ntt_mac";
  const u64 *ct;  // [items][2][L][N] coefficient form
  const u64 *D;   // [items][L][N] (dstride = L*N) or shared (dstride = 0)
  size_t dstride;
  u64 *sum;
  const DevConsts *C;
  TwRef tw;
  int first;  // 1: overwrite
  int comps;  // polynomials per item in `ct` (2: whole ciphertexts, 1: one component)
  size_t sum_stride, sum_off;  // sum limb (item, c, i) lives at sum + item*sum_stride + sum_off + (c*L + i)*N
  u64 *ntt_out;                // optional: NTT_i(ct) itself (canonical), same indexing as ct
  const u32 *didx;  // optional: item -> diagonal index (blocks sharing a SHAKE counter share their diagonals)
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    u64 *sm = reinterpret_cast<u64 *>(smem);
    const int L = C->L, i = bid % L;
    const size_t item = bid / (comps * L);
    const int cl = bid % (comps * L);
    const DevMod mi = C->mod[i];
    const u64 q = mi.q;
    const u64 *src = ct + static_cast<size_t>(bid) * S;
    const u64 *d = D + (didx ? didx[item] : item) * dstride + static_cast<size_t>(i) * S;
    u64 *dst = sum + item * sum_stride + sum_off + static_cast<size_t>(cl) * S;
    u64 *nout = ntt_out ? ntt_out + static_cast<size_t>(bid) * S : nullptr;
    if (C->f64[i]) {
      double *fm = reinterpret_cast<double *>(smem);
      const double qd = C->qf[i], qi = C->qinvf[i];
      ntt_fwd_core_f64<LOGS, 0, 2>(fm, tw.fwd_f(i), qd, qi, 0, nt, LoadU64{src});
      FOR_THREADS(tid, nt) {
        for (int j = tid; j < S; j += nt) {
          if (nout) nout[j] = f_canonical(fm[pidx(j)], qd, qi);
          u64 v = f_canonical(f_mulmod_var(fm[pidx(j)], u_to_f(d[j]), qd, qi), qd, qi);
          if (!first) v = add_mod(v, dst[j], q);
          dst[j] = v;
        }
      }
      return;
    }
    FOR_THREADS(tid, nt) {
      for (int j = tid; j < S; j += nt) sm[pidx(j)] = src[j];
    }
    SYNC();
    ntt_fwd_core<LOGS>(sm, tw.fwd(i), q, 1, nt);
    FOR_THREADS(tid, nt) {
      for (int j = tid; j < S; j += nt) {
        if (nout) nout[j] = barrett64(sm[pidx(j)], mi);
        u64 v = mul_mod(sm[pidx(j)], d[j], mi);
        if (!first) v = add_mod(v, dst[j], q);
        dst[j] = v;
      }
    }
  }
};

// The inner sum of one giant step of PASTA_SEAL::babystep_giantstep in one pass: inner = sum_{j < J} rot_j (.) D_j with the J baby
// rotations and the J lifted diagonals resident in HBM (pasta_3_seal.cpp:349-357: multiply_plain + add_inplace per j). Each residue
// of `inner` is produced by one thread and written once: 16 x (2 + 1) MiB read + 2 MiB written per block instead of 16 x 7 MiB.
struct DyadicMacNBody {
  static constexpr const char *kName = "dyadic_mac";
  const u64 *rot;  // [J][items][2][L][N] NTT form
  const u64 *D;    // [J][nd][L][N]
  u64 *inner;      // [items][2][L][N]
  const DevConsts *C;
  int J;
  size_t rstride;  // words between consecutive j in rot (items * 2 * L * N)
  size_t jstride;  // words between consecutive j in D (nd * L * N)
  size_t dstride;  // L * N
  const u32 *didx;  // optional: item -> diagonal index
  size_t total;     // items * 2 * L * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    const u32 L = static_cast<u32>(C->L);
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const u32 limb = static_cast<u32>(g >> C->logn);
        const u32 i = limb % L, item = limb / (2 * L);
        const u64 *dp = D + static_cast<size_t>(didx ? didx[item] : item) * dstride + static_cast<size_t>(i) * N + (g & (N - 1));
        const u64 *rp = rot + g;
        if (C->f64[i]) {  // |each product| <= 0.7q, 16 of them stay far below 2^53
          const double q = C->qf[i], qi = C->qinvf[i];
          double acc = 0.0;
          for (int j0 = 0; j0 < J; j0 += 4) {
            u64 rv[4], dv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = j0 + u < J ? j0 + u : j0;
              rv[u] = rp[static_cast<size_t>(j) * rstride];
              dv[u] = dp[static_cast<size_t>(j) * jstride];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (j0 + u < J) acc = f_add(acc, f_mulmod_var(u_to_f(rv[u]), u_to_f(dv[u]), q, qi));
          }
          inner[g] = f_canonical(acc, q, qi);
        } else {
          const DevMod mi = C->mod[i];
          u64 acc = 0;
          for (int j = 0; j < J; ++j)
            acc = add_mod(acc, mul_mod(rp[static_cast<size_t>(j) * rstride], dp[static_cast<size_t>(j) * jstride], mi), mi.q);
          inner[g] = acc;
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// BEHZ multiplication (Evaluator::bfv_multiply, seal/evaluator.h:214; RNSTool, seal/util/rns.h:213-228)

// fastbconv_m_tilde + sm_mrq: x (base q, coefficient form) -> x in base Bsk.  One thread per (poly, coefficient).
struct BehzExtendBody {
  static constexpr const char *kName = "behz_extend";
  const u64 *x;  // [polys][L][N]
  u64 *xb;       // [polys][L+1][N]
  const DevConsts *C;
  size_t total;  // polys * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const int L = C->L;
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t p = g >> C->logn, j = g & (N - 1);
        u64 z[kMaxLimbs];
        u32 mt = 0;
        for (int i = 0; i < L; ++i) {
          z[i] = mul_shoup(x[(p * L + i) * N + j], C->mtilde_ipq[i], C->mod[i].q);
          mt += static_cast<u32>(z[i]) * C->q2mt[i];
        }
        const u32 r = mt * C->neg_inv_q_mt;
        for (int bb = 0; bb <= L; ++bb) {
          const DevMod mb = C->mod[C->K + bb];
          Acc128 s;  // one reduction per output residue (modarith.h)
          for (int i = 0; i < L; ++i) s.mac(z[i], C->q2bsk[bb][i]);
          const u64 rp = r >= 0x80000000u ? mb.q - (0x100000000ULL - r) : r;
          s.mac(rp, C->q_mod_bsk[bb]);
          xb[(p * (L + 1) + bb) * N + j] = mul_shoup(s.reduce(mb), C->inv_mt_bsk[bb], mb.q);
        }
      }
    }
  }
};

// NTT-domain tensor product for one base: d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1.
// a, b: [items][2][limbs][N]; d: [items][3][limbs][N]; table id of limb l is tab0 + l.
struct TensorBody {
  static constexpr const char *kName = "behz_tensor";
  const u64 *a, *b;
  u64 *d;
  const DevConsts *C;
  int limbs, tab0;
  size_t total;  // items * limbs * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t j = g & (N - 1);
        const u32 lidx = static_cast<u32>(g >> C->logn);
        const size_t l = lidx % static_cast<u32>(limbs), item = lidx / static_cast<u32>(limbs);
        const DevMod m = C->mod[tab0 + l];
        const size_t poly = static_cast<size_t>(limbs) * N, off = l * N + j;
        const u64 a0 = a[item * 2 * poly + off], a1 = a[item * 2 * poly + poly + off];
        const u64 b0 = b[item * 2 * poly + off], b1 = b[item * 2 * poly + poly + off];
        u64 *o = d + item * 3 * poly + off;
        o[0] = mul_mod(a0, b0, m);
        o[poly] = add_mod(mul_mod(a0, b1, m), mul_mod(a1, b0, m), m.q);
        o[2 * poly] = mul_mod(a1, b1, m);
      }
    }
  }
};

// multiply by t, fast_floor, fastbconv_sk: (d in base q, d in base Bsk; coefficient form) -> round(t*d/Q) in base q
struct BehzScaleRoundBody {
  static constexpr const char *kName = "behz_scale_round";
  const u64 *dq;  // [polys][L][N]
  const u64 *db;  // [polys][L+1][N]
  u64 *out;       // [polys][L][N]
  const DevConsts *C;
  size_t total;  // polys * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const int L = C->L, K = C->K;
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t p = g >> C->logn, j = g & (N - 1);
        u64 z[kMaxLimbs], f[kMaxLimbs] = {};
        for (int i = 0; i < L; ++i) z[i] = mul_shoup(dq[(p * L + i) * N + j], C->t_ipq[i], C->mod[i].q);
        for (int bb = 0; bb <= L; ++bb) {
          const DevMod mb = C->mod[K + bb];
          Acc128 s;  // one reduction per inner product (modarith.h)
          for (int i = 0; i < L; ++i) s.mac(z[i], C->q2bsk[bb][i]);
          const u64 tb = mul_shoup(db[(p * (L + 1) + bb) * N + j], C->t_mod_bsk[bb], mb.q);
          f[bb] = mul_shoup(sub_mod(tb, s.reduce(mb), mb.q), C->inv_q_bsk[bb], mb.q);
        }
        const u64 f_sk = f[L];
        const DevMod msk = C->mod[K + L];
        Acc128 sk;
        for (int i = 0; i < L; ++i) {
          f[i] = mul_shoup(f[i], C->inv_punct_b[i], C->mod[K + i].q);
          sk.mac(f[i], C->b2msk[i]);
        }
        const u64 alpha = mul_shoup(sub_mod(sk.reduce(msk), f_sk, msk.q), C->inv_pb_msk, msk.q);
        const bool upper = alpha > (msk.q >> 1);
        for (int jq = 0; jq < L; ++jq) {
          const DevMod mq = C->mod[jq];
          Acc128 s;
          for (int i = 0; i < L; ++i) s.mac(f[i], C->b2q[jq][i]);
          if (upper)
            s.mac(msk.q - alpha, C->pb_mod_q[jq]);
          else
            s.mac(alpha, mq.q - C->pb_mod_q[jq]);
          out[(p * L + jq) * N + j] = s.reduce(mq);
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Pieces of the NTT-resident rotation chain of the PASTA diagonal loop (Engine::affine_diagonal_resident).
// Component 0 of the rotating state is kept in the NTT domain for the whole layer:
//   c0' = galois(c0) + k0,  k0 = (INTT(acc0) - corr) * q_sp^-1,  corr[j] = (r0[j] mod q_i) - (half mod q_i)
//   =>  NTT(c0') = perm(NTT(c0)) + (acc0 - NTT(corr)) * q_sp^-1          (perm = the automorphism on NTT slots)
// which needs one inverse NTT (special limb) + L forward NTTs of corr instead of K inverse + L forward NTTs, and the
// plaintext product for component 0 becomes element-wise. Exact: every step is linear over Z_q_i.

// rows x words strided copy
struct StridedCopyBody {
  static constexpr const char *kName = "strided_copy";
  const u64 *src;
  u64 *dst;
  size_t sstride, dstride, words, total;  // total = rows * words
  HD void operator()(int bid, int nt, unsigned char *) const {
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t r = g / words, w = g % words;
        dst[r * dstride + w] = src[r * sstride + w];
      }
    }
  }
};

// dst[r][l][j] = src[r][l][perm[j]]: a limb-wise slot permutation while copying (start of the resident chain: the transforms of the
// state are stored the way the first rotation will read them)
struct PermCopyBody {
  static constexpr const char *kName = "perm_copy";
  const u64 *src;
  u64 *dst;
  const u32 *perm;
  size_t sstride, dstride;  // words between rows
  const DevConsts *C;
  int limbs;
  size_t total;  // rows * limbs * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t j = g & (N - 1), limb = g >> C->logn, r = limb / limbs, l = limb % limbs;
        dst[r * dstride + l * N + j] = src[r * sstride + l * N + perm[j]];
      }
    }
  }
};

// ModDown of component 1 only: c1[i][j] = (acc1[i][j] - (r1[j] mod q_i) + half_i) * q_sp^-1, acc1 in coefficient form
struct ModDownC1Body {
  static constexpr const char *kName = "moddown_c1";
  const u64 *acc;  // [count][2][K][N]
  u64 *out;        // [count][L][N]
  const DevConsts *C;
  size_t total;  // count * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const int K = C->K, L = C->L;
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t b = g >> C->logn, j = g & (N - 1);
        const DevMod msp = C->mod[K - 1];
        const u64 *a = acc + ((b * 2 + 1) * K) * N + j;
        const u64 r = csub(a[static_cast<size_t>(K - 1) * N] + C->half_sp, msp.q);
        for (int i = 0; i < L; ++i) {
          const DevMod mi = C->mod[i];
          const u64 ri = msp.q > mi.q ? barrett64(r, mi) : r;
          u64 v = add_mod(sub_mod(a[static_cast<size_t>(i) * N], ri, mi.q), C->half_sp_mod_q[i], mi.q);
          out[(b * L + i) * N + j] = mul_shoup(v, C->inv_sp_mod_q[i], mi.q);
        }
      }
    }
  }
};

// NTT of the component-0 correction + update of the NTT-resident c0 + plaintext product. CTA per (item, limb i).
template <int LOGS>
struct Corr0MacBody {
  static constexpr const char *kName = "corr0_mac";
  const u64 *acc;     // [items][2][K][N]: [0][K-1] coefficient form (after the inverse NTT), [0][i<L] NTT form
  const u64 *c0_in;   // [items][L][N] NTT form
  u64 *c0_out;        // [items][L][N]
  const u32 *perm;    // NTT-slot permutation of the Galois element
  const u64 *D;       // [items][L][N], or shared by all items (dstride = 0)
  u64 *sum;           // [items][2][L][N], component 0 updated
  const DevConsts *C;
  TwRef tw;
  size_t dstride;     // L*N or 0
  const u32 *didx;    // optional: item -> diagonal index
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    const int L = C->L, K = C->K, i = bid % L;
    const size_t item = bid / L;
    const DevMod mi = C->mod[i], msp = C->mod[K - 1];
    const u64 q = mi.q;
    const u64 *sp = acc + ((item * 2) * K + (K - 1)) * S;
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[i], qi = C->qinvf[i];
    ntt_fwd_core_f64<LOGS, 0, 2>(fm, tw.fwd_f(i), qd, qi, 0, nt, LoadCorr{sp, C->half_sp, C->half_sp_mod_q[i], q, mi, msp});
    const u64 *a0 = acc + ((item * 2) * K + i) * S;
    const u64 *cin = c0_in + (item * L + i) * S;
    u64 *cout = c0_out + (item * L + i) * S;
    const u64 *d = D + (didx ? didx[item] : item) * dstride + static_cast<size_t>(i) * S;
    u64 *s0 = sum + (item * 2 * L + i) * S;
    FOR_THREADS(tid, nt) {
      constexpr int U = 4;  // independent gathers in flight per thread (perm -> c0 is a dependent pair of loads)
      for (int j0 = tid; j0 < S; j0 += nt * U) {
        u32 pj[U];
        u64 cv[U], av[U], dv[U], sv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) pj[u] = perm[j0 + u * nt < S ? j0 + u * nt : j0];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt < S ? j0 + u * nt : j0;
          cv[u] = cin[pj[u]];
          av[u] = a0[j];
          dv[u] = d[j];
          sv[u] = s0[j];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt;
          if (j < S) {
            const u64 t = f_canonical(fm[pidx(j)], qd, qi);
            const u64 k0 = mul_shoup(sub_mod(av[u], t, q), C->inv_sp_mod_q[i], q);
            const u64 c = add_mod(cv[u], k0, q);
            cout[j] = c;
            s0[j] = add_mod(sv[u], mul_mod(c, dv[u], mi), q);
          }
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Half-limb versions (FP64 path, two CTAs per SM) of lift_ntt / ntt_mac / corr0_mac: CTA = (limb, half).
template <int LOGH>
struct LiftNttHalfBody {
  static constexpr const char *kName = "lift_ntt";
  static constexpr int kMaxThreads = 512, kMinBlocks = 2;
  const u64 *pt;  // [items][N]
  u64 *out;       // [items][L][N]
  const DevConsts *C;
  TwRef tw;
  const u32 *nolift;  // optional, per item: 1 = monomial plaintext (no centred lift, see LiftNttBody)
  HD void operator()(int bid, int, unsigned char *smem) const {
    constexpr int nt = half_threads(LOGH);
    constexpr int S = 1 << LOGH;
    const int h = bid & 1, lb = bid >> 1;
    const int L = C->L, i = lb % L;
    const size_t item = lb / L;
    const u64 q = C->mod[i].q;
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[i], qi = C->qinvf[i];
    const F64Tw twk = tw.fwd_f(i);
    fwd_half_transform_f64<LOGH>(fm, twk, qd, qi, h, nt,
                                 RawLift{pt + item * (2 * S), (nolift && nolift[item]) ? ~static_cast<u64>(0) : C->half_t, q - C->t});
    u64 *dst = out + static_cast<size_t>(lb) * (2 * S) + static_cast<size_t>(h) * S;
    FOR_THREADS(tid, nt) {
#pragma unroll 4
      for (int j = tid; j < S; j += nt) dst[j] = f_canonical(fm[pidx(j)], qd, qi);
    }
  }
};

template <int LOGH>
struct NttMacHalfBody {
  static constexpr const char *kName = "ntt_mac";
  static constexpr int kMaxThreads = 512, kMinBlocks = 2;
  const u64 *ct;  // [items][comps][L][N] coefficient form
  const u64 *D;   // [items][L][N] (dstride = L*N) or shared (dstride = 0)
  size_t dstride;
  u64 *sum;
  const DevConsts *C;
  TwRef tw;
  int first;  // 1: overwrite
  int comps;
  size_t sum_stride, sum_off;
  u64 *ntt_out;  // optional: NTT_i(ct) itself (canonical), same indexing as ct
  const u32 *didx;  // optional: item -> diagonal index (blocks sharing a SHAKE counter share their diagonals)
  int pf, nl;       // L2 prefetch distance in limbs (0: off), total limbs of the launch
  HD void operator()(int bid, int, unsigned char *smem) const {
    constexpr int nt = half_threads(LOGH);
    constexpr int S = 1 << LOGH;
    const int h = bid & 1, lb = bid >> 1;
    const int L = C->L, i = lb % L;
    const size_t item = lb / (comps * L);
    const int cl = lb % (comps * L);
    const size_t hoff = static_cast<size_t>(h) * S;
    if (pf && !h && lb + pf < nl) {  // operands of the CTA pair that runs two waves later: its limb, diagonal limb and running sum
      const int lf = lb + pf;
      const size_t itf = lf / (comps * L);
      cta_prefetch_l2(ct + static_cast<size_t>(lf) * (2 * S), sizeof(u64) * 2 * S);
      if (dstride) cta_prefetch_l2(D + (didx ? didx[itf] : itf) * dstride + static_cast<size_t>(lf % L) * (2 * S), sizeof(u64) * 2 * S);
      if (!first) cta_prefetch_l2(sum + itf * sum_stride + sum_off + static_cast<size_t>(lf % (comps * L)) * (2 * S), sizeof(u64) * 2 * S);
    }
    const u64 *d = D + (didx ? didx[item] : item) * dstride + static_cast<size_t>(i) * (2 * S) + hoff;
    u64 *dst = sum + item * sum_stride + sum_off + static_cast<size_t>(cl) * (2 * S) + hoff;
    u64 *nout = ntt_out ? ntt_out + static_cast<size_t>(lb) * (2 * S) + hoff : nullptr;
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[i], qi = C->qinvf[i];
    const F64Tw twk = tw.fwd_f(i);
    fwd_half_transform_f64<LOGH>(fm, twk, qd, qi, h, nt, RawU64{ct + static_cast<size_t>(lb) * (2 * S)});
    FOR_THREADS(tid, nt) {
      constexpr int U = 4;
      for (int j0 = tid; j0 < S; j0 += nt * U) {
        u64 dv[U], sv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt < S ? j0 + u * nt : j0;
          dv[u] = d[j];
          sv[u] = first ? 0 : dst[j];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt;
          if (j < S) {
            const double x = fm[pidx(j)];
            if (nout) nout[j] = f_canonical(x, qd, qi);
            // |x| <= 10q: the product is bounded by 2.4q, plus the canonical running sum
            dst[j] = f_canonical(f_add(f_mulmod_var(x, u_to_f(dv[u]), qd, qi), u_to_f(sv[u])), qd, qi);
          }
        }
      }
    }
  }
};

// Both components of the rotated ciphertext, NTT-resident, in one launch (CTA = (item, limb i, component c, half h)):
//   c = 0:  c0n'[i] = perm(c0n[i]) + (acc0[i] - NTT_i(corr_i(r_0))) * q_sp^-1,   sum0[i] += c0n'[i] (.) D[i]
//   c = 1:  c1n'[i] =                (acc1[i] - NTT_i(corr_i(r_1))) * q_sp^-1,   sum1[i] += c1n'[i] (.) D[i]
// with corr_i(r)[j] = (r[j] mod q_i) - (half mod q_i), r_c = INTT(acc_c[special]) + half mod q_sp. The second line is the same
// identity as the first: NTT_i is linear, so NTT_i of the ModDown output (INTT_i(acc1[i]) - corr_i) * q_sp^-1 needs only the
// transform of the correction. It replaces the forward transform of the coefficient-form c1 (ntt_mac on c1c), so the chain
// never stores or re-reads c1 in coefficient form (only its Galois image, the next key switch's digits), and the two
// CTAs that need the diagonal's limb D[i] run next to each other (the second read is an L2 hit).
// comps = 1 runs component 0 only (the round-1 corr0_mac).
template <int LOGH>
struct Corr0MacHalfBody {
  static constexpr const char *kName = "corr_mac";
  static constexpr int kMaxThreads = 512, kMinBlocks = 2;
  const u64 *acc;     // [items][2][K][N]: [c][K-1] coefficient form (after the inverse NTT), [c][i<L] NTT form
  const u64 *c0_in;   // [items][L][N] NTT form
  u64 *c0_out;        // [items][L][N]
  const u32 *perm;    // NTT-slot permutation of the Galois element
  const u64 *D;       // [items][L][N], or shared by all items (dstride = 0)
  u64 *sum;           // [items][2][L][N]
  const DevConsts *C;
  TwRef tw;
  size_t dstride;     // L*N or 0
  const u32 *didx;    // optional: item -> diagonal index
  int pf, nl;         // L2 prefetch distance in limbs (0: off), total limbs of the launch
  int comps;          // 1: component 0 only; 2: both
  u64 *c1_out;        // [items][L][N] NTT form of the new component 1 (comps == 2)
  size_t cin_stride, cout_stride;  // words between items in c0_in and in c0_out / c1_out (L*N when they are dense per component)
  // pinv != nullptr: c0_in is stored ALREADY PERMUTED (read linearly) and both outputs are stored permuted for their next reader
  // (scatter through the inverse permutation): the dependent index -> gather pair of loads leaves the epilogue's critical path, and
  // ks_digits reads its reused digit as consecutive words. A store does not stall anybody; a gather stalls its warp for two round trips.
  const u32 *pinv;
  // D == nullptr: no plaintext product (the baby rotations of the BSGS layer only want the rotated ciphertext in NTT form)
  HD void operator()(int bid, int, unsigned char *smem) const {
    constexpr int nt = half_threads(LOGH);
    constexpr int S = 1 << LOGH;
    const int h = bid & 1, c = comps == 2 ? (bid >> 1) & 1 : 0, lb = comps == 2 ? bid >> 2 : bid >> 1;
    const int L = C->L, K = C->K, i = lb % L;
    const size_t item = lb / L, N = 2 * S, hoff = static_cast<size_t>(h) * S;
    if (pf && !h && !c && lb + pf < nl) {  // operands of the CTAs that run two waves later
      const int lf = lb + pf, fi = lf % L;
      const size_t itf = lf / L;
      if (fi == 0) cta_prefetch_l2(acc + ((itf * 2) * K + (K - 1)) * N, sizeof(u64) * N);
      cta_prefetch_l2(acc + ((itf * 2) * K + fi) * N, sizeof(u64) * N);
      cta_prefetch_l2(c0_in + itf * cin_stride + static_cast<size_t>(fi) * N, sizeof(u64) * N);
      if (D && dstride) cta_prefetch_l2(D + (didx ? didx[itf] : itf) * dstride + static_cast<size_t>(fi) * N, sizeof(u64) * N);
      if (D) cta_prefetch_l2(sum + (itf * 2 * L + fi) * N, sizeof(u64) * N);
    }
    const DevMod mi = C->mod[i], msp = C->mod[K - 1];
    const u64 *sp = acc + ((item * 2 + c) * K + (K - 1)) * N;
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[i], qi = C->qinvf[i];
    const F64Tw twk = tw.fwd_f(i);
    const u64 *a0 = acc + ((item * 2 + c) * K + i) * N + hoff;
    const u64 *cin = c0_in + item * cin_stride + static_cast<size_t>(i) * N;
    u64 *cout = (c ? c1_out : c0_out) + item * cout_stride + static_cast<size_t>(i) * N + hoff;
    const bool mac = D != nullptr;
    const u64 *d = mac ? D + (didx ? didx[item] : item) * dstride + static_cast<size_t>(i) * N + hoff : nullptr;
    u64 *s0 = mac ? sum + ((item * 2 + c) * L + i) * N + hoff : nullptr;
    fwd_half_transform_f64<LOGH, RawCorr, SmemIO, kF64AnyOut16, 512, false>(fm, twk, qd, qi, h, nt,
                                                                            RawCorr{sp, C->half_sp, C->half_sp_mod_q[i], mi.q, mi, msp});
    const u32 *pm = perm + hoff;
    const D2 isp = C->inv_sp_f[i];
    FOR_THREADS(tid, nt) {
      constexpr int U = 4;  // independent gathers in flight per thread (perm -> c0 is a dependent pair of loads)
      for (int j0 = tid; j0 < S; j0 += nt * U) {
        u32 pj[U];
        u64 cv[U], av[U], dv[U], sv[U];
        if (pinv) {
#pragma unroll
          for (int u = 0; u < U; ++u) pj[u] = pinv[hoff + (j0 + u * nt < S ? j0 + u * nt : j0)];  // where this output is stored
        } else if (!c) {
#pragma unroll
          for (int u = 0; u < U; ++u) pj[u] = pm[j0 + u * nt < S ? j0 + u * nt : j0];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt < S ? j0 + u * nt : j0;
          cv[u] = c ? 0 : (pinv ? cin[hoff + j] : cin[pj[u]]);
          av[u] = a0[j];
          dv[u] = mac ? d[j] : 0;
          sv[u] = mac ? s0[j] : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt;
          if (j < S) {
            // k = (acc - NTT(corr)) * q_sp^-1: |acc - NTT(corr)| <= 11q, |k| <= 2.6q
            const double k0 = f_mulmod_const(f_add(u_to_f(av[u]), -fm[pidx(j)]), isp, qd);
            const u64 cc = f_canonical(f_add(u_to_f(cv[u]), k0), qd, qi);
            if (pinv)
              (cout - hoff)[pj[u]] = cc;
            else
              cout[j] = cc;
            if (mac) s0[j] = f_canonical(f_add(f_mulmod_var(u_to_f(cc), u_to_f(dv[u]), qd, qi), u_to_f(sv[u])), qd, qi);
          }
        }
      }
    }
  }
};

// Inverse NTT of the component-1 accumulator limbs of a key switch with the rounding ModDown and the Galois map of the
// NEXT rotation fused into the store (NTT-resident rotation chain, FP64 path). CTA per (item, limb i < L):
//   c1[i][j]       = (INTT(acc1[i])[j] - ((r1[j] + half) mod q_sp) + half_i) * q_sp^-1  mod q_i     (switch_key_inplace)
//   g1[i][pi(j)]   = +- c1[i][j]   with pi(j) = j * elt mod N, sign from bit log N of j * elt        (apply_galois)
// r1 = acc1[special] must already be in coefficient form (a two-limb launch of NttBody precedes this kernel).
struct StoreModDownGalois {
  static constexpr bool kLoad = false, kStore = true, kGroupOut = false;
  const u64 *sp;
  u64 *c1, *g1;
  D2 ninv, isp;
  double q, qinv, qsp, half_sp, half_i;
  u64 qu;
  u32 elt, nmask;
  int logn;
  HD double load(int) const { return 0.0; }
  HD void group_out(int, const double *) const {}
  struct Aux {
    u64 s;
  };
  HD Aux aux(int j) const { return Aux{sp[j]}; }
  HD void store(int j, double v) const { store(j, v, aux(j)); }
  HD void store(int j, double v, const Aux &ax) const {
    const double x = f_mulmod_const(v, ninv, q);
    double t = f_add(u_to_f(ax.s), half_sp);
    if (t >= qsp) t = f_add(t, -qsp);
    const double y = f_add(f_add(x, -t), half_i);  // |y| <= 1.4q + q_sp + q/2 < 5q
    const u64 c = f_canonical(f_mulmod_const(y, isp, q), q, qinv);
    if (c1) c1[j] = c;  // the coefficient form itself is only kept where a caller needs it
    const u32 raw = static_cast<u32>(j) * elt;
    g1[raw & nmask] = ((raw >> logn) & 1) && c ? qu - c : c;
  }
};

template <int LOGS>
struct InttModDownBody {
  static constexpr const char *kName = "intt_moddown";
  const u64 *acc;  // [items][2][K][N]: [1][i<L] NTT form, [1][K-1] coefficient form
  u64 *c1;         // [items][L][N] coefficient form
  u64 *g1;         // [items][L][N] Galois image of c1
  const DevConsts *C;
  TwRef tw;
  u32 elt;
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    const int L = C->L, K = C->K, i = bid % L;
    const size_t item = bid / L;
    const u64 *src = acc + ((item * 2 + 1) * K + i) * S;
    const u64 *sp = acc + ((item * 2 + 1) * K + (K - 1)) * S;
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[i], qi = C->qinvf[i];
    FOR_THREADS(tid, nt) {
      for (int j = tid; j < S; j += nt) fm[pidx(j)] = u_to_f(src[j]);
    }
    SYNC();
    const size_t o = (item * L + i) * S;
    const StoreModDownGalois st{sp, c1 + o, g1 + o, C->n_inv_f[i], C->inv_sp_f[i], qd, qi, C->qf[K - 1], static_cast<double>(C->half_sp),
                                static_cast<double>(C->half_sp_mod_q[i]), C->mod[i].q, elt, static_cast<u32>(S - 1), LOGS};
    ntt_inv_core_f64<LOGS, 0, StoreModDownGalois, whole_threads(LOGS)>(fm, tw.inv_f(i), qd, qi, 0, nt, st);
  }
};

// Generic key switch (rotate / relinearize outside the resident chain), FP64 path: inverse NTT of the accumulator limbs
// acc[c][i < L] with the rounding ModDown and the final add fused into the store. CTA per (item, component c, limb i):
//   out[c][i][j] = base_c[i][j] + (INTT(acc[c][i])[j] - ((r_c[j] + half) mod q_sp) + half_i) * q_sp^-1  mod q_i
// r_c = acc[c][special] must already be in coefficient form (a two-limb launch of NttBody precedes this kernel).
struct StoreModDownAdd {
  static constexpr bool kLoad = false, kStore = true, kGroupOut = false;
  const u64 *sp, *base;
  u64 *out;
  D2 ninv, isp;
  double q, qinv, qsp, half_sp, half_i;
  const u64 *add2;  // optional second addend with out's indexing (a running sum the result is accumulated into; may alias out)
  HD double load(int) const { return 0.0; }
  HD void group_out(int, const double *) const {}
  struct Aux {
    u64 s, b, a2;
  };
  HD Aux aux(int j) const { return Aux{sp[j], base ? base[j] : 0, add2 ? add2[j] : 0}; }
  HD void store(int j, double v) const { store(j, v, aux(j)); }
  HD void store(int j, double v, const Aux &ax) const {
    const double x = f_mulmod_const(v, ninv, q);
    double t = f_add(u_to_f(ax.s), half_sp);
    if (t >= qsp) t = f_add(t, -qsp);
    const double y = f_add(f_add(x, -t), half_i);  // |y| < 5q
    double c = f_mulmod_const(y, isp, q);          // |c| <= 1.5q
    if (base) c = f_add(c, u_to_f(ax.b));
    if (add2) c = f_add(c, u_to_f(ax.a2));         // <= 3.5q
    out[j] = f_canonical(c, q, qinv);
  }
};

template <int LOGS>
struct InttModDownAddBody {
  static constexpr const char *kName = "intt_moddown";
  const u64 *acc;    // [items][2][K][N]
  const u64 *base0;  // item b at base0 + b*bstride : [L][N] (may be null)
  const u64 *base1;
  size_t bstride;
  u64 *out;  // [items][2][L][N]
  const DevConsts *C;
  TwRef tw;
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int S = 1 << LOGS;
    const int L = C->L, K = C->K, i = bid % L, c = (bid / L) & 1;
    const size_t item = bid / (2 * L);
    const u64 *src = acc + ((item * 2 + c) * K + i) * S;
    const u64 *sp = acc + ((item * 2 + c) * K + (K - 1)) * S;
    const u64 *bs = c ? base1 : base0;
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[i], qi = C->qinvf[i];
    FOR_THREADS(tid, nt) {
      for (int j = tid; j < S; j += nt) fm[pidx(j)] = u_to_f(src[j]);
    }
    SYNC();
    const StoreModDownAdd st{sp, bs ? bs + item * bstride + static_cast<size_t>(i) * S : nullptr, out + ((item * 2 + c) * L + i) * S,
                             C->n_inv_f[i], C->inv_sp_f[i], qd, qi, C->qf[K - 1], static_cast<double>(C->half_sp),
                             static_cast<double>(C->half_sp_mod_q[i]), nullptr};
    ntt_inv_core_f64<LOGS, 0, StoreModDownAdd, whole_threads(LOGS)>(fm, tw.inv_f(i), qd, qi, 0, nt, st);
  }
};

// ------------------------------------------------------------------------------------------------------------
// Inverse transforms as two-CTA clusters (FP64 path). The last Gentleman-Sande stage of an N-point inverse transform
// pairs residue i with i + N/2; everything before it happens inside the two halves independently. CTA h of a cluster
// runs the N/2-point sub-transform of half h in its own 68 KiB of shared memory (two CTAs per SM, as for the forward
// half-limb kernels), the pair meets at a cluster barrier, and each CTA then finishes one quarter of the butterflies of
// the last stage, reading the partner's half through distributed shared memory. The Plan supplies, per limb, the source,
// the NTT table and the store functor that receives the (unscaled) outputs: plain scaling, or the fused ModDown variants.
template <int LOGH, class Plan>
struct InvClusterBody {
  static constexpr const char *kName = Plan::kName;
  static constexpr int kMaxThreads = 512, kMinBlocks = 2;
  Plan plan;
  const DevConsts *C;
  TwRef tw;
  int pf, nl;  // L2 prefetch distance in limbs (0: off), total limbs of the launch
  HD void phase1(int bid, int, unsigned char *smem) const {
    constexpr int S = 1 << LOGH;
    constexpr int nt = half_threads(LOGH);
    const int h = bid & 1, lb = bid >> 1;
    if (pf && !h && lb + pf < nl) plan.prefetch(lb + pf, C, 2 * S);
    const int tab = plan.tab(lb, C);
    const u64 *src = plan.src(lb, C, 2 * S) + static_cast<size_t>(h) * S;
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[tab], qi = C->qinvf[tab];
    FOR_THREADS(tid, nt) {
#pragma unroll 8
      for (int j = tid; j < S; j += nt) fm[pidx(j)] = u_to_f(src[j]);
    }
    SYNC();
    ntt_inv_core_f64<LOGH, 1, SmemIO, nt>(fm, tw.inv_f(tab), qd, qi, h, nt);
  }
  HD void phase2(int bid, int, const unsigned char *smem, const unsigned char *peer) const {
    constexpr int S = 1 << LOGH;
    constexpr int nt = half_threads(LOGH);
    const int h = bid & 1, lb = bid >> 1;
    const int tab = plan.tab(lb, C);
    const double qd = C->qf[tab], qi = C->qinvf[tab];
    const double *lo = reinterpret_cast<const double *>(h ? peer : smem);  // residues [0, S) of the limb
    const double *hi = reinterpret_cast<const double *>(h ? smem : peer);  // residues [S, 2S)
    const double w = tw.inv_f(tab).idx[1];
    const auto st = plan.store(lb, C, 2 * S);
    FOR_THREADS(tid, nt) {
      // one of the two operands comes from the partner CTA over distributed shared memory (long latency): issue the loads of
      // U butterflies before the first use
      // and the global operands of the fused store (special-limb residue, addend) likewise: the stores in between keep the
      // compiler from hoisting those loads on its own
      using Aux = typename decltype(st)::Aux;
      constexpr int U = std::is_empty<Aux>::value ? 4 : 2;
      const int end = (h + 1) * (S / 2);
      for (int j0 = h * (S / 2) + tid; j0 < end; j0 += nt * U) {
        double a[U], b[U];
        Aux x0[U], x1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt < end ? j0 + u * nt : j0;
          a[u] = lo[pidx(j)];  // |.| <= 4.5q
          b[u] = hi[pidx(j)];
          x0[u] = st.aux(j);
          x1[u] = st.aux(j + S);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * nt;
          if (j < end) {
            st.store(j, f_add(a[u], b[u]), x0[u]);
            st.store(j + S, f_mulmod_var(f_add(a[u], -b[u]), w, qd, qi), x1[u]);
          }
        }
      }
    }
  }
};

// Forward counterpart for whole-limb transforms (FP64 path): the pair of half-limb CTAs of a limb forms a two-CTA cluster. Each
// folds the first Cooley-Tukey stage into its load (reading the whole limb), the cluster barrier separates all reads of the limb
// from the first store, so the transform may run in place; the passes and the store then touch only the CTA's own half.
template <int LOGH>
struct NttFwdClusterBody {
  static constexpr const char *kName = "ntt";
  static constexpr int kMaxThreads = 512, kMinBlocks = 2;
  static constexpr bool kPeerSmem = false;  // the cluster barrier only orders the pair's global reads before the first write
  const u64 *in;
  u64 *out;
  const DevConsts *C;
  TwRef tw;
  TabMap map;
  int limbs;
  size_t istride, lstride;
  int pf, nl;  // L2 prefetch distance in limbs (0: off), total limbs of the launch
  HD size_t at(int lb) const { return static_cast<size_t>(lb / limbs) * istride + static_cast<size_t>(lb % limbs) * lstride; }
  HD void phase1(int bid, int, unsigned char *smem) const {
    constexpr int nt = half_threads(LOGH);
    const int h = bid & 1, lb = bid >> 1, tab = map.id[lb % limbs];
    if (pf && !h && lb + pf < nl) cta_prefetch_l2(in + at(lb + pf), sizeof(u64) << (LOGH + 1));
    double *fm = reinterpret_cast<double *>(smem);
    // every global read of the limb happens here (the cluster barrier that follows separates them from the in-place stores)
    fwd_half_transform_f64<LOGH>(fm, tw.fwd_f(tab), C->qf[tab], C->qinvf[tab], h, nt, RawU64{in + at(lb)});
  }
  HD void phase2(int bid, int, unsigned char *smem, const unsigned char *) const {
    constexpr int nt = half_threads(LOGH);
    constexpr int S = 1 << LOGH;
    const int h = bid & 1, lb = bid >> 1, tab = map.id[lb % limbs];
    double *fm = reinterpret_cast<double *>(smem);
    const double qd = C->qf[tab], qi = C->qinvf[tab];
    u64 *dst = out + at(lb) + static_cast<size_t>(h) * S;
    FOR_THREADS(tid, nt) {
#pragma unroll 4
      for (int j = tid; j < S; j += nt) dst[j] = f_canonical(fm[pidx(j)], qd, qi);
    }
  }
};

struct PlanScaled {  // NttBody's inverse branch: out = INTT(in) (scaled by 1/N, canonical)
  static constexpr const char *kName = "ntt";
  const u64 *in;
  u64 *out;
  TabMap map;
  int limbs;
  size_t istride, lstride;
  HD int tab(int lb, const DevConsts *) const { return map.id[lb % limbs]; }
  HD size_t at(int lb) const { return static_cast<size_t>(lb / limbs) * istride + static_cast<size_t>(lb % limbs) * lstride; }
  HD const u64 *src(int lb, const DevConsts *, int) const { return in + at(lb); }
  HD void prefetch(int lb, const DevConsts *, int N) const { cta_prefetch_l2(in + at(lb), sizeof(u64) * N); }
  HD StoreScaled store(int lb, const DevConsts *C, int) const {
    const int t = tab(lb, C);
    return StoreScaled{out + at(lb), C->n_inv_f[t], C->qf[t], C->qinvf[t]};
  }
};
struct PlanModDownGalois {  // InttModDownBody
  static constexpr const char *kName = "intt_moddown";
  const u64 *acc;
  u64 *c1, *g1;
  u32 elt;
  int logn;
  HD int tab(int lb, const DevConsts *C) const { return lb % C->L; }
  HD const u64 *src(int lb, const DevConsts *C, int N) const {
    const size_t item = lb / C->L;
    return acc + ((item * 2 + 1) * C->K + lb % C->L) * N;
  }
  HD void prefetch(int lb, const DevConsts *C, int N) const { cta_prefetch_l2(src(lb, C, N), sizeof(u64) * N); }
  HD StoreModDownGalois store(int lb, const DevConsts *C, int N) const {
    const int L = C->L, K = C->K, i = lb % L;
    const size_t item = lb / L, o = (item * L + i) * N;
    return StoreModDownGalois{acc + ((item * 2 + 1) * K + (K - 1)) * N, c1 ? c1 + o : nullptr, g1 + o, C->n_inv_f[i], C->inv_sp_f[i], C->qf[i], C->qinvf[i],
                              C->qf[K - 1], static_cast<double>(C->half_sp), static_cast<double>(C->half_sp_mod_q[i]), C->mod[i].q, elt,
                              static_cast<u32>(N - 1), logn};
  }
};
struct PlanModDownAdd {  // InttModDownAddBody
  static constexpr const char *kName = "intt_moddown";
  const u64 *acc, *base0, *base1;
  size_t bstride;
  u64 *out;
  const u64 *accum;  // optional [items][2][L][N]: out = accum + key-switch result (+ base); may alias out
  HD int tab(int lb, const DevConsts *C) const { return lb % C->L; }
  HD const u64 *src(int lb, const DevConsts *C, int N) const {
    const int L = C->L;
    const size_t item = lb / (2 * L);
    return acc + ((item * 2 + ((lb / L) & 1)) * C->K + lb % L) * N;
  }
  HD void prefetch(int lb, const DevConsts *C, int N) const {
    cta_prefetch_l2(src(lb, C, N), sizeof(u64) * N);
    const int L = C->L;
    const u64 *bs = ((lb / L) & 1) ? base1 : base0;
    if (bs) cta_prefetch_l2(bs + static_cast<size_t>(lb / (2 * L)) * bstride + static_cast<size_t>(lb % L) * N, sizeof(u64) * N);
  }
  HD StoreModDownAdd store(int lb, const DevConsts *C, int N) const {
    const int L = C->L, K = C->K, i = lb % L, c = (lb / L) & 1;
    const size_t item = lb / (2 * L);
    const u64 *bs = c ? base1 : base0;
    return StoreModDownAdd{acc + ((item * 2 + c) * K + (K - 1)) * N, bs ? bs + item * bstride + static_cast<size_t>(i) * N : nullptr,
                           out + ((item * 2 + c) * L + i) * N, C->n_inv_f[i], C->inv_sp_f[i], C->qf[i], C->qinvf[i], C->qf[K - 1],
                           static_cast<double>(C->half_sp), static_cast<double>(C->half_sp_mod_q[i]),
                           accum ? accum + ((item * 2 + c) * L + i) * N : nullptr};
  }
};

// The two kernels that finish a rotation of the resident chain -- corr_mac (both NTT-resident components + plaintext products) and
// intt_moddown (component 1 back to coefficient form, ModDown, next Galois map) -- read the same key-switch accumulators and do
// not depend on each other: one launch runs both (CTA pairs [0, n_corr) are corr_mac's, the rest intt_moddown's two-CTA clusters),
// so they overlap instead of running back to back and a rotation is three launches (ks_digits, special-limb INTT, this).
template <int LOGH>
struct RotTailBody {
  static constexpr const char *kName = "rot_tail";
  static constexpr int kMaxThreads = 512, kMinBlocks = 2;
  Corr0MacHalfBody<LOGH> corr;
  InvClusterBody<LOGH, PlanModDownGalois> inv;
  int n_corr;  // even
  HD void phase1(int bid, int nt, unsigned char *smem) const {
    if (bid < n_corr)
      corr(bid, nt, smem);
    else
      inv.phase1(bid - n_corr, nt, smem);
  }
  HD void phase2(int bid, int nt, const unsigned char *smem, const unsigned char *peer) const {
    if (bid >= n_corr) inv.phase2(bid - n_corr, nt, smem, peer);
  }
};

// value -> (value, floor(value * 2^64 / q)) for uploaded key-switching keys (one-time, at hhe_load_ksk)
struct ShoupifyBody {
  static constexpr const char *kName = "shoupify";
  const u64 *in;  // [L][2][K][N]
  W2 *out;        // integer limbs: W2 per residue; compact FP64 mode (every key limb on the FP64 path): double per residue
  const DevConsts *C;
  size_t total;
  int compact_f64;
  int group_major;  // FP64 compact keys for KsDigitsTmemBody: residue 8g+e of each half limb goes to e*(N/16)+g
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const int limb = static_cast<int>(static_cast<u32>(g >> C->logn) % static_cast<u32>(C->K));
        const DevMod m = C->mod[limb];
        const u64 w = in[g];
        if (compact_f64) {
          size_t dst = g;
          if (group_major) {
            const size_t half = N >> 1, p = g & (N - 1), r = p & (half - 1);
            dst = (g - p) + (p - r) + (r & 7) * (half >> 3) + (r >> 3);
          }
          reinterpret_cast<double *>(out)[dst] = u_to_f(w);
        } else {
        // floor(w * 2^64 / q): Barrett estimate from floor(2^128/q), then exact correction
        u64 est = mulhi64(w, m.cr0) + w * m.cr1;
        // remainder check: w*2^64 - est*q must be in [0, q)
        u64 rem = 0 - est * m.q;  // low 64 bits of (w<<64) - est*q
        while (rem >= m.q) {
          rem -= m.q;
          ++est;
        }
        out[g] = W2{w, est};
        }
      }
    }
  }
};

}  // namespace hhe
