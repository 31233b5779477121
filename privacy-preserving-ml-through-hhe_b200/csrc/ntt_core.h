// Shared-memory negacyclic NTT core (host+device). One CTA owns one (sub-)transform of S = 2^LOGS residues staged
// in shared memory; each thread runs radix-8 (three radix-2 stages) butterflies in registers between barriers, so a
// 2^14 transform needs 5 passes over shared memory instead of 14. Radix-8 keeps a thread at <= 64 registers, which
// lets a 1024-thread CTA (32 warps per SM) hide the IMAD-pipe and L2 latencies; measured on B200 it beats the
// radix-16 / 512-thread / 128-register variant by 17% (HHE_RADIX_LOG=4 HHE_MAX_THREADS=512 rebuilds that one).
//
// Semantics follow SEAL's transform (seal/util/dwthandler.h:94-356, seal/util/ntt.h): forward = Cooley-Tukey,
// natural-order input, bit-reversed output, stage with m groups uses twiddles fwd[m + i] = psi^bitrev(m+i);
// inverse = Gentleman-Sande with inv[m + i] on the same (m, i) pairing. Harvey lazy butterflies: forward values
// stay in [0, 4q), inverse values in [0, 2q); callers canonicalise on the way out.
//
// A transform of size S may be a *sub-transform* of a larger N-point one: after the first log2(N/S) Cooley-Tukey
// stages the N/S contiguous chunks are independent; `mc` = (N/S) + chunk selects the twiddle rows of that chunk.
#pragma once
#include "modarith.h"
#include "modarith_f64.h"

namespace hhe {

// Shared-memory index padding: one extra word per 16 keeps the stride-16/stride-1 mixes of the register passes
// spread over the banks.
HD int pidx(int i) { return i + (i >> 4); }
// A pass whose element stride is 8 (lg == 3) would put lanes 0-7 and 8-15 of a half warp on overlapping banks (their
// 64-residue blocks are 4 padded words apart): exchanging bits 3 and 4 of the group index inside each warp pairs blocks
// that are 8 words apart instead. A bijection on the groups of a warp, so nothing else changes (tools: /tmp-free check
// in DESIGN.md; ncu: 22% of the shared-memory wavefronts were conflict replays before).
HD int stride8_group(int g) { return (g & ~0x18) | ((g & 8) << 1) | ((g & 16) >> 1); }
HD constexpr size_t ntt_smem_words(int S) { return static_cast<size_t>(S) + (S >> 4); }

HD void fwd_bfly(u64 &a, u64 &b, W2 w, u64 q, u64 two_q) {
  u64 x = a >= two_q ? a - two_q : a;
  u64 t = mul_shoup_lazy(b, w.w, w.ws, q);
  a = x + t;
  b = x + two_q - t;
}
HD void inv_bfly(u64 &a, u64 &b, W2 w, u64 q, u64 two_q) {
  u64 s = a + b;
  u64 d = a + two_q - b;
  a = s >= two_q ? s - two_q : s;
  b = mul_shoup_lazy(d, w.w, w.ws, q);
}

// wide-slack forward butterfly (q < 2^57): no conditional subtraction; each output exceeds its input bound by < 4q
HD void fwd_bfly_wide(u64 &a, u64 &b, W2 w, u64 nq, u64 four_q) {
  const u64 t = mul_shoup_wide(b, w.w, w.ws, nq);
  b = a + four_q - t;
  a = a + t;
}

// One register pass over stages [s0, s0+R) for group g (2^R residues).
template <int R, bool WIDE = false>
HD void fwd_group(u64 *sm, const W2 *__restrict__ tw, u64 q, int logS, int s0, u32 mc, int g) {
  constexpr int E = 1 << R;
  const int lg = logS - s0 - R;
  if (lg == 3) g = stride8_group(g);
  const int lo = g & ((1 << lg) - 1), hi = g >> lg;
  const int base = (hi << (logS - s0)) + lo;
  const u64 two_q = q << 1, nq = 0 - q, four_q = q << 2;
  // all 2^R - 1 twiddles of the group are requested first (one L2 round trip for the whole group), then the residues
  W2 wv[E];
#pragma unroll
  for (int d = 0; d < R; ++d) {
    const u32 tb = (mc << (s0 + d)) + (static_cast<u32>(hi) << d);
#pragma unroll
    for (int j = 0; j < (1 << d); ++j) wv[(1 << d) + j] = tw[tb + j];
  }
  u64 x[E];
#pragma unroll
  for (int e = 0; e < E; ++e) x[e] = sm[pidx(base + (e << lg))];
#pragma unroll
  for (int d = 0; d < R; ++d) {
    const int half = E >> (d + 1);
#pragma unroll
    for (int j = 0; j < (1 << d); ++j) {
      const W2 w = wv[(1 << d) + j];
#pragma unroll
      for (int k = 0; k < half; ++k) {
        if (WIDE)
          fwd_bfly_wide(x[2 * j * half + k], x[2 * j * half + k + half], w, nq, four_q);
        else
          fwd_bfly(x[2 * j * half + k], x[2 * j * half + k + half], w, q, two_q);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) sm[pidx(base + (e << lg))] = x[e];
}

template <int R>
HD void inv_group(u64 *sm, const W2 *__restrict__ tw, u64 q, int logS, int s0, u32 mc, int g) {
  constexpr int E = 1 << R;
  const int lg = logS - s0 - R;
  if (lg == 3) g = stride8_group(g);
  const int lo = g & ((1 << lg) - 1), hi = g >> lg;
  const int base = (hi << (logS - s0)) + lo;
  const u64 two_q = q << 1;
  u64 x[E];
#pragma unroll
  for (int e = 0; e < E; ++e) x[e] = sm[pidx(base + (e << lg))];
#pragma unroll
  for (int d = R - 1; d >= 0; --d) {
    const int half = E >> (d + 1);
    const u32 tb = (mc << (s0 + d)) + (static_cast<u32>(hi) << d);
#pragma unroll
    for (int j = 0; j < (1 << d); ++j) {
      const W2 w = tw[tb + j];
#pragma unroll
      for (int k = 0; k < half; ++k) inv_bfly(x[2 * j * half + k], x[2 * j * half + k + half], w, q, two_q);
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) sm[pidx(base + (e << lg))] = x[e];
}

#ifndef HHE_RADIX_LOG
#define HHE_RADIX_LOG 3
#endif
constexpr int kRadixLog = HHE_RADIX_LOG;  // stages per register pass (4: radix-16, 3: radix-8)
HD constexpr int NttSchedule_first(int logs) { return logs - kRadixLog * ((logs - 1) / kRadixLog); }
template <int LOGS>
struct NttSchedule {
  static constexpr int kFirst = LOGS - kRadixLog * ((LOGS - 1) / kRadixLog);  // stages in the odd-sized pass
};

// Forward transform of the S residues in sm (padded layout). Ends with a barrier.
//   q >= 2^57 (BEHZ auxiliary primes): Harvey lazy butterflies, in [0, 4q) -> out [0, 4q).
//   q <  2^57: wide-slack butterflies, in [0, B) -> out [0, B + 4q*LOGS); callers reduce with barrett64 or feed the
//              result to a Shoup/Barrett multiplication that accepts any 64-bit operand.
template <int LOGS>
HD void ntt_fwd_core(u64 *sm, const W2 *__restrict__ tw, u64 q, u32 mc, int nt) {
  constexpr int R0 = NttSchedule<LOGS>::kFirst;
  if (q < kWideSlackLimit) {
    FOR_THREADS(tid, nt) {
      for (int g = tid; g < (1 << (LOGS - R0)); g += nt) fwd_group<R0, true>(sm, tw, q, LOGS, 0, mc, g);
    }
    SYNC();
    for (int s0 = R0; s0 < LOGS; s0 += kRadixLog) {
      FOR_THREADS(tid, nt) {
        for (int g = tid; g < (1 << (LOGS - kRadixLog)); g += nt) fwd_group<kRadixLog, true>(sm, tw, q, LOGS, s0, mc, g);
      }
      SYNC();
    }
    return;
  }
  FOR_THREADS(tid, nt) {
    for (int g = tid; g < (1 << (LOGS - R0)); g += nt) fwd_group<R0>(sm, tw, q, LOGS, 0, mc, g);
  }
  SYNC();
  for (int s0 = R0; s0 < LOGS; s0 += kRadixLog) {
    FOR_THREADS(tid, nt) {
      for (int g = tid; g < (1 << (LOGS - kRadixLog)); g += nt) fwd_group<kRadixLog>(sm, tw, q, LOGS, s0, mc, g);
    }
    SYNC();
  }
}

// Inverse transform (without the 1/N scaling). In: [0, 2q). Out: [0, 2q). Ends with a barrier.
template <int LOGS>
HD void ntt_inv_core(u64 *sm, const W2 *__restrict__ tw, u64 q, u32 mc, int nt) {
  constexpr int R0 = NttSchedule<LOGS>::kFirst;
  for (int s0 = LOGS - kRadixLog; s0 >= R0; s0 -= kRadixLog) {
    FOR_THREADS(tid, nt) {
      for (int g = tid; g < (1 << (LOGS - kRadixLog)); g += nt) inv_group<kRadixLog>(sm, tw, q, LOGS, s0, mc, g);
    }
    SYNC();
  }
  FOR_THREADS(tid, nt) {
    for (int g = tid; g < (1 << (LOGS - R0)); g += nt) inv_group<R0>(sm, tw, q, LOGS, 0, mc, g);
  }
  SYNC();
}

// ---------------------------------------------------------------------------------------------------------------
// FP64-pipe variants (q <= 2^49, see modarith_f64.h). Shared memory holds doubles (signed integers, |x| < 16q <= 2^53).
//
// Twiddles are plain doubles w (8 bytes); the quotient of a butterfly product is estimated from the product itself
// (f_mulmod_var: rint(RN(b*w) * (1/q))), so no w/q table is needed.
//
// Bound discipline (compile-time, in units of q/16): a product with an operand of magnitude beta*q has magnitude
// <= (1/2 + 3 beta 2^-53 2^49) q = (1/2 + 0.1875 beta) q (f64_tbound16), a forward stage adds that to every residue, an
// inverse stage doubles the sums. Every intermediate must stay below 16q (exact integers in a double); the chains
// below keep it under 12q. A forward register pass runs in one of two modes:
//   kNone  no reduction at all
//   kHalf  the four "a" inputs of the pass's last stage (the residues that are not multiplied there) are reduced to
//          |x| <= q/2 + 1 first, so the pass's outputs are bounded by q/2 + one product whatever came in.
// Inverse passes reduce every residue on load (three doublings follow).
//
// Two table layouts per (modulus, direction), both of N doubles (F64Tw):
//   idx[k]          index-major, SEAL's order (psi^bitrev(k)): used by passes whose twiddles are warp-uniform
//   cm[off(g0) + c*2^g0 + H]  component-major for a 3-stage register pass starting at global stage g0: component
//                   c = 2^d - 1 + j is twiddle j of stage g0+d, H = index of the radix-8 group's 2^(logN-g0) block.
//                   Consecutive lanes read consecutive doubles (2 L1 wavefronts per component instead of up to 16).
struct F64Tw {
  const double *idx;
  const double *cm;
  int gmin;  // first global stage that has a component-major table (= logN mod 3, or 3 if that is 0)
};
HD constexpr size_t f64tw_offset_c(int g0, int gmin) {
  size_t off = 0;
  for (int g = gmin; g < g0; g += 3) off += static_cast<size_t>(7) << g;
  return off;
}
HD size_t f64tw_offset(int g0, int gmin) {
  size_t off = 0;
  for (int g = gmin; g < g0; g += 3) off += static_cast<size_t>(7) << g;
  return off;
}

// bound (units of q/16) of a product whose variable operand is bounded by b16, with one unit of slack
HD constexpr int f64_tbound16(int b16) { return 8 + (3 * b16 + 15) / 16 + 1; }
constexpr int kF64Reduced16 = 9;   // q/2 + 1
constexpr int kF64PassLimit16 = 120;  // a pass may leave at most 7.5q to the next one (a kHalf pass then peaks below 12q)
constexpr int kF64AnyOut16 = 160;     // what f_canonical / f_reduce callers accept from a transform
enum F64Mode { kNone = 0, kHalf = 1, kFull = 2 };
// output bound of a forward pass of R stages entered with bound b16
HD constexpr int f64_fwd_out16(int b16, int R, bool half) {
  for (int s = 0; s < R; ++s) b16 = ((half && s == R - 1) ? kF64Reduced16 : b16) + f64_tbound16(b16);
  return b16;
}
// largest intermediate of a kHalf pass: the residues entering its last stage (operands of that stage's products)
HD constexpr int f64_fwd_peak16(int b16, int R) {
  for (int s = 0; s + 1 < R; ++s) b16 += f64_tbound16(b16);
  return b16;
}

// Register pass over local stages [S0, S0+R) of chunk `chunk` of a transform that was split into 2^LM chunks.
// Every stride is a compile-time constant, so shared-memory and twiddle accesses use immediate offsets.
// IO: optional global-memory side of the pass. The forward transform's first pass (S0 == 0) can take its inputs from
// IO::load(i) and the inverse transform's last pass (S0 == 0) can hand its outputs to IO::store(i, v) instead of going
// through shared memory: in both, consecutive threads touch consecutive residues (coalesced), and one shared-memory
// round trip plus one barrier per transform disappear.
// A third hook, IO::group_out(g, x), receives the 8 outputs of every group of the forward transform's LAST pass in
// registers (the key-switch kernel multiplies them by the key and accumulates without another shared-memory round trip).
struct SmemIO {
  static constexpr bool kLoad = false, kStore = false, kGroupOut = false;
  HD double load(int) const { return 0.0; }
  HD void store(int, double) const {}
  HD void group_out(int, const double *) const {}
};

// Twiddles of group g of the pass (R stages from local stage S0, transform split into 2^LM chunks): wv[(1 << d) + j] is
// twiddle j of the pass's stage d. Separate from the butterflies so that a caller can issue these global loads one group
// ahead of their use (FwdChainF64 / InvChainF64 device paths).
template <int R, int LOGS, int S0, int LM>
HD void group_tw_f64(F64Tw tw, int chunk, int g, double *wv) {
  constexpr int E = 1 << R;
  constexpr int LG = LOGS - S0 - R;
  constexpr int G0 = S0 + LM;
  if (LG == 3) g = stride8_group(g);
  const int hi = g >> LG;
  const int H = (chunk << S0) + hi;  // global block index at stage G0
  if (R == 3 && G0 >= 1 && (G0 % 3) == ((LOGS + LM) % 3)) {
    constexpr int kGmin = (LOGS + LM) % 3 ? (LOGS + LM) % 3 : 3;  // == tw.gmin (Engine: logn % 3, or 3), known at compile time here
    constexpr size_t kOff = f64tw_offset_c(G0, kGmin);
    const double *T = tw.cm + kOff + H;
#pragma unroll
    for (int c = 0; c < E - 1; ++c) wv[c + 1] = T[static_cast<size_t>(c) << G0];
  } else {
#pragma unroll
    for (int d = 0; d < R; ++d) {
#pragma unroll
      for (int j = 0; j < (1 << d); ++j) wv[(1 << d) + j] = tw.idx[(static_cast<size_t>(1) << (G0 + d)) + (static_cast<size_t>(H) << d) + j];
    }
  }
}

// The R butterfly stages of one radix-2^R group held in registers (x[0 .. 2^R)), twiddles in wv (see group_tw_f64).
template <int R, bool INVERSE, int MODE>
HD void group_math_f64(double *x, const double *wv, double q, double qinv) {
  constexpr int E = 1 << R;
  if (MODE == kFull) {
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = f_reduce(x[e], q, qinv);
  }
  if (!INVERSE) {
#pragma unroll
    for (int d = 0; d < R; ++d) {
      const int half = E >> (d + 1);
      if (MODE == kHalf && d == R - 1) {
#pragma unroll
        for (int e = 0; e < E; e += 2) x[e] = f_reduce(x[e], q, qinv);
      }
#pragma unroll
      for (int j = 0; j < (1 << d); ++j) {
        const double w = wv[(1 << d) + j];
#pragma unroll
        for (int k = 0; k < half; ++k) {
          double &a = x[2 * j * half + k], &b = x[2 * j * half + k + half];
          const double t = f_mulmod_var(b, w, q, qinv);
          b = f_add(a, -t);
          a = f_add(a, t);
        }
      }
    }
  } else {
#pragma unroll
    for (int d = R - 1; d >= 0; --d) {
      const int half = E >> (d + 1);
#pragma unroll
      for (int j = 0; j < (1 << d); ++j) {
        const double w = wv[(1 << d) + j];
#pragma unroll
        for (int k = 0; k < half; ++k) {
          double &a = x[2 * j * half + k], &b = x[2 * j * half + k + half];
          const double dlt = f_add(a, -b);
          a = f_add(a, b);
          b = f_mulmod_var(dlt, w, q, qinv);
        }
      }
    }
  }
}

// The butterflies of group g with its twiddles in wv: residues from shared memory (or IO::load), results to shared memory
// (or IO::store / IO::group_out).
template <int R, bool INVERSE, int LOGS, int S0, int LM, int MODE, class IO = SmemIO>
HD void group_core_f64(double *sm, const double *wv, double q, double qinv, int g, const IO &io = IO()) {
  constexpr int E = 1 << R;
  constexpr int LG = LOGS - S0 - R;  // log2 of the element stride inside the group
  if (LG == 3) g = stride8_group(g);
  const int lo = g & ((1 << LG) - 1), hi = g >> LG;
  // padded shared-memory offsets: pidx(base + (e << LG)) = a0 + off(e) with compile-time off(e)
  const int a0 = pidx(hi << (LOGS - S0)) + lo + (LG >= 4 ? (lo >> 4) : 0);
  auto off = [](int e) constexpr { return LG >= 4 ? e * ((1 << LG) + (1 << (LG >= 4 ? LG - 4 : 0))) : (e << LG) + (e >> (LG < 4 ? 4 - LG : 0)); };
  constexpr bool kGlobalIn = IO::kLoad && !INVERSE && S0 == 0;
  constexpr bool kGlobalOut = IO::kStore && INVERSE && S0 == 0;
  constexpr bool kGroupOut = IO::kGroupOut && !INVERSE && S0 + R == LOGS;
  double x[E];
#pragma unroll
  for (int e = 0; e < E; ++e) x[e] = kGlobalIn ? io.load((hi << (LOGS - S0)) + lo + (e << LG)) : sm[a0 + off(e)];
  group_math_f64<R, INVERSE, MODE>(x, wv, q, qinv);
  if constexpr (kGroupOut) {
    io.group_out(g, x);
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (kGlobalOut)
        io.store((hi << (LOGS - S0)) + lo + (e << LG), x[e]);
      else
        sm[a0 + off(e)] = x[e];
    }
  }
}

template <int R, bool INVERSE, int LOGS, int S0, int LM, int MODE, class IO = SmemIO>
HD void group_f64(double *sm, F64Tw tw, double q, double qinv, int chunk, int g, const IO &io = IO()) {
  double wv[1 << R];
  group_tw_f64<R, LOGS, S0, LM>(tw, chunk, g, wv);
  group_core_f64<R, INVERSE, LOGS, S0, LM, MODE, IO>(sm, wv, q, qinv, g, io);
}

// Synchronisation between two consecutive register passes. The groups g in [hi * 2^LGDOM, (hi + 1) * 2^LGDOM) of the
// producing pass write exactly the residues that the same range of groups of the consuming pass reads (LGDOM = the larger
// element-stride exponent of the two passes), and group g belongs to thread g mod NT. When the CTA size NT is known at
// compile time only those 2^LGDOM threads have to meet: a warp-level sync for <= 32 threads, a named barrier for a
// few warps, the CTA barrier otherwise.
template <int LGDOM, int NT>
HD void sync_domain() {
#if defined(__CUDA_ARCH__)
#if defined(HHE_FULL_BARRIERS)
  if constexpr (true) {
#else
  if constexpr (NT == 0 || (1 << LGDOM) >= NT || (NT >> LGDOM) > 15) {
#endif
    __syncthreads();
  } else if constexpr (LGDOM <= 5) {
    __syncwarp();
  } else {
    const unsigned id = 1u + (threadIdx.x >> LGDOM);
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(1u << LGDOM) : "memory");
  }
#endif
}

// Compile-time chain of forward passes. B16 = bound (units of q/16) of the values entering the pass; MAXOUT16 = the
// largest bound the consumer of the transform accepts. A pass runs unreduced (kNone) when what it leaves is within
// kF64PassLimit16 (MAXOUT16 for the last pass), otherwise in kHalf mode.
// NT > 0: the CTA size is known at compile time (the group loops unroll, index arithmetic folds).
template <int LOGS, int LM, int S0, int B16, int MAXOUT16, int NT = 0>
struct FwdChainF64 {
  static constexpr int R = S0 == 0 ? NttSchedule<LOGS>::kFirst : kRadixLog;
  static constexpr bool kLast = S0 + R >= LOGS;
  static constexpr int kLimit = kLast ? MAXOUT16 : kF64PassLimit16;
  static constexpr bool kHalfMode = f64_fwd_out16(B16, R, false) > kLimit;
  static constexpr int kOut = f64_fwd_out16(B16, R, kHalfMode);
  static_assert(f64_fwd_peak16(B16, R) <= 192, "FP64 transform: intermediate bound above 12q");
  static_assert(kOut <= kLimit, "FP64 transform: pass output above the consumer's bound");
  using Next = FwdChainF64<LOGS, LM, (kLast ? 0 : S0 + R), (kLast ? 16 : kOut), MAXOUT16, NT>;
  static constexpr int kGroups = 1 << (LOGS - R);
#if defined(__CUDA_ARCH__)
  // Device path for a compile-time CTA size: the twiddles of a group are requested one group ahead of their use (the second
  // group's while the first one computes, the next pass's first group's before the barrier that ends this pass), so their
  // L1/L2 latency overlaps FP64 work instead of stalling the first product of every group.
  static __device__ __forceinline__ void tw_first(F64Tw tw, int chunk, double *wv) {
    group_tw_f64<R, LOGS, S0, LM>(tw, chunk, static_cast<int>(threadIdx.x), wv);
  }
  template <class IO>
  static __device__ __forceinline__ void run_dev(double *sm, F64Tw tw, double q, double qinv, int chunk, const IO &io, double *w0) {
    constexpr int GPT = kGroups / NT;
    static_assert(GPT >= 1 && GPT * NT == kGroups, "pipelined passes need a whole number of groups per thread");
    const int tid = static_cast<int>(threadIdx.x);
    double w1[8];
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
      double *cur = (j & 1) ? w1 : w0, *nxt = (j & 1) ? w0 : w1;
      if (j + 1 < GPT)
        group_tw_f64<R, LOGS, S0, LM>(tw, chunk, tid + (j + 1) * NT, nxt);
      else {
        if constexpr (!kLast) Next::tw_first(tw, chunk, nxt);
      }
      group_core_f64<R, false, LOGS, S0, LM, kHalfMode ? kHalf : kNone, IO>(sm, cur, q, qinv, tid + j * NT, io);
    }
    if (kLast)
      SYNC();
    else
      sync_domain<LOGS - S0 - R, NT>();
    if constexpr (!kLast) {
      // the next pass's first twiddles sit in the array the last group did not use
      if constexpr (GPT & 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) w0[c] = w1[c];
      }
      Next::run_dev(sm, tw, q, qinv, chunk, io, w0);
    }
  }
#endif
  template <class IO = SmemIO>
  static HD void run(double *sm, F64Tw tw, double q, double qinv, int chunk, int nt, const IO &io = IO()) {
#if defined(__CUDA_ARCH__) && !defined(HHE_NO_TW_PIPELINE)
    if constexpr (NT > 0 && kGroups % (NT > 0 ? NT : 1) == 0 && kGroups >= NT) {
      double w0[8];
      tw_first(tw, chunk, w0);
      run_dev(sm, tw, q, qinv, chunk, io, w0);
      return;
    }
#endif
    const int st = NT ? NT : nt;
    FOR_THREADS(tid, nt) {
#pragma unroll
      for (int g = tid; g < (1 << (LOGS - R)); g += st)
        group_f64<R, false, LOGS, S0, LM, kHalfMode ? kHalf : kNone, IO>(sm, tw, q, qinv, chunk, g, io);
    }
    if (kLast)
      SYNC();
    else
      sync_domain<LOGS - S0 - R, NT>();  // this pass's element stride is the larger one
    if (!kLast) Next::run(sm, tw, q, qinv, chunk, nt, io);
  }
};

// Forward transform on doubles. B2IN = twice the input bound in units of q (2 for canonical residues).
// Output bound <= MAXOUT16 / 16 q.
template <int LOGS, int LM, int B2IN, class IO = SmemIO, int MAXOUT16 = kF64AnyOut16, int NT = 0>
HD void ntt_fwd_core_f64(double *sm, F64Tw tw, double q, double qinv, int chunk, int nt, const IO &io = IO()) {
  static_assert(kRadixLog == 3, "FP64 path is written for radix-8 register passes");
  FwdChainF64<LOGS, LM, 0, B2IN * 8, MAXOUT16, NT>::run(sm, tw, q, qinv, chunk, nt, io);
}

// Same, entering the chain at local stage S0 (the caller already performed the stages below S0).
template <int LOGS, int LM, int S0, int B2IN, class IO = SmemIO, int MAXOUT16 = kF64AnyOut16, int NT = 0>
HD void ntt_fwd_core_f64_from(double *sm, F64Tw tw, double q, double qinv, int chunk, int nt, const IO &io = IO()) {
  FwdChainF64<LOGS, LM, S0, B2IN * 8, MAXOUT16, NT>::run(sm, tw, q, qinv, chunk, nt, io);
}

// Inverse passes, highest stages first; every pass reduces on load (q/2 + 1, then three doublings: <= 4.5q).
template <int LOGS, int LM, int S0, int NT = 0>
struct InvChainF64 {
  static constexpr int R0 = NttSchedule<LOGS>::kFirst;
  static constexpr int R = S0 == 0 ? R0 : kRadixLog;
  static constexpr int kNext = S0 - kRadixLog >= R0 ? S0 - kRadixLog : 0;  // first stage of the pass that consumes this one
  static constexpr int kGroups = 1 << (LOGS - R);
  using Next = InvChainF64<LOGS, LM, kNext, NT>;
#if defined(__CUDA_ARCH__)
  // device path for a compile-time CTA size: twiddles requested one group ahead (see FwdChainF64::run_dev)
  static __device__ __forceinline__ void tw_first(F64Tw tw, int chunk, double *wv) {
    group_tw_f64<R, LOGS, S0, LM>(tw, chunk, static_cast<int>(threadIdx.x), wv);
  }
  template <class IO>
  static __device__ __forceinline__ void run_dev(double *sm, F64Tw tw, double q, double qinv, int chunk, const IO &io, double *w0) {
    constexpr int GPT = kGroups / NT;
    static_assert(GPT >= 1 && GPT * NT == kGroups, "pipelined passes need a whole number of groups per thread");
    const int tid = static_cast<int>(threadIdx.x);
    double w1[8];
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
      double *cur = (j & 1) ? w1 : w0, *nxt = (j & 1) ? w0 : w1;
      if (j + 1 < GPT)
        group_tw_f64<R, LOGS, S0, LM>(tw, chunk, tid + (j + 1) * NT, nxt);
      else {
        if constexpr (S0 > 0) Next::tw_first(tw, chunk, nxt);
      }
      group_core_f64<R, true, LOGS, S0, LM, kFull, IO>(sm, cur, q, qinv, tid + j * NT, io);
    }
    if (S0 == 0)
      SYNC();
    else
      sync_domain<LOGS - kNext - (kNext == 0 ? R0 : kRadixLog), NT>();
    if constexpr (S0 > 0) {
      if constexpr (GPT & 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) w0[c] = w1[c];
      }
      Next::run_dev(sm, tw, q, qinv, chunk, io, w0);
    }
  }
#endif
  template <class IO = SmemIO>
  static HD void run(double *sm, F64Tw tw, double q, double qinv, int chunk, int nt, const IO &io = IO()) {
#if defined(__CUDA_ARCH__) && !defined(HHE_NO_TW_PIPELINE)
    if constexpr (NT > 0 && kGroups % (NT > 0 ? NT : 1) == 0 && kGroups >= NT) {
      double w0[8];
      tw_first(tw, chunk, w0);
      run_dev(sm, tw, q, qinv, chunk, io, w0);
      return;
    }
#endif
    const int st = NT ? NT : nt;
    FOR_THREADS(tid, nt) {
#pragma unroll
      for (int g = tid; g < (1 << (LOGS - R)); g += st) group_f64<R, true, LOGS, S0, LM, kFull, IO>(sm, tw, q, qinv, chunk, g, io);
    }
    if (S0 == 0)
      SYNC();
    else
      sync_domain<LOGS - kNext - (kNext == 0 ? R0 : kRadixLog), NT>();  // the consumer's element stride is the larger one
    if (S0 > 0) Next::run(sm, tw, q, qinv, chunk, nt, io);
  }
};

// Inverse transform on doubles (without 1/N unless the IO functor applies it). With a storing IO functor the results
// leave through IO::store and shared memory holds garbage afterwards.
template <int LOGS, int LM, class IO = SmemIO, int NT = 0>
HD void ntt_inv_core_f64(double *sm, F64Tw tw, double q, double qinv, int chunk, int nt, const IO &io = IO()) {
  constexpr int R0 = NttSchedule<LOGS>::kFirst;
  InvChainF64<LOGS, LM, (LOGS - kRadixLog >= R0 ? LOGS - kRadixLog : 0), NT>::run(sm, tw, q, qinv, chunk, nt, io);
}

}  // namespace hhe
