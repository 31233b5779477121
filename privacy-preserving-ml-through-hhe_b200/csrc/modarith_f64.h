// Exact modular arithmetic on the FP64 pipe for moduli q <= 2^49 (all BFVDefault primes up to N = 16384).
//
// Why: on B200 the 64-bit integer Shoup product is bound by the FMA-heavy pipe (IMAD.WIDE / IMAD.HI issue every 4
// cycles per sub-partition: ~28 pipe cycles per product, ncu profiles/r1_*), while the FP64 pipe of the same SM runs
// DFMA/DMUL/DADD at one warp instruction every 2 cycles per sub-partition. Integer multiplies do NOT overlap with FP64
// work (tools/microbench/pipes.cu: DFMA x8 + IMAD.WIDE x4 takes the sum of both), but the FP64 round-to-integer
// conversion FRND.F64 runs on the conversion unit (one warp instruction per ~9 cycles) concurrently with the FP64 pipe
// (DFMA x7 + FRND costs the same as DFMA x7). A modular product is therefore 5 FP64 instructions + 1 FRND, a butterfly
// 7 FP64 + 1 FRND.
//
// Representation: residues are doubles holding (signed) integers of magnitude < 2^53 (16q), so every value, sum and
// difference below is an exactly representable integer. All operations are written with explicit fma / mul / add (no
// contraction) and are exact:
//   h  = RN(b*w)            l = fma(b, w, -h) = b*w - h exactly (error-free product)
//   qh = rint(RN(h * qinv)) qinv = RN(1/q): three roundings, |b*w/q - qh| <= 1/2 + 3 |b| 2^-53
//   r  = fma(-qh, q, h)     exact: |h - qh*q| is a small multiple of q, far below 2^53
//   b*w mod q  ==  r + l    an integer of magnitude <= (1/2 + 3 |b| 2^-53) q  (f64_tbound16 in ntt_core.h tracks this)
// Only the final canonical residue in [0, q) leaves a kernel, so results are bit-identical with integer arithmetic.
#pragma once
#include <cstring>

#include "hd.h"

#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstdlib>
#endif

namespace hhe {

constexpr u64 kF64ModLimit = 1ULL << 49;  // q < 2^49  =>  16q < 2^53: residues up to 12q plus one product stay exact integers

struct D2 {  // FP64 twiddle / key element: value and value/q
  double w, winv;
};

HD double f_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
HD double f_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;  // volatile: keep the host compiler from contracting into a later add
  return r;
#endif
}
HD double f_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
HD double bits_to_double(u64 b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(static_cast<long long>(b));
#else
  double d;
  std::memcpy(&d, &b, 8);
  return d;
#endif
}
HD u64 double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
  return static_cast<u64>(__double_as_longlong(d));
#else
  u64 b;
  std::memcpy(&b, &d, 8);
  return b;
#endif
}

constexpr u64 kTwo52Bits = 0x4330000000000000ULL;  // 2^52 as a bit pattern

// unsigned integer < 2^52 -> double (exact): splice into the mantissa of 2^52, subtract 2^52
HD double u_to_f(u64 x) { return f_add(bits_to_double(x | kTwo52Bits), -4503599627370496.0); }
// double holding an integer in [0, 2^52) -> unsigned
HD u64 f_to_u(double d) { return double_to_bits(f_add(d, 4503599627370496.0)) & ((1ULL << 52) - 1); }

// round-to-nearest-even integer: FRND.F64 on the device (conversion unit, overlaps with the FP64 pipe)
HD double f_rint(double x) {
#if defined(__CUDA_ARCH__)
  return rint(x);
#else
  return std::nearbyint(x);
#endif
}

HD void f_check_operand(double b, double q) {
#if defined(HHE_EMULATE)
  if (!(std::fabs(b) <= 12.0 * q)) std::abort();  // bound discipline check (test harness only)
#else
  (void)b;
  (void)q;
#endif
}

// b * w mod q for a precomputed constant (w, winv ~ w/q): |result| <= (1/2 + 2 |b| 2^-53 + 2^-3) q, for |b| <= 12q
HD double f_mulmod_const(double b, D2 c, double q) {
  f_check_operand(b, q);
  const double qh = f_rint(f_mul(b, c.winv));
  const double h = f_mul(b, c.w);
  const double l = f_fma(b, c.w, -h);
  return f_add(f_fma(-qh, q, h), l);
}

// a * b mod q for two variable operands (|a| <= 12q, |b| <= q): quotient from h * (1/q);
// |result| <= (1/2 + 3 |a| 2^-53) q: below 0.7q for |a| <= q, 1.25q for |a| <= 4q, 2q for |a| <= 8q
HD double f_mulmod_var(double a, double b, double q, double qinv) {
  f_check_operand(a, q);
  const double h = f_mul(a, b);
  const double l = f_fma(a, b, -h);
  const double qh = f_rint(f_mul(h, qinv));
  return f_add(f_fma(-qh, q, h), l);
}

// x mod q into [-q/2 - 1, q/2 + 1] for |x| < 2^53
HD double f_reduce(double x, double q, double qinv) { return f_fma(-f_rint(f_mul(x, qinv)), q, x); }

// any |x| < 2^53 -> canonical residue in [0, q) as an unsigned integer
HD u64 f_canonical(double x, double q, double qinv) {
  double r = f_reduce(x, q, qinv);
  if (r < 0.0) r = f_add(r, q);
  if (r >= q) r = f_add(r, -q);
  return f_to_u(r);
}

}  // namespace hhe
