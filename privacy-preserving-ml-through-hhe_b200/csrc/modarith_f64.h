// Exact modular arithmetic on the FP64 pipe for moduli q <= 2^49 (all BFVDefault primes up to N = 16384).
//
// Why: on B200 the 64-bit integer Shoup product is bound by the FMA-heavy pipe (IMAD.WIDE / IMAD.HI issue every 4
// cycles per sub-partition: ~28 pipe cycles per product, ncu profiles/r1_*), while the FP64 pipe of the same SM runs
// DFMA at 16 lanes/clk per sub-partition (2 cycles per warp instruction) and is otherwise idle. A modular product
// needs 6 FP64 instructions here, a butterfly 8.
//
// Representation: residues are doubles holding (signed) integers; every operand of a product has magnitude <= 4q < 2^51
// and every sum stays below 2^53, so every value, sum and difference below is an exactly representable integer.
// All operations are written with explicit fma / mul / add (no contraction) and are exact:
//   h  = RN(b*w)            l = fma(b, w, -h) = b*w - h exactly (error-free product)
//   qh = rint(b * winv)     winv ~ w/q (relative error <= 2^-52);  |b*w/q - qh| <= 1 for |b| <= 4q < 2^51
//   r  = fma(-qh, q, h)     exact: |h - qh*q| <= q + |l| < 2^51
//   b*w mod q  ==  r + l    an integer of magnitude <= q
// Only the final canonical residue in [0, q) leaves a kernel, so results are bit-identical with integer arithmetic.
#pragma once
#include <cstring>

#include "hd.h"

#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstdlib>
#endif

namespace hhe {

constexpr u64 kF64ModLimit = 1ULL << 49;  // q < 2^49  =>  4q < 2^51 (f_rint_mul), 8q <= 2^52 (key inner product)

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

// round-to-nearest integer of x*y for |x*y| < 2^51: adding 1.5 * 2^52 moves the product into [2^52, 2^53) where the
// spacing of doubles is 1, so the fused multiply-add itself rounds to an integer; subtracting the constant is exact.
// (No sign handling, no extra registers: 2 FP64 instructions.) Every caller keeps |x| <= 4q < 2^51 and 0 <= y <= 1.
HD double f_rint_mul(double x, double y) {
  constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52
#if defined(HHE_EMULATE)
  if (!(std::fabs(x * y) < 2251799813685248.0)) std::abort();  // bound discipline check (test harness only)
#endif
  return f_add(f_fma(x, y, kMagic), -kMagic);
}

// b * w mod q for a precomputed constant (w, winv ~ w/q): result is an integer with |result| <= q, for |b| <= 4q
HD double f_mulmod_const(double b, D2 c, double q) {
  const double qh = f_rint_mul(b, c.winv);
  const double h = f_mul(b, c.w);
  const double l = f_fma(b, c.w, -h);
  return f_add(f_fma(-qh, q, h), l);
}

// a * b mod q for two variable operands (|a| <= 4q, 0 <= b < q): quotient from h * (1/q); |result| <= 2q
HD double f_mulmod_var(double a, double b, double q, double qinv) {
  const double h = f_mul(a, b);
  const double l = f_fma(a, b, -h);
  const double qh = f_rint_mul(h, qinv);
  return f_add(f_fma(-qh, q, h), l);
}

// x mod q into [-q/2 - 1, q/2 + 1] for |x| < 2^53
HD double f_reduce(double x, double q, double qinv) { return f_fma(-f_rint_mul(x, qinv), q, x); }

// any |x| < 2^53 -> canonical residue in [0, q) as an unsigned integer
HD u64 f_canonical(double x, double q, double qinv) {
  double r = f_reduce(x, q, qinv);
  if (r < 0.0) r = f_add(r, q);
  if (r >= q) r = f_add(r, -q);
  return f_to_u(r);
}

}  // namespace hhe
