// 64-bit modular arithmetic primitives (host+device). Moduli are < 2^61.
//   Shoup multiplication by a constant with precomputed quotient (seal/util/uintarithsmallmod.h:255-326)
//   Barrett 128->64 and 64->64 reduction (seal/util/uintarithsmallmod.h:167-230)
// Only the mathematical residue matters for parity: every value that leaves a kernel is canonical in [0, q).
#pragma once
#include "hd.h"

namespace hhe {

struct Twiddle;  // params.h; layout {w, ws}

struct DevMod {
  u64 q;
  u64 cr0, cr1;  // floor(2^128 / q), low and high words
};

struct W2 {  // device view of a Shoup pair
  u64 w, ws;
};

// x * w mod q, result in [0, 2q), any 64-bit x
HD u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q) { return x * w - mulhi64(x, ws) * q; }
HD u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) {
  u64 r = mul_shoup_lazy(x, w, ws, q);
  return r >= q ? r - q : r;
}
HD u64 mul_shoup(u64 x, W2 c, u64 q) { return mul_shoup(x, c.w, c.ws, q); }

// ---- "wide-slack" arithmetic for moduli below 2^57 -------------------------------------------------------------
// floor(x*ws / 2^64) computed from three 32x32 products (the lo*lo partial product and one carry are dropped):
// the estimate is low by at most 2, so the Shoup product below lands in [0, 4q) instead of [0, 2q). With q < 2^57
// a 14-stage transform can let values grow by 4q per stage (< 60q < 2^64) and skip every conditional subtraction.
HD u64 mulhi_approx(u64 x, u64 ws) {
  const u32 xl = static_cast<u32>(x), xh = static_cast<u32>(x >> 32);
  const u32 wl = static_cast<u32>(ws), wh = static_cast<u32>(ws >> 32);
#if defined(__CUDA_ARCH__)
  const u32 c1 = __umulhi(xl, wh), c2 = __umulhi(xh, wl);
#else
  const u32 c1 = static_cast<u32>((static_cast<u64>(xl) * wh) >> 32), c2 = static_cast<u32>((static_cast<u64>(xh) * wl) >> 32);
#endif
  return static_cast<u64>(xh) * wh + c1 + c2;
}
// x * w mod q up to a multiple of q: result in [0, 4q) for any 64-bit x. nq = 2^64 - q.
HD u64 mul_shoup_wide(u64 x, u64 w, u64 ws, u64 nq) { return x * w + mulhi_approx(x, ws) * nq; }
constexpr u64 kWideSlackLimit = 1ULL << 57;

HD u64 add_mod(u64 a, u64 b, u64 q) {
  u64 s = a + b;
  return s >= q ? s - q : s;
}
HD u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
HD u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }
HD u64 csub(u64 a, u64 q) { return a >= q ? a - q : a; }

// x mod q for any 64-bit x
HD u64 barrett64(u64 x, const DevMod &m) {
  u64 r = x - mulhi64(x, m.cr1) * m.q;
  return r >= m.q ? r - m.q : r;
}

// (hi:lo) mod q, requires the result of the estimate to fit, i.e. hi < q (always true for products of residues)
HD u64 barrett128(u64 lo, u64 hi, const DevMod &m) {
  // qhat = floor((hi:lo) * (cr1:cr0) / 2^128), assembled from 64-bit partial products
  u64 c = mulhi64(lo, m.cr0);
  u64 p1l = lo * m.cr1, p1h = mulhi64(lo, m.cr1);
  u64 s1 = p1l + c;
  u64 k1 = p1h + (s1 < p1l);
  u64 p2l = hi * m.cr0, p2h = mulhi64(hi, m.cr0);
  u64 s2 = s1 + p2l;
  u64 k2 = p2h + (s2 < s1);
  u64 qhat = hi * m.cr1 + k1 + k2;
  u64 r = lo - qhat * m.q;
  return r >= m.q ? r - m.q : r;
}

HD u64 mul_mod(u64 a, u64 b, const DevMod &m) { return barrett128(a * b, mulhi64(a, b), m); }
// a*b + c mod q  (c < 2^64, a*b + c must stay below q * 2^64)
HD u64 mul_add_mod(u64 a, u64 b, u64 c, const DevMod &m) {
  u64 lo = a * b, hi = mulhi64(a, b);
  lo += c;
  hi += (lo < c);
  return barrett128(lo, hi, m);
}

// Sum of up to 32 products of residues (each factor < 2^61) kept as a 128-bit integer and reduced ONCE: the inner products of the
// BEHZ base conversions (sum_i z_i * c_i mod m) cost one multiplication per term instead of a multiplication plus a 128-bit Barrett
// reduction per term. 32 * 2^122 < 2^128, so the accumulator cannot overflow.
struct Acc128 {
  u64 lo = 0, hi = 0;
  HD void mac(u64 a, u64 b) {
    const u64 pl = a * b;
    lo += pl;
    hi += mulhi64(a, b) + (lo < pl);
  }
  // (hi:lo) mod q for ANY 128-bit value: hi is reduced first so that barrett128's precondition (hi < q) holds
  HD u64 reduce(const DevMod &m) const { return barrett128(lo, barrett64(hi, m), m); }
};

}  // namespace hhe
