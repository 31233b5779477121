// BLAKE2b / BLAKE2xb (RFC 7693; the BLAKE2X extendable-output construction of the BLAKE2 reference code) as host+device functions.
//
// Why it is here: SEAL 4.0's default random generator (seal/randomgen.h: Blake2xbPRNG) produces its byte stream as
//     refill k = blake2xb(out = 4096 bytes, in = counter k as 8 little-endian bytes, key = the 64-byte seed),  k = 0, 1, 2, ...
// and Encryptor::encrypt draws the ternary polynomial u and the two noise polynomials from that stream (seal/util/rlwe.h). To
// reproduce seal::Encryptor::encrypt on the GPU bit for bit the engine regenerates the same stream:
// every 64-byte output block of every refill is an independent BLAKE2b instance (that is what makes BLAKE2X parallel), one CUDA
// thread per block. Written from the RFC; no code of the reference tree is used.
#pragma once
#include "hd.h"

namespace hhe {

HD u64 blake2_iv(int i) {
  switch (i) {
    case 0: return 0x6a09e667f3bcc908ULL;
    case 1: return 0xbb67ae8584caa73bULL;
    case 2: return 0x3c6ef372fe94f82bULL;
    case 3: return 0xa54ff53a5f1d36f1ULL;
    case 4: return 0x510e527fade682d1ULL;
    case 5: return 0x9b05688c2b3e6c1fULL;
    case 6: return 0x1f83d9abfb41bd6bULL;
    default: return 0x5be0cd19137e2179ULL;
  }
}

HD u64 blake2_rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }

// message word schedule of round r (rounds 10 and 11 repeat 0 and 1), packed 4 bits per entry
HD int blake2_sigma(int r, int i) {
  constexpr u64 kS[10] = {0xfedcba9876543210ULL, 0x357b20c16df984aeULL, 0x491763eadf250c8bULL, 0x8f04a562ebcd1397ULL,
                          0xd386cb1efa427509ULL, 0x91ef57d438b0a6c2ULL, 0xb8293670a4def15cULL, 0xa2684f05931ce7bdULL,
                          0x5a417d2c803b9ef6ULL, 0x0dc3e9bf5167482aULL};
  return static_cast<int>((kS[r % 10] >> (4 * i)) & 15);
}

// one compression: h <- F(h, m, t, last)   (byte counter t < 2^64 is all this use needs)
HD void blake2b_compress(u64 *h, const u64 *m, u64 t, bool last) {
  u64 v[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = h[i];
    v[8 + i] = blake2_iv(i);
  }
  v[12] ^= t;
  if (last) v[14] = ~v[14];
#define HHE_B2G(a, b, c, d, x, y)          \
  v[a] = v[a] + v[b] + (x);                \
  v[d] = blake2_rotr(v[d] ^ v[a], 32);     \
  v[c] = v[c] + v[d];                      \
  v[b] = blake2_rotr(v[b] ^ v[c], 24);     \
  v[a] = v[a] + v[b] + (y);                \
  v[d] = blake2_rotr(v[d] ^ v[a], 16);     \
  v[c] = v[c] + v[d];                      \
  v[b] = blake2_rotr(v[b] ^ v[c], 63);
  for (int r = 0; r < 12; ++r) {
    HHE_B2G(0, 4, 8, 12, m[blake2_sigma(r, 0)], m[blake2_sigma(r, 1)])
    HHE_B2G(1, 5, 9, 13, m[blake2_sigma(r, 2)], m[blake2_sigma(r, 3)])
    HHE_B2G(2, 6, 10, 14, m[blake2_sigma(r, 4)], m[blake2_sigma(r, 5)])
    HHE_B2G(3, 7, 11, 15, m[blake2_sigma(r, 6)], m[blake2_sigma(r, 7)])
    HHE_B2G(0, 5, 10, 15, m[blake2_sigma(r, 8)], m[blake2_sigma(r, 9)])
    HHE_B2G(1, 6, 11, 12, m[blake2_sigma(r, 10)], m[blake2_sigma(r, 11)])
    HHE_B2G(2, 7, 8, 13, m[blake2_sigma(r, 12)], m[blake2_sigma(r, 13)])
    HHE_B2G(3, 4, 9, 14, m[blake2_sigma(r, 14)], m[blake2_sigma(r, 15)])
  }
#undef HHE_B2G
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[8 + i];
}

// Parameter block words 0..2 of BLAKE2b (the other five are zero here: no salt, no personalisation):
//   byte 0 digest_length, 1 key_length, 2 fanout, 3 depth, 4-7 leaf_length, 8-11 node_offset, 12-15 xof_length,
//   16 node_depth, 17 inner_length
HD void blake2b_init(u64 *h, u32 digest_len, u32 key_len, u32 fanout, u32 depth, u32 leaf_len, u32 node_offset, u32 xof_len,
                     u32 node_depth, u32 inner_len) {
  const u64 p0 = static_cast<u64>(digest_len) | (static_cast<u64>(key_len) << 8) | (static_cast<u64>(fanout) << 16) |
                 (static_cast<u64>(depth) << 24) | (static_cast<u64>(leaf_len) << 32);
  const u64 p1 = static_cast<u64>(node_offset) | (static_cast<u64>(xof_len) << 32);
  const u64 p2 = static_cast<u64>(node_depth) | (static_cast<u64>(inner_len) << 8);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = blake2_iv(i);
  h[0] ^= p0;
  h[1] ^= p1;
  h[2] ^= p2;
}

// 64-byte output block `block` of blake2xb(out_len = xof_len bytes, in = `counter` (8 bytes, little endian), key = seed[8] (64 bytes)):
// root = BLAKE2b-64 keyed hash of the counter with the XOF parameter block, out block i = BLAKE2b-64(root) with node_offset = i,
// fanout = depth = 0, leaf_length = inner_length = 64 (blake2xb_final of the BLAKE2 reference implementation).
HD void blake2xb_block(const u64 *seed, u64 counter, u32 block, u32 xof_len, u64 *out) {
  u64 h[8], m[16];
  blake2b_init(h, 64, 64, 1, 1, 0, 0, xof_len, 0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m[i] = seed[i];  // the key block: 64 key bytes + 64 zero bytes
    m[8 + i] = 0;
  }
  blake2b_compress(h, m, 128, false);
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = 0;
  m[0] = counter;
  blake2b_compress(h, m, 136, true);  // h = root hash
  u64 c[8];
  blake2b_init(c, 64, 0, 0, 0, 64, block, xof_len, 0, 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m[i] = h[i];
    m[8 + i] = 0;
  }
  blake2b_compress(c, m, 64, true);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = c[i];
}

}  // namespace hhe
