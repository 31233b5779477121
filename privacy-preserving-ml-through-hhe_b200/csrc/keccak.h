// SHAKE128 + PASTA-3 round-material generation on the device (host+device code).
// Replaces Pasta::init_shake / generate_random_field_element / get_random_matrix / calculate_row / get_rc_vec
// (src/pasta/pasta_3_plain.cpp:56-129,286-295) and the XKCP sponge they call (libs/keccak/KeccakHash.c,
// KeccakSponge.inc: rate 1344 bits, suffix 0x1F).  One CTA per SHAKE stream (= PASTA block counter).
#pragma once
#include "devconsts.h"
#include "kernels.h"

namespace hhe {

HD u64 rotl64(u64 x, int n) { return (x << n) | (x >> (64 - n)); }

HD void keccak_f1600(u64 *s) {
  const u64 rc[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
                      0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                      0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
                      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  for (int round = 0; round < 24; ++round) {
    // theta
    u64 c0 = s[0] ^ s[5] ^ s[10] ^ s[15] ^ s[20];
    u64 c1 = s[1] ^ s[6] ^ s[11] ^ s[16] ^ s[21];
    u64 c2 = s[2] ^ s[7] ^ s[12] ^ s[17] ^ s[22];
    u64 c3 = s[3] ^ s[8] ^ s[13] ^ s[18] ^ s[23];
    u64 c4 = s[4] ^ s[9] ^ s[14] ^ s[19] ^ s[24];
    u64 d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1),
        d4 = c3 ^ rotl64(c0, 1);
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
      s[y] ^= d0;
      s[y + 1] ^= d1;
      s[y + 2] ^= d2;
      s[y + 3] ^= d3;
      s[y + 4] ^= d4;
    }
    // rho + pi
    u64 b[25];
    b[0] = s[0];
    b[10] = rotl64(s[1], 1);
    b[20] = rotl64(s[2], 62);
    b[5] = rotl64(s[3], 28);
    b[15] = rotl64(s[4], 27);
    b[16] = rotl64(s[5], 36);
    b[1] = rotl64(s[6], 44);
    b[11] = rotl64(s[7], 6);
    b[21] = rotl64(s[8], 55);
    b[6] = rotl64(s[9], 20);
    b[7] = rotl64(s[10], 3);
    b[17] = rotl64(s[11], 10);
    b[2] = rotl64(s[12], 43);
    b[12] = rotl64(s[13], 25);
    b[22] = rotl64(s[14], 39);
    b[23] = rotl64(s[15], 41);
    b[8] = rotl64(s[16], 45);
    b[18] = rotl64(s[17], 15);
    b[3] = rotl64(s[18], 21);
    b[13] = rotl64(s[19], 8);
    b[14] = rotl64(s[20], 18);
    b[24] = rotl64(s[21], 2);
    b[9] = rotl64(s[22], 61);
    b[19] = rotl64(s[23], 56);
    b[4] = rotl64(s[24], 14);
    // chi
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
      s[y] = b[y] ^ (~b[y + 1] & b[y + 2]);
      s[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]);
      s[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]);
      s[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y]);
      s[y + 4] = b[y + 4] ^ (~b[y] & b[y + 1]);
    }
    s[0] ^= rc[round];
  }
}

constexpr int kShakeRateLanes = 21;  // 168 bytes

struct ShakeStream {
  u64 s[25];
  int lane;
  // seed = BE64(nonce) || BE64(counter): 16 bytes = lanes 0,1 (little-endian lanes -> byte-swap), pad 0x1F / 0x80
  HD void init(u64 nonce, u64 counter) {
    for (int i = 0; i < 25; ++i) s[i] = 0;
    s[0] = bswap(nonce);
    s[1] = bswap(counter);
    s[2] = 0x1F;
    s[kShakeRateLanes - 1] ^= 0x8000000000000000ULL;
    keccak_f1600(s);
    lane = 0;
  }
  static HD u64 bswap(u64 x) {
    x = ((x & 0x00ff00ff00ff00ffULL) << 8) | ((x >> 8) & 0x00ff00ff00ff00ffULL);
    x = ((x & 0x0000ffff0000ffffULL) << 16) | ((x >> 16) & 0x0000ffff0000ffffULL);
    return (x << 32) | (x >> 32);
  }
  // next 8 output bytes interpreted big-endian (be64toh of the squeezed bytes)
  HD u64 next_be64() {
    if (lane == kShakeRateLanes) {
      keccak_f1600(s);
      lane = 0;
    }
    return bswap(s[lane++]);
  }
  // generate_random_field_element: mask to bit-length of p, reject >= p (and 0 if !allow_zero)
  HD u32 field_element(u64 p, u64 mask, bool allow_zero) {
    for (;;) {
      const u64 e = next_be64() & mask;
      if (!allow_zero && e == 0) continue;
      if (e < p) return static_cast<u32>(e);
    }
  }
};

// Generates, for one block: 4 layers x {mat1, mat2} (128x128, row-major) and 4 x 256 round constants.
// Phase 1 (thread 0): squeeze first rows + round constants in stream order mat1, mat2, rc1, rc2 per layer.
// Phase 2: rows 1..127 by the companion recurrence row_i[j] = first[j]*row_{i-1}[127] + row_{i-1}[j-1] mod p,
//          thread j owns column j of one of the 8 matrices (256 threads = 2 matrices at a time).
struct MaterialBody {
  static constexpr const char *kName = "pasta_material";
  const u64 *counters;  // [blocks]
  u64 nonce;
  u32 *out;  // [blocks][kMaterialWords]
  u64 p;
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int T = kPastaT;
    u32 *o = out + static_cast<size_t>(bid) * kMaterialWords;
    u32 *prev = reinterpret_cast<u32 *>(smem);  // [2][T] previous row of the two matrices in flight
    FOR_THREADS(tid, nt) {
      if (tid == 0) {
        u64 mask = 1;
        while (mask <= p) mask <<= 1;
        mask -= 1;
        ShakeStream sh;
        sh.init(nonce, counters[bid]);
        for (int layer = 0; layer < 4; ++layer) {
          for (int m = 0; m < 2; ++m) {
            u32 *row0 = o + (static_cast<size_t>(layer) * 2 + m) * T * T;
            for (int j = 0; j < T; ++j) row0[j] = sh.field_element(p, mask, false);
          }
          u32 *rc = o + kMatWords + layer * 2 * T;
          for (int j = 0; j < 2 * T; ++j) rc[j] = sh.field_element(p, mask, true);
        }
      }
    }
    SYNC();
    for (int layer = 0; layer < 4; ++layer) {
      FOR_THREADS(tid, nt) {
        if (tid < 2 * T) prev[tid] = o[(static_cast<size_t>(layer) * 2 + (tid >> 7)) * T * T + (tid & (T - 1))];
      }
      SYNC();
      for (int i = 1; i < T; ++i) {
        // two-phase update through a second buffer so that the read of prev[j-1] never races the write of prev[j]
        u32 *next = prev + 2 * T + ((i & 1) ? 0 : 2 * T);
        const u32 *cur = (i == 1) ? prev : prev + 2 * T + ((i & 1) ? 2 * T : 0);
        FOR_THREADS(tid, nt) {
          if (tid < 2 * T) {
            const int m = tid >> 7, j = tid & (T - 1);
            const u32 *M = o + (static_cast<size_t>(layer) * 2 + m) * T * T;
            u64 v = static_cast<u64>(M[j]) * cur[m * T + T - 1] % p;
            if (j) v = (v + cur[m * T + j - 1]) % p;
            next[tid] = static_cast<u32>(v);
            o[(static_cast<size_t>(layer) * 2 + m) * T * T + static_cast<size_t>(i) * T + j] = static_cast<u32>(v);
          }
        }
        SYNC();
      }
    }
  }
};
constexpr size_t kMaterialSmem = sizeof(u32) * 2 * kPastaT * 3;

// Plain (non-homomorphic) PASTA-3: Pasta::gen_keystream with its round functions (src/pasta/pasta_3_plain.cpp:156-171,198-282) and
// the word-wise PASTA::encrypt / decrypt (:9-47) for bulk data owners. One CTA per block, consuming the round material
// MaterialBody produced for that block's counter; thread tid < 256 owns state word (half = tid >> 7, i = tid & 127).
struct alignas(16) U32x4 {
  u32 x, y, z, w;
};
struct PastaPlainBody {
  static constexpr const char *kName = "pasta_plain";
  const u32 *material;  // [blocks][kMaterialWords]
  const u64 *key;       // [256] symmetric key words (< p)
  const u64 *in;        // [n_words] plaintext (encrypt) or ciphertext (decrypt)
  u64 *out;
  size_t first_word;    // index of this launch's first word inside in/out
  size_t n_words;       // total length of in/out
  u64 p;
  int decrypt;
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    constexpr int T = kPastaT;
    const u32 *mat = material + static_cast<size_t>(bid) * kMaterialWords;
    u32 *st = reinterpret_cast<u32 *>(smem), *nw = st + 2 * T;
    FOR_THREADS(tid, nt) {
      if (tid < 2 * T) st[tid] = static_cast<u32>(key[tid] % p);
    }
    SYNC();
    for (int layer = 0; layer < 4; ++layer) {
      FOR_THREADS(tid, nt) {  // affine layer: row i of the half's matrix times the half's state, plus round constant
        if (tid < 2 * T) {
          const int half = tid >> 7, i = tid & (T - 1);
          const u32 *row = mat + (static_cast<size_t>(layer) * 2 + half) * T * T + static_cast<size_t>(i) * T;
          const u32 *sv = st + half * T;
          u64 acc = 0;  // 128 products below 2^34 each: no overflow
          for (int j = 0; j < T; j += 4) {
            const U32x4 m4 = *reinterpret_cast<const U32x4 *>(row + j);
            acc += static_cast<u64>(m4.x) * sv[j] + static_cast<u64>(m4.y) * sv[j + 1] + static_cast<u64>(m4.z) * sv[j + 2] +
                   static_cast<u64>(m4.w) * sv[j + 3];
          }
          nw[tid] = static_cast<u32>((acc % p + mat[kMatWords + layer * 2 * T + tid]) % p);
        }
      }
      SYNC();
      FOR_THREADS(tid, nt) {  // mix: (a, b) -> (2a + b, a + 2b)
        if (tid < 2 * T) {
          const int i = tid & (T - 1);
          const u64 sum = (static_cast<u64>(nw[i]) + nw[T + i]) % p;
          st[tid] = static_cast<u32>((nw[tid] + sum) % p);
        }
      }
      SYNC();
      if (layer == 3) break;
      FOR_THREADS(tid, nt) {  // S-box: Feistel x_i += x_{i-1}^2 (rounds 1, 2), cube (round 3), per half
        if (tid < 2 * T) {
          const int i = tid & (T - 1);
          const u64 v = st[tid];
          if (layer == 2) {
            nw[tid] = static_cast<u32>(v * v % p * v % p);
          } else {
            const u64 prev = i ? st[tid - 1] : 0;
            nw[tid] = static_cast<u32>((prev * prev + v) % p);
          }
        }
      }
      SYNC();
      FOR_THREADS(tid, nt) {
        if (tid < 2 * T) st[tid] = nw[tid];
      }
      SYNC();
    }
    FOR_THREADS(tid, nt) {  // the keystream is the first half of the state
      const size_t w = first_word + static_cast<size_t>(bid) * T + tid;
      if (tid < T && w < n_words) {
        const u64 ks = st[tid];
        u64 v = in[w];
        if (decrypt) {
          if (ks > v) v += p;  // exactly PASTA::decrypt (:39-41): no reduction of the input
          out[w] = v - ks;
        } else {
          out[w] = (v + ks) % p;
        }
      }
    }
  }
};
constexpr size_t kPastaPlainSmem = sizeof(u32) * 4 * kPastaT;

}  // namespace hhe
