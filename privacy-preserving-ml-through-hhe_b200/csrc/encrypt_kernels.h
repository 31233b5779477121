// Kernels of the GPU port of seal::Encryptor::encrypt (public-key BFV encryption): the client-side
// operation behind sealhelper::encrypt_weight_mat (src/util/sealhelper.cpp:123-142) and pastahelper::encrypt_symmetric_key
// (src/util/pastahelper.cpp:355-377). SEAL 4.0's algorithm (seal/encryptor.h:132, seal/util/rlwe.h:52-105), per ciphertext:
//   prng  <- Blake2xbPRNG(seed)                                      a fresh generator per encryption
//   u     <- ternary:  per coefficient one 32-bit draw x, r = (3 x) >> 32, redrawn when x = 0 (libstdc++'s
//                      uniform_int_distribution<uint64_t>(0, 2) on a 32-bit engine: Lemire's method, as compiled into libseal-4.0.a)
//   c_j   =  INTT(NTT(u) (.) pk_j) + e_j,  j = 0, 1,  at the KEY level (K limbs); e_j <- centred binomial from 6 bytes:
//            popcount(b0) + popcount(b1) + popcount(b2 & 31) - popcount(b3) - popcount(b4) - popcount(b5 & 31)
//            (draw order: all of u, then all of e_0, then all of e_1)
//   (c_0, c_1) <- rounding division by the special prime (divide_and_round_q_last: the ModDown of a key switch), data level
//   c_0  += Delta-scaled plaintext (multiply_add_plain_with_scaling_variant, the add_plain of SURVEY.md A.4)
// Bit-exact with SEAL for the same seed (tests compare with the reference's Encryptor under a seeded Blake2xbPRNGFactory).
#pragma once
#include "blake2.h"
#include "devconsts.h"
#include "modarith.h"

namespace hhe {

constexpr u32 kPrngRefillBytes = 4096;  // Blake2xbPRNG buffer size (seal/randomgen.h)

// stream bytes one encryption can consume: 4 N (u) + 2 * 6 N (noise) + one spare refill for redrawn ternary samples
HD size_t enc_stream_refills(size_t N) { return (16 * N + kPrngRefillBytes - 1) / kPrngRefillBytes + 1; }

// The generator's byte stream: thread per 64-byte block of every refill of every ciphertext's generator.
struct PrngStreamBody {
  static constexpr const char *kName = "prng_stream";
  const u64 *seeds;  // [count][8]
  u64 *stream;       // [count][refills * 512] (64-bit words, little endian = the byte stream)
  size_t refills, total;  // total = count * refills * 64 blocks
  HD void operator()(int bid, int nt, unsigned char *) const {
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const size_t per = refills * 64, ct = g / per, r = (g % per) / 64;
        const u32 blk = static_cast<u32>(g % 64);
        blake2xb_block(seeds + ct * 8, r, blk, kPrngRefillBytes, stream + g * 8);
      }
    }
  }
};

HD u32 stream_u32(const u64 *s, size_t word) { return static_cast<u32>(s[word >> 1] >> (32 * (word & 1))); }
HD u32 stream_u8(const u64 *s, size_t byte) { return static_cast<u32>(s[byte >> 3] >> (8 * (byte & 7))) & 0xffu; }
HD int popc8(u32 v) {
  v = (v & 0x55u) + ((v >> 1) & 0x55u);
  v = (v & 0x33u) + ((v >> 2) & 0x33u);
  return static_cast<int>((v + (v >> 4)) & 0x0fu);
}

// u and the two noise polynomials of one ciphertext, as residues of every key-level limb. CTA per ciphertext.
// Almost always no ternary draw is rejected (probability N 2^-32 per ciphertext) and coefficient j reads stream word j; otherwise
// thread 0 rebuilds the word index of every coefficient sequentially before the CTA continues.
struct EncSampleBody {
  static constexpr const char *kName = "enc_sample";
  const u64 *stream;  // [count][words]
  size_t words;       // 64-bit words per ciphertext stream
  u64 *u;             // [count][K][N]
  u64 *e;             // [count][2][K][N]
  u32 *pos;           // [count][N] scratch for the rare path
  const DevConsts *C;
  HD void operator()(int bid, int nt, unsigned char *smem) const {
    const size_t N = C->N;
    const int K = C->K;
    const u64 *s = stream + static_cast<size_t>(bid) * words;
    u32 *flag = reinterpret_cast<u32 *>(smem);  // [0] any rejected draw, [1] 32-bit words the ternary polynomial consumed
    u32 *p = pos + static_cast<size_t>(bid) * N;
    FOR_THREADS(tid, nt) {
      if (tid == 0) flag[0] = 0, flag[1] = static_cast<u32>(N);
    }
    SYNC();
    FOR_THREADS(tid, nt) {
      for (size_t j = tid; j < N; j += nt)
        if (stream_u32(s, j) == 0) flag[0] = 1;
    }
    SYNC();
    FOR_THREADS(tid, nt) {
      if (tid == 0 && flag[0]) {
        size_t w = 0;
        for (size_t j = 0; j < N; ++j) {
          while (stream_u32(s, w) == 0) ++w;
          p[j] = static_cast<u32>(w++);
        }
        flag[1] = static_cast<u32>(w);
      }
    }
    SYNC();
    const bool slow = flag[0] != 0;
    const size_t e_base = static_cast<size_t>(flag[1]) * 4;  // byte offset of the first noise sample
    FOR_THREADS(tid, nt) {
      for (size_t j = tid; j < N; j += nt) {
        const u32 x = stream_u32(s, slow ? p[j] : j);
        const u32 r = static_cast<u32>((static_cast<u64>(x) * 3) >> 32);  // 0, 1, 2 -> -1, 0, 1
        for (int k = 0; k < K; ++k) u[(static_cast<size_t>(bid) * K + k) * N + j] = r ? r - 1 : C->mod[k].q - 1;
        for (int c = 0; c < 2; ++c) {
          const size_t b0 = e_base + (static_cast<size_t>(c) * N + j) * 6;
          const int noise = popc8(stream_u8(s, b0)) + popc8(stream_u8(s, b0 + 1)) + popc8(stream_u8(s, b0 + 2) & 0x1f) -
                            popc8(stream_u8(s, b0 + 3)) - popc8(stream_u8(s, b0 + 4)) - popc8(stream_u8(s, b0 + 5) & 0x1f);
          for (int k = 0; k < K; ++k)
            e[((static_cast<size_t>(bid) * 2 + c) * K + k) * N + j] = noise < 0 ? C->mod[k].q - static_cast<u64>(-noise) : static_cast<u64>(noise);
        }
      }
    }
  }
};

// c[ct][j][k] = NTT(u)[ct][k] (.) pk[j][k]   (element-wise, NTT domain, key level)
struct EncPkMulBody {
  static constexpr const char *kName = "enc_pk_mul";
  const u64 *u_ntt;  // [count][K][N]
  const u64 *pk;     // [2][K][N] NTT form (seal::PublicKey::data())
  u64 *c;            // [count][2][K][N]
  const DevConsts *C;
  size_t total;  // count * 2 * K * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const size_t N = C->N;
    const u32 K = static_cast<u32>(C->K);
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) {
        const u32 limb = static_cast<u32>(g >> C->logn);
        const u32 k = limb % K, comp = (limb / K) & 1, ct = limb / (2 * K);
        const size_t j = g & (N - 1);
        c[g] = mul_mod(u_ntt[(static_cast<size_t>(ct) * K + k) * N + j], pk[(static_cast<size_t>(comp) * K + k) * N + j], C->mod[k]);
      }
    }
  }
};

// c += e (coefficient form, key level)
struct EncAddNoiseBody {
  static constexpr const char *kName = "enc_add_noise";
  u64 *c;
  const u64 *e;
  const DevConsts *C;
  size_t total;  // count * 2 * K * N
  HD void operator()(int bid, int nt, unsigned char *) const {
    const u32 K = static_cast<u32>(C->K);
    FOR_THREADS(tid, nt) {
      const size_t g = static_cast<size_t>(bid) * nt + tid;
      if (g < total) c[g] = add_mod(c[g], e[g], C->mod[static_cast<u32>(g >> C->logn) % K].q);
    }
  }
};

}  // namespace hhe
