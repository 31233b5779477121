// SEAL 4.0 wire format at the engine's boundary: the byte streams the reference's gRPC payloads and
// .bin checkpoints carry are seal::Ciphertext / GaloisKeys / RelinKeys ::save output (src/examples/CSP/CSP.cpp:131,201,509,537,587,
// src/examples/Analyst/Analyst.cpp:258-341). This module reads and writes that format directly from/to the engine's SEAL-layout
// buffers, so a service can hand serialized bytes to the GPU engine without instantiating SEAL objects.
//
// Format (libs/seal/include/SEAL-4.0/seal/serialization.h:49-91 and the layouts observed from the reference, pinned by
// tests/test_seal_codec.py against SEAL itself):
//   object      = SEALHeader{u16 magic 0xA15E, u8 header_size 16, u8 major 4, u8 minor 0, u8 compr_mode, u16 reserved, u64 size}
//                 followed by the body, raw (compr_mode 0), zlib-deflated (1) or one zstd frame (2); `size` counts header + body as stored
//   Ciphertext  = object{ parms_id u64[4], is_ntt_form u8, size u64, poly_modulus_degree u64, coeff_modulus_size u64, scale f64,
//                         correction_factor u64, DynArray }
//   DynArray    = object(compr_mode 0){ count u64, count x u64 }
//   KSwitchKeys = object{ parms_id u64[4] (key level), dim1 u64, dim1 x ( dim2 u64, dim2 x PublicKey ) },  PublicKey = Ciphertext object
//                 (compr_mode 0, NTT form, size 2, key level); GaloisKeys index = (galois_elt - 1) / 2, RelinKeys index 0.
//   parms_id    = BLAKE2b-256 over the u64 words [scheme (BFV = 1), N, q_0 .. q_{n-1}, t] (EncryptionParameters::compute_parms_id)
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace hhe {
namespace sealio {

struct Ring {
  uint64_t N, t;
  std::vector<uint64_t> q;  // key-level primes (last = special prime)
  int L() const { return static_cast<int>(q.size()) - 1; }
  int K() const { return static_cast<int>(q.size()); }
};

enum { kComprNone = 0, kComprZlib = 1, kComprZstd = 2 };

void blake2b_256(const void *in, size_t len, uint8_t out[32]);
// level 0: first (data-level) parameters, 1: key-level parameters
void parms_id(const Ring &r, int level, uint64_t out[4]);

size_t ct_save_bound(const Ring &r, int size);
// returns bytes written; throws std::invalid_argument (bad arguments / buffer too small), std::logic_error (compression failed)
size_t ct_save(const Ring &r, const uint64_t *ct, int size, int compr, uint8_t *out, size_t cap);
// Ciphertext::load(context, ...): validates header, parms_id, dimensions and residue ranges (std::logic_error otherwise).
// dst is filled by `sink(ptr, words)` so the caller can direct the payload into pinned memory. Returns bytes consumed.
size_t ct_load(const Ring &r, const uint8_t *in, size_t len, uint64_t *ct, size_t cap_words, int *size);

// Walks a serialized GaloisKeys / RelinKeys object; for every non-empty index calls on_key(index, ksk) with ksk laid out
// [L digits][2][K][N] (the engine's hhe_load_ksk layout). Returns bytes consumed.
size_t keys_walk(const Ring &r, const uint8_t *in, size_t len, const std::function<void(uint64_t index, const uint64_t *ksk)> &on_key);

}  // namespace sealio
}  // namespace hhe
