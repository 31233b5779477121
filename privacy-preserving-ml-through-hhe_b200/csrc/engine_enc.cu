// Engine implementation (5/5): public-key BFV encryption on the GPU (seal::Encryptor::encrypt).
#include "encrypt_kernels.h"
#include "engine_impl.h"

namespace hhe {

// out[ct] = Encryptor(context, pk).encrypt(plain[ct]) with the generator seeded by seeds[ct] (prng_seed_type, 8 x u64).
// d_pk: seal::PublicKey::data() = [2][K][N], NTT form, key level. d_pt: [count][N] coefficients in [0, t). All device pointers.
void Engine::encrypt(const u64 *d_pk, const u64 *d_seeds, const u64 *d_pt, size_t count, u64 *d_out) {
  if (!count) return;
  Scope sc(*this);
  const size_t N = P_.N;
  const int K = P_.K;
  const size_t refills = enc_stream_refills(N), words = refills * (kPrngRefillBytes / 8);
  const size_t key_words = static_cast<size_t>(K) * N;
  u64 *stream = scratch(count * words), *u = scratch(count * key_words), *e = scratch(count * 2 * key_words),
      *c = scratch(count * 2 * key_words);
  u32 *pos = reinterpret_cast<u32 *>(scratch((count * N + 1) / 2));
  PrngStreamBody prng{d_seeds, stream, refills, count * refills * 64};
  dev_.launch(prng, ew_grid(count * refills * 64), kEwThreads, 0);
  EncSampleBody smp{stream, words, u, e, pos, dC_};
  dev_.launch(smp, count, 256, 16);
  ntt(u, u, count, K, map_mod(K, K, 0), false);
  EncPkMulBody mul{u, d_pk, c, dC_, count * 2 * key_words};
  dev_.launch(mul, ew_grid(count * 2 * key_words), kEwThreads, 0);
  ntt(c, c, count, 2 * K, map_mod(2 * K, K, 0), true);
  EncAddNoiseBody add{c, e, dC_, count * 2 * key_words};
  dev_.launch(add, ew_grid(count * 2 * key_words), kEwThreads, 0);
  // divide_and_round_q_last (seal/util/rns.h): the rounding ModDown of a key switch without addend
  ModDownBody md{c, nullptr, nullptr, 0, d_out, dC_, count * N};
  dev_.launch(md, ew_grid(count * N), kEwThreads, 0);
  add_plain(d_out, d_pt, N, d_out, count, false);
}

}  // namespace hhe
