// C ABI of libhhe_b200.so (include/hhe_b200.h): host-buffer marshalling around hhe::Engine.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hhe_b200.h"
#include "engine.h"
#include "seal_codec.h"

using namespace hhe;

struct hhe_ctx {
  std::unique_ptr<Engine> eng;
};

namespace {

thread_local std::string g_error;

// Every entry point runs with its context's device current and puts the caller's device back afterwards: contexts on
// different GPUs may be used from one process (scratch allocations, key uploads and launches all follow the current device).
struct DeviceScope {
  int prev = -1;
  ~DeviceScope() {
#ifdef HHE_CUDA
    if (prev >= 0) cudaSetDevice(prev);
#endif
  }
};
thread_local DeviceScope *g_scope = nullptr;

template <class F>
int guarded(F &&f) {
  DeviceScope scope;
  DeviceScope *outer = g_scope;
  g_scope = &scope;
  struct Restore {
    DeviceScope *o;
    ~Restore() { g_scope = o; }
  } restore{outer};
  try {
    f();
    return HHE_OK;
  } catch (const std::invalid_argument &e) {
    g_error = e.what();
    return HHE_ERR_INVALID;
  } catch (const std::logic_error &e) {
    g_error = e.what();
    return HHE_ERR_LOGIC;
  } catch (const std::exception &e) {
    g_error = e.what();
    return g_error.rfind("NO_DEVICE", 0) == 0 ? HHE_ERR_NO_DEVICE : HHE_ERR_RUNTIME;
  }
}

Engine &E(hhe_ctx *c) {
  if (!c || !c->eng) throw std::invalid_argument("null context");
#ifdef HHE_CUDA
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur != c->eng->device()) {
    if (g_scope && g_scope->prev < 0) g_scope->prev = cur;
    cuda_check(cudaSetDevice(c->eng->device()), "cudaSetDevice");
  }
#endif
  return *c->eng;
}

u64 *up(Engine &e, const u64 *host, size_t words) {
  u64 *d = e.scratch(words);
  e.dev().h2d(d, host, words * 8);
  return d;
}

void down(Engine &e, u64 *host, const u64 *d, size_t words) { e.dev().d2h(host, d, words * 8); }

// Results of a chunked call leave through two device buffers and the copy stream (Device::d2h_begin / d2h_end): the
// device-to-host copy of chunk i overlaps the computation of chunk i + 1. `buf(i)` is the device buffer chunk i writes to
// (its previous contents have been delivered), `send(i, ...)` starts the delivery, the destructor / finish() waits for all.
struct OverlappedOut {
  Engine &e;
  u64 *d[2];
  OverlappedOut(Engine &eng, size_t words_per_chunk, bool two) : e(eng) {
    d[0] = e.scratch(words_per_chunk);
    d[1] = two ? e.scratch(words_per_chunk) : d[0];
  }
  u64 *buf(size_t i) {
    e.dev().d2h_end(static_cast<int>(i & 1));
    return d[i & 1];
  }
  void send(size_t i, u64 *host, size_t words) { e.dev().d2h_begin(static_cast<int>(i & 1), host, d[i & 1], words * 8); }
  void finish() {
    e.dev().d2h_end(0);
    e.dev().d2h_end(1);
    e.dev().sync();
  }
  ~OverlappedOut() {
    try {
      finish();
    } catch (...) {
    }
  }
};

// run `fn(first, n)` over [0, count) in chunks of the engine's batch limit
template <class F>
void chunked(Engine &e, size_t count, F &&fn) {
  const size_t step = static_cast<size_t>(std::max(1, e.batch_limit()));
  for (size_t off = 0; off < count; off += step) fn(off, std::min(step, count - off));
}

}  // namespace

extern "C" {

const char *hhe_last_error(void) { return g_error.c_str(); }
#ifdef HHE_CUDA
const char *hhe_version(void) { return "hhe_b200 0.2 (sm_100a)"; }
int hhe_build_is_cuda(void) { return 1; }
#else
const char *hhe_version(void) { return "hhe_b200 0.2 (HOST EMULATION: test harness of the kernel index arithmetic, not a product build)"; }
int hhe_build_is_cuda(void) { return 0; }
#endif

int hhe_ctx_create(hhe_ctx **out, uint64_t N, uint64_t t, const uint64_t *q, int nq, int device, void *stream) {
  if (!out) return HHE_ERR_INVALID;
  *out = nullptr;
  return guarded([&] {
    Params p = Params::derive(N, t, q, nq);
    auto ctx = std::make_unique<hhe_ctx>();
    ctx->eng = std::make_unique<Engine>(p, device, stream);
    *out = ctx.release();
  });
}

void hhe_ctx_destroy(hhe_ctx *ctx) {
  guarded([&] {
    if (ctx && ctx->eng) E(ctx);  // the engine's device current while its memory is released; the caller's restored afterwards
    delete ctx;
  });
}

int hhe_ctx_info(const hhe_ctx *ctx, uint64_t *info) {
  return guarded([&] {
    Engine &e = E(const_cast<hhe_ctx *>(ctx));
    info[0] = e.params().N;
    info[1] = e.params().L;
    info[2] = e.params().K;
    info[3] = e.params().t;
    info[4] = e.batch_limit();
    info[5] = e.dev().sm_count;
    info[6] = 0;
    for (int i = 0; i < 2 * e.params().K; ++i) info[6] += table_is_f64(e.params(), i) ? 1 : 0;
  });
}

void *hhe_ctx_stream(const hhe_ctx *ctx) {
#ifdef HHE_CUDA
  return ctx && ctx->eng ? static_cast<void *>(ctx->eng->dev().stream) : nullptr;
#else
  (void)ctx;
  return nullptr;
#endif
}

int hhe_set_batch(hhe_ctx *ctx, int blocks) {
  return guarded([&] {
    Engine &e = E(ctx);
    if (blocks < 0) throw std::invalid_argument("batch must be >= 0");
    e.set_batch(blocks ? blocks : e.auto_batch());
  });
}

uint32_t hhe_galois_elt(const hhe_ctx *ctx, int step) {
  return ctx && ctx->eng ? ctx->eng->params().galois_elt_from_step(step) : 0;
}

int hhe_ctx_constants(const hhe_ctx *ctx, uint64_t *out) {
  return guarded([&] {
    const Params &p = E(const_cast<hhe_ctx *>(ctx)).params();
    size_t o = 0;
    for (int i = 0; i < p.K; ++i) out[o++] = p.tab[i].psi;
    out[o++] = p.tab[p.tab_plain()].psi;
    out[o++] = p.m_sk;
    out[o++] = p.gamma;
    out[o++] = p.m_tilde;
    for (int i = 0; i < p.L; ++i) out[o++] = p.bsk[i];
    for (int i = 0; i <= p.L; ++i) out[o++] = p.tab[p.tab_bsk(i)].psi;
  });
}

int hhe_load_ksk(hhe_ctx *ctx, int kind, uint32_t galois_elt, const uint64_t *ksk) {
  return guarded([&] {
    if (!ksk) throw std::invalid_argument("null key");
    E(ctx).load_ksk(kind, galois_elt, ksk);
  });
}

int hhe_clear_keyset(hhe_ctx *ctx, int kind) {
  return guarded([&] { E(ctx).clear_keyset(kind); });
}

int hhe_has_ksk(const hhe_ctx *ctx, int kind, uint32_t galois_elt) {
  return ctx && ctx->eng && ctx->eng->find_key(kind, galois_elt) ? 1 : 0;
}

uint64_t hhe_launch_count(const hhe_ctx *ctx) { return ctx && ctx->eng ? ctx->eng->dev().launches : 0; }

int hhe_profile_enable(hhe_ctx *ctx, int on) {
  return guarded([&] {
    E(ctx).dev().profile_resolve();
    E(ctx).dev().profiling = on != 0;
  });
}
int hhe_profile_reset(hhe_ctx *ctx) {
  return guarded([&] { E(ctx).dev().profile_reset(); });
}
int hhe_profile_report(hhe_ctx *ctx, char *buf, size_t cap) {
  return guarded([&] {
    Device &d = E(ctx).dev();
    d.profile_resolve();
    std::string js = "{";
    for (size_t i = 0; i < d.stats.size(); ++i) {
      if (i) js += ", ";
      js += "\"" + std::string(d.stats[i].name) + "\": {\"launches\": " + std::to_string(d.stats[i].launches) +
            ", \"ms\": " + std::to_string(d.stats[i].ms) + "}";
    }
    js += "}";
    if (js.size() + 1 > cap) throw std::invalid_argument("profile report buffer too small");
    std::memcpy(buf, js.c_str(), js.size() + 1);
  });
}

// ---------------------------------------------------------------------------------------------- primitives
int hhe_ntt(hhe_ctx *ctx, int limb, int inverse, uint64_t *data, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    const Params &p = e.params();
    if (limb < 0 || limb >= 2 * p.K) throw std::invalid_argument("limb index out of range");
    Engine::Scope sc(e);
    const size_t words = count * p.N;
    u64 *d = up(e, data, words);
    TabMap m{};
    m.id[0] = static_cast<unsigned char>(limb);
    e.ntt(d, d, count, 1, m, inverse != 0);
    down(e, data, d, words);
    e.dev().sync();
  });
}

int hhe_encode(hhe_ctx *ctx, const uint64_t *slots, size_t n_slots, uint64_t *pt, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    const Params &p = e.params();
    if (n_slots > p.N) throw std::invalid_argument("values_matrix size exceeds slot count");
    for (size_t i = 0; i < n_slots * count; ++i)
      if (slots[i] >= p.t) throw std::invalid_argument("input value is larger than plain_modulus");
    Engine::Scope sc(e);
    u64 *ds = up(e, slots, std::max<size_t>(1, n_slots * count));
    u64 *dp = e.scratch(count * p.N);
    e.encode_slots(ds, n_slots, nullptr, static_cast<u32>(n_slots), dp, count);
    down(e, pt, dp, count * p.N);
    e.dev().sync();
  });
}

int hhe_add(hhe_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words();
    u64 *da = up(e, a, w), *db = up(e, b, w);
    e.add(da, db, da, count);
    down(e, out, da, w);
    e.dev().sync();
  });
}

int hhe_negate(hhe_ctx *ctx, const uint64_t *a, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words();
    u64 *da = up(e, a, w);
    e.negate(da, da, count);
    down(e, out, da, w);
    e.dev().sync();
  });
}

int hhe_add_plain(hhe_ctx *ctx, const uint64_t *a, const uint64_t *pt, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words(), N = e.params().N;
    u64 *da = up(e, a, w), *dp = up(e, pt, count * N);
    e.add_plain(da, dp, N, da, count, false);
    down(e, out, da, w);
    e.dev().sync();
  });
}

int hhe_multiply_plain(hhe_ctx *ctx, const uint64_t *a, const uint64_t *pt, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words(), N = e.params().N;
    // SEAL's multiply_plain: a zero plaintext is refused; a MONOMIAL plaintext (one nonzero coefficient) takes the
    // negacyclic_multiply_poly_mono_coeffmod branch, which with fast plain lift multiplies by the coefficient as it is
    // (no centred lift even in the upper half): flagged per item for the lift kernel
    std::vector<u32> mono(count, 0);
    bool any_mono = false;
    for (size_t it = 0; it < count; ++it) {
      size_t nz = 0;
      for (size_t j = 0; j < N; ++j) nz += pt[it * N + j] != 0;
      if (!nz) throw std::logic_error("result ciphertext is transparent");
      mono[it] = nz == 1;
      any_mono = any_mono || nz == 1;
    }
    u64 *da = up(e, a, w), *dp = up(e, pt, count * N), *dout = e.scratch(w);
    u32 *dm = nullptr;
    if (any_mono) {
      dm = reinterpret_cast<u32 *>(e.scratch((count + 1) / 2));
      e.dev().h2d(dm, mono.data(), count * 4);
    }
    e.multiply_plain(da, dp, N, dout, count, dm);
    down(e, out, dout, w);
    e.dev().sync();
  });
}

int hhe_rotate_rows(hhe_ctx *ctx, const uint64_t *a, int steps, int keyset, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words();
    u64 *da = up(e, a, w), *dout = e.scratch(w);
    e.rotate_rows(da, steps, keyset, dout, count);
    down(e, out, dout, w);
    e.dev().sync();
  });
}

int hhe_rotate_columns(hhe_ctx *ctx, const uint64_t *a, int keyset, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words();
    u64 *da = up(e, a, w), *dout = e.scratch(w);
    e.rotate_columns(da, keyset, dout, count);
    down(e, out, dout, w);
    e.dev().sync();
  });
}

int hhe_multiply(hhe_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out3, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words();
    u64 *da = up(e, a, w), *db = (a == b) ? da : up(e, b, w), *dout = e.scratch(count * e.ct_words(3));
    e.multiply(da, db, dout, count);
    down(e, out3, dout, count * e.ct_words(3));
    e.dev().sync();
  });
}

int hhe_square(hhe_ctx *ctx, const uint64_t *a, uint64_t *out3, size_t count) { return hhe_multiply(ctx, a, a, out3, count); }

int hhe_relinearize(hhe_ctx *ctx, const uint64_t *a3, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    u64 *da = up(e, a3, count * e.ct_words(3)), *dout = e.scratch(count * e.ct_words());
    e.relinearize(da, dout, count);
    down(e, out, dout, count * e.ct_words());
    e.dev().sync();
  });
}

int hhe_exponentiate3(hhe_ctx *ctx, const uint64_t *a, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words();
    u64 *da = up(e, a, w), *dout = e.scratch(w);
    e.exponentiate3(da, dout, count);
    down(e, out, dout, w);
    e.dev().sync();
  });
}

// ---------------------------------------------------------------------------------------------- hot path
static void decompose_host(Engine &e, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n_words, size_t records,
                           uint64_t nonce, uint64_t first_counter, int use_bsgs, uint64_t *out) {
  const Params &p = e.params();
  if (!enc_key || !sym_ct || !out) throw std::invalid_argument("null buffer");
  if (n_words == 0 || records == 0) return;
  for (size_t i = 0; i < n_words * records; ++i)
    if (sym_ct[i] >= p.t) throw std::invalid_argument("input value is larger than plain_modulus");
  const size_t bpr = (n_words + kPastaT - 1) / kPastaT, nblocks = bpr * records;
  std::vector<u64> sym(nblocks * kPastaT, 0), counters(nblocks);
  std::vector<u32> lens(nblocks);
  for (size_t r = 0; r < records; ++r)
    for (size_t b = 0; b < bpr; ++b) {
      const size_t cnt = std::min<size_t>(kPastaT, n_words - b * kPastaT), blk = r * bpr + b;
      std::memcpy(&sym[blk * kPastaT], sym_ct + r * n_words + b * kPastaT, cnt * 8);
      lens[blk] = static_cast<u32>(cnt);
      counters[blk] = first_counter + b;
    }
  Engine::Scope sc(e);
  const size_t ctw = e.ct_words(), step = static_cast<size_t>(std::max(1, e.batch_limit()));
  u64 *d_key = up(e, enc_key, ctw);
  // all inputs go up once (128 words + a length per block); every chunk's ciphertexts come back on the copy stream while the
  // next chunk is being transciphered
  u64 *d_sym = up(e, sym.data(), nblocks * kPastaT);
  u32 *d_lens = reinterpret_cast<u32 *>(e.scratch((nblocks + 1) / 2));
  e.dev().h2d(d_lens, lens.data(), nblocks * 4);
  // records restart their counters: the bpr keystream ciphertexts are the same for every record and are computed once per call
  // (Engine::pasta_keystreams); every block then only subtracts its keystream from its own encoded words
  u64 *d_ks = nullptr;
  u32 *d_idx = nullptr;
  if (records > 1 && e.share_keystreams()) {
    std::vector<u64> uniq(counters.begin(), counters.begin() + bpr);
    std::vector<u32> idx(nblocks);
    for (size_t blk = 0; blk < nblocks; ++blk) idx[blk] = static_cast<u32>(blk % bpr);
    d_ks = e.scratch(bpr * ctw);
    e.pasta_keystreams(d_key, uniq, nonce, use_bsgs != 0, d_ks);
    d_idx = reinterpret_cast<u32 *>(e.scratch((nblocks + 1) / 2));
    e.dev().h2d(d_idx, idx.data(), nblocks * 4);
    e.dev().sync();
  }
  OverlappedOut oo(e, std::min(step, nblocks) * ctw, nblocks > step);
  size_t chunk = 0;
  chunked(e, nblocks, [&](size_t off, size_t nb) {
    Engine::Scope inner(e);
    u64 *d_out = oo.buf(chunk);
    if (d_ks) {
      e.pasta_finish(d_ks, d_idx + off, d_sym + off * kPastaT, d_lens + off, nb, d_out);
    } else {
      std::vector<u64> ctr(counters.begin() + off, counters.begin() + off + nb);
      e.pasta_decompose(d_key, d_sym + off * kPastaT, d_lens + off, ctr, nonce, use_bsgs != 0, d_out);
    }
    oo.send(chunk++, out + off * ctw, nb * ctw);
  });
  oo.finish();
}

int hhe_pasta3_decompose(hhe_ctx *ctx, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n_words, uint64_t nonce,
                         uint64_t first_counter, int use_bsgs, uint64_t *out) {
  return guarded([&] { decompose_host(E(ctx), enc_key, sym_ct, n_words, 1, nonce, first_counter, use_bsgs, out); });
}

int hhe_pasta3_decompose_records(hhe_ctx *ctx, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n_words,
                                 size_t records, uint64_t nonce, uint64_t first_counter, int use_bsgs, uint64_t *out) {
  return guarded([&] { decompose_host(E(ctx), enc_key, sym_ct, n_words, records, nonce, first_counter, use_bsgs, out); });
}

// BaseCSP::decompose (src/examples/CSP/CSP.cpp:235-283) as one call: every record is transciphered (counters restart per
// record), the last block is optionally masked, and the record's blocks are flattened into one ciphertext. The
// intermediate ciphertexts never leave HBM. apply_mask = 0 reproduces the reference service (its mask is applied to a
// copy, CSP.cpp:262-269, so it has no effect); apply_mask = 1 masks in place like the monolithic demo
// (src/examples/hhe_pktnn_examples.cpp:620-624).
int hhe_csp_decompose(hhe_ctx *ctx, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n_words, size_t records, uint64_t nonce,
                      int use_bsgs, int apply_mask, int flatten_keyset, uint64_t *out) {
  return guarded([&] {
    Engine &e = E(ctx);
    const Params &p = e.params();
    if (!enc_key || !sym_ct || !out) throw std::invalid_argument("null buffer");
    if (n_words == 0 || records == 0) return;
    for (size_t i = 0; i < n_words * records; ++i)
      if (sym_ct[i] >= p.t) throw std::invalid_argument("input value is larger than plain_modulus");
    const size_t bpr = (n_words + kPastaT - 1) / kPastaT, rem = n_words % kPastaT, ctw = e.ct_words();
    Engine::Scope sc(e);
    u64 *d_key = up(e, enc_key, ctw);
    u64 *d_ones = nullptr;
    if (apply_mask && rem) {
      std::vector<u64> ones(rem, 1);
      d_ones = up(e, ones.data(), rem);
    }
    // every record uses counters 0 .. bpr-1: their keystream ciphertexts are computed once for the whole call
    u64 *d_ks = nullptr;
    if (records > 1 && e.share_keystreams()) {
      std::vector<u64> uniq(bpr);
      for (size_t b = 0; b < bpr; ++b) uniq[b] = b;
      d_ks = e.scratch(bpr * ctw);
      e.pasta_keystreams(d_key, uniq, nonce, use_bsgs != 0, d_ks);
    }
    // chunk over whole records so a chunk's blocks can be flattened on the device
    const size_t rec_per_chunk = std::max<size_t>(1, static_cast<size_t>(e.batch_limit()) / bpr);
    OverlappedOut oo(e, std::min(rec_per_chunk, records) * ctw, records > rec_per_chunk);
    size_t chunk = 0;
    for (size_t r0 = 0; r0 < records; r0 += rec_per_chunk) {
      const size_t nr = std::min(rec_per_chunk, records - r0), nb = nr * bpr;
      Engine::Scope inner(e);
      std::vector<u64> sym(nb * kPastaT, 0), ctr(nb);
      std::vector<u32> lens(nb);
      for (size_t r = 0; r < nr; ++r)
        for (size_t b = 0; b < bpr; ++b) {
          const size_t cnt = std::min<size_t>(kPastaT, n_words - b * kPastaT), blk = r * bpr + b;
          std::memcpy(&sym[blk * kPastaT], sym_ct + (r0 + r) * n_words + b * kPastaT, cnt * 8);
          lens[blk] = static_cast<u32>(cnt);
          ctr[blk] = b;
        }
      u64 *d_sym = up(e, sym.data(), nb * kPastaT);
      u32 *d_lens = reinterpret_cast<u32 *>(e.scratch((nb + 1) / 2));
      e.dev().h2d(d_lens, lens.data(), nb * 4);
      u64 *d_blocks = e.scratch(nb * ctw), *d_flat = oo.buf(chunk);
      if (d_ks) {
        std::vector<u32> idx(nb);
        for (size_t blk = 0; blk < nb; ++blk) idx[blk] = static_cast<u32>(blk % bpr);
        u32 *d_idx = reinterpret_cast<u32 *>(e.scratch((nb + 1) / 2));
        e.dev().h2d(d_idx, idx.data(), nb * 4);
        e.dev().sync();
        e.pasta_finish(d_ks, d_idx, d_sym, d_lens, nb, d_blocks);
      } else {
        e.pasta_decompose(d_key, d_sym, d_lens, ctr, nonce, use_bsgs != 0, d_blocks);
      }
      if (d_ones) {  // mask the last block of every record in place
        u64 *last = e.scratch(nr * ctw), *masked = e.scratch(nr * ctw);
        e.strided_copy(d_blocks + (bpr - 1) * ctw, bpr * ctw, last, ctw, ctw, nr);
        e.mask(last, d_ones, static_cast<u32>(rem), masked, nr);
        e.strided_copy(masked, ctw, d_blocks + (bpr - 1) * ctw, bpr * ctw, ctw, nr);
      }
      e.flatten(d_blocks, bpr, flatten_keyset, d_flat, nr);
      oo.send(chunk++, out + r0 * ctw, nr * ctw);
    }
    oo.finish();
  });
}

int hhe_mask(hhe_ctx *ctx, const uint64_t *a, const uint64_t *mask, size_t n_mask, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    if (n_mask > e.params().N) throw std::invalid_argument("values_matrix size exceeds slot count");
    bool zero = true;
    for (size_t i = 0; i < n_mask; ++i) {
      if (mask[i] >= e.params().t) throw std::invalid_argument("input value is larger than plain_modulus");
      zero = zero && mask[i] == 0;
    }
    if (zero) throw std::logic_error("result ciphertext is transparent");
    Engine::Scope sc(e);
    const size_t w = count * e.ct_words();
    u64 *da = up(e, a, w), *dm = up(e, mask, std::max<size_t>(1, n_mask)), *dout = e.scratch(w);
    e.mask(da, dm, static_cast<u32>(n_mask), dout, count);
    down(e, out, dout, w);
    e.dev().sync();
  });
}

int hhe_flatten(hhe_ctx *ctx, const uint64_t *in, size_t per, int keyset, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    if (per == 0) throw std::invalid_argument("flatten needs at least one ciphertext per group");
    Engine::Scope sc(e);
    const size_t ctw = e.ct_words();
    u64 *din = up(e, in, count * per * ctw), *dout = e.scratch(count * ctw);
    e.flatten(din, per, keyset, dout, count);
    down(e, out, dout, count * ctw);
    e.dev().sync();
  });
}

int hhe_vec_sum(hhe_ctx *ctx, const uint64_t *a, size_t n, int keyset, uint64_t *out, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    const size_t ctw = e.ct_words(), step = static_cast<size_t>(std::max(1, e.batch_limit()));
    Engine::Scope outer(e);
    OverlappedOut oo(e, std::min(step, count) * ctw, count > step);
    size_t chunk = 0;
    chunked(e, count, [&](size_t off, size_t nb) {
      Engine::Scope sc(e);
      u64 *da = up(e, a + off * ctw, nb * ctw), *dout = oo.buf(chunk);
      e.vec_sum(da, n, keyset, dout, nb);
      oo.send(chunk++, out + off * ctw, nb * ctw);
    });
    oo.finish();
  });
}

int hhe_fc_rows(hhe_ctx *ctx, const uint64_t *x, size_t samples, const uint64_t *w, size_t rows, size_t n, int keyset,
                uint64_t *out) {
  return guarded([&] {
    Engine &e = E(ctx);
    const size_t ctw = e.ct_words();
    Engine::Scope sc(e);
    u64 *dw = up(e, w, rows * ctw);
    // items are (sample, row) pairs; chunk over samples so a chunk holds whole rows-groups
    const size_t per_chunk = std::max<size_t>(1, static_cast<size_t>(e.batch_limit()) / std::max<size_t>(1, rows));
    OverlappedOut oo(e, std::min(per_chunk, samples) * rows * ctw, samples > per_chunk);
    size_t chunk = 0;
    for (size_t s0 = 0; s0 < samples; s0 += per_chunk) {
      const size_t ns = std::min(per_chunk, samples - s0), items = ns * rows;
      Engine::Scope inner(e);
      u64 *dx = up(e, x + s0 * ctw, ns * ctw);
      u64 *A = e.scratch(items * ctw), *B = oo.buf(chunk), *t3 = e.scratch(items * e.ct_words(3));
      for (size_t s = 0; s < ns; ++s) {
        e.broadcast(dx + s * ctw, A + s * rows * ctw, ctw, rows);
        e.dev().d2d(B + s * rows * ctw, dw, rows * ctw * 8);
      }
      e.multiply(A, B, t3, items);   // sealhelper::packed_enc_multiply
      e.relinearize(t3, A, items);   // Evaluator::relinearize_inplace
      e.vec_sum(A, n, keyset, B, items);  // sealhelper::encrypted_vec_sum
      oo.send(chunk++, out + s0 * rows * ctw, items * ctw);
    }
    oo.finish();
  });
}

// CSP_hhe_pktnn_1fc::evaluateModel (src/examples/CSP/CSP.cpp:288-323): every processed record is multiplied with every
// encrypted weight row, relinearised and summed (packed_enc_multiply -> relinearize_inplace -> encrypted_vec_sum).
int hhe_csp_evaluate_model(hhe_ctx *ctx, const uint64_t *records_ct, size_t records, const uint64_t *enc_weights, size_t rows,
                           size_t input_len, int sum_keyset, uint64_t *out) {
  return hhe_fc_rows(ctx, records_ct, records, enc_weights, rows, input_len, sum_keyset, out);
}

// ---------------------------------------------------------------------------------------------- device-resident API
int hhe_dev_alloc(hhe_ctx *ctx, size_t bytes, void **dptr) {
  return guarded([&] { *dptr = E(ctx).dev().dmalloc(bytes); });
}
int hhe_dev_free(hhe_ctx *ctx, void *dptr) {
  return guarded([&] { E(ctx).dev().dfree(dptr); });
}
int hhe_dev_upload(hhe_ctx *ctx, void *dptr, const void *host, size_t bytes) {
  return guarded([&] {
    E(ctx).dev().h2d(dptr, host, bytes);
    E(ctx).dev().sync();
  });
}
int hhe_dev_download(hhe_ctx *ctx, void *host, const void *dptr, size_t bytes) {
  return guarded([&] {
    E(ctx).dev().d2h(host, dptr, bytes);
    E(ctx).dev().sync();
  });
}
int hhe_sync(hhe_ctx *ctx) {
  return guarded([&] { E(ctx).dev().sync(); });
}

int hhe_dev_ntt(hhe_ctx *ctx, int limb, int inverse, uint64_t *d_data, size_t count) {
  return guarded([&] {
    Engine &e = E(ctx);
    if (limb < 0 || limb >= 2 * e.params().K) throw std::invalid_argument("limb index out of range");
    TabMap m{};
    m.id[0] = static_cast<unsigned char>(limb);
    e.ntt(d_data, d_data, count, 1, m, inverse != 0);
  });
}

int hhe_dev_rotate_rows(hhe_ctx *ctx, const uint64_t *d_a, int steps, int keyset, uint64_t *d_out, size_t count) {
  return guarded([&] { E(ctx).rotate_rows(d_a, steps, keyset, d_out, count); });
}

int hhe_dev_relinearize(hhe_ctx *ctx, const uint64_t *d_a3, uint64_t *d_out, size_t count) {
  return guarded([&] { E(ctx).relinearize(d_a3, d_out, count); });
}

int hhe_dev_multiply(hhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out3, size_t count) {
  return guarded([&] { E(ctx).multiply(d_a, d_b, d_out3, count); });
}

int hhe_dev_pasta3_decompose(hhe_ctx *ctx, const uint64_t *d_enc_key, const uint64_t *d_sym, const uint32_t *lens,
                             const uint64_t *counters, size_t nblocks, uint64_t nonce, int use_bsgs, uint64_t *d_out) {
  return guarded([&] {
    Engine &e = E(ctx);
    Engine::Scope sc(e);
    u32 *d_lens = reinterpret_cast<u32 *>(e.scratch((nblocks + 1) / 2));
    e.dev().h2d(d_lens, lens, nblocks * 4);
    std::vector<u64> ctr(counters, counters + nblocks);
    e.pasta_decompose(d_enc_key, d_sym, d_lens, ctr, nonce, use_bsgs != 0, d_out);
  });
}

int hhe_pasta3_plain(hhe_ctx *ctx, const uint64_t *key256, const uint64_t *in, size_t n_words, uint64_t nonce, uint64_t first_counter,
                     int decrypt, uint64_t *out) {
  return guarded([&] {
    Engine &e = E(ctx);
    if (!key256 || (!in && n_words) || (!out && n_words)) throw std::invalid_argument("null buffer");
    if (!n_words) return;
    Engine::Scope sc(e);
    u64 *dk = up(e, key256, 2 * kPastaT), *din = up(e, in, n_words), *dout = e.scratch(n_words);
    e.pasta_plain(dk, din, n_words, nonce, first_counter, decrypt != 0, dout);
    down(e, out, dout, n_words);
    e.dev().sync();
  });
}

int hhe_pasta_layer_material(hhe_ctx *ctx, uint64_t nonce, uint64_t counter, int layer, uint32_t *mat1, uint32_t *mat2,
                             uint32_t *rc) {
  return guarded([&] {
    Engine &e = E(ctx);
    if (layer < 0 || layer > 3) throw std::invalid_argument("layer must be 0..3");
    Engine::Scope sc(e);
    u64 *d_ctr = up(e, &counter, 1);
    u32 *d_mat = reinterpret_cast<u32 *>(e.scratch((kMaterialWords + 1) / 2));
    e.material(d_ctr, 1, nonce, d_mat);
    const size_t mw = static_cast<size_t>(kPastaT) * kPastaT;
    e.dev().d2h(mat1, d_mat + (static_cast<size_t>(layer) * 2) * mw, mw * 4);
    e.dev().d2h(mat2, d_mat + (static_cast<size_t>(layer) * 2 + 1) * mw, mw * 4);
    e.dev().d2h(rc, d_mat + kMatWords + layer * 2 * kPastaT, 2 * kPastaT * 4);
    e.dev().sync();
  });
}

// ---------------------------------------------------------------------------------------------- client side: encryption
namespace {
// seeds for callers that pass none: 64 bytes per ciphertext from the operating system, as SEAL's random_uint64() (getrandom /
// /dev/urandom) seeds a fresh Blake2xbPRNG for every encryption
void os_random(u64 *out, size_t words) {
  FILE *f = std::fopen("/dev/urandom", "rb");
  if (!f || std::fread(out, 8, words, f) != words) {
    if (f) std::fclose(f);
    throw std::runtime_error("cannot read /dev/urandom for the encryption seeds");
  }
  std::fclose(f);
}

void encrypt_host(Engine &e, const uint64_t *pk, const uint64_t *seeds, const uint64_t *plain, const uint64_t *slots, size_t n_slots,
                  size_t count, uint64_t *out) {
  const Params &p = e.params();
  if (!pk || !out || (!plain && !slots)) throw std::invalid_argument("null buffer");
  if (!count) return;
  const size_t N = p.N, ctw = e.ct_words();
  if (slots) {
    if (n_slots > N) throw std::invalid_argument("values_matrix size exceeds slot count");
    for (size_t i = 0; i < n_slots * count; ++i)
      if (slots[i] >= p.t) throw std::invalid_argument("input value is larger than plain_modulus");
  } else {
    for (size_t i = 0; i < N * count; ++i)
      if (plain[i] >= p.t) throw std::invalid_argument("plain is not valid for encryption parameters");
  }
  for (size_t i = 0; i < static_cast<size_t>(2) * p.K * N; ++i)
    if (pk[i] >= p.q[(i / N) % p.K]) throw std::invalid_argument("public key is not valid for encryption parameters");
  std::vector<u64> own;
  if (!seeds) {
    own.resize(count * 8);
    os_random(own.data(), own.size());
    seeds = own.data();
  }
  Engine::Scope sc(e);
  u64 *d_pk = up(e, pk, static_cast<size_t>(2) * p.K * N);
  const size_t step = static_cast<size_t>(std::max(1, e.batch_limit()));
  OverlappedOut oo(e, std::min(step, count) * ctw, count > step);
  size_t chunk = 0;
  chunked(e, count, [&](size_t off, size_t nb) {
    Engine::Scope inner(e);
    u64 *d_seeds = up(e, seeds + off * 8, nb * 8);
    u64 *d_pt;
    if (slots) {
      u64 *ds = up(e, slots + off * n_slots, std::max<size_t>(1, n_slots * nb));
      d_pt = e.scratch(nb * N);
      e.encode_slots(ds, n_slots, nullptr, static_cast<u32>(n_slots), d_pt, nb);
    } else {
      d_pt = up(e, plain + off * N, nb * N);
    }
    e.encrypt(d_pk, d_seeds, d_pt, nb, oo.buf(chunk));
    oo.send(chunk++, out + off * ctw, nb * ctw);
  });
  oo.finish();
}
}  // namespace

int hhe_encrypt(hhe_ctx *ctx, const uint64_t *pk, const uint64_t *seeds, const uint64_t *plain, size_t count, uint64_t *out) {
  return guarded([&] { encrypt_host(E(ctx), pk, seeds, plain, nullptr, 0, count, out); });
}

int hhe_encrypt_slots(hhe_ctx *ctx, const uint64_t *pk, const uint64_t *seeds, const uint64_t *slots, size_t n_slots, size_t count,
                      uint64_t *out) {
  return guarded([&] { encrypt_host(E(ctx), pk, seeds, nullptr, slots, n_slots, count, out); });
}

// ---------------------------------------------------------------------------------------------- SEAL wire format
namespace {
sealio::Ring ring_of(const hhe_seal_ring *r) {
  if (!r || !r->q || r->nq < 2) throw std::invalid_argument("null or incomplete hhe_seal_ring");
  return sealio::Ring{r->N, r->t, std::vector<uint64_t>(r->q, r->q + r->nq)};
}
sealio::Ring ring_from_params(const Params &p) { return sealio::Ring{p.N, p.t, std::vector<uint64_t>(p.q.begin(), p.q.end())}; }
}  // namespace

int hhe_seal_parms_id(const hhe_seal_ring *ring, int level, uint64_t out[4]) {
  return guarded([&] {
    if (!out || level < 0 || level > 1) throw std::invalid_argument("level must be 0 (data) or 1 (key)");
    sealio::parms_id(ring_of(ring), level, out);
  });
}

size_t hhe_seal_ct_save_bound(const hhe_seal_ring *ring, int size) {
  if (!ring || ring->nq < 2 || size < 1) return 0;
  return sealio::ct_save_bound(sealio::Ring{ring->N, ring->t, std::vector<uint64_t>(static_cast<size_t>(ring->nq), 0)}, size);
}

int hhe_seal_ct_save(const hhe_seal_ring *ring, const uint64_t *ct, int size, int compr_mode, uint8_t *out, size_t cap,
                     size_t *written) {
  return guarded([&] {
    const size_t n = sealio::ct_save(ring_of(ring), ct, size, compr_mode, out, cap);
    if (written) *written = n;
  });
}

int hhe_seal_ct_load(const hhe_seal_ring *ring, const uint8_t *in, size_t len, uint64_t *ct, size_t cap_words, int *size,
                     size_t *consumed) {
  return guarded([&] {
    const size_t n = sealio::ct_load(ring_of(ring), in, len, ct, cap_words, size);
    if (consumed) *consumed = n;
  });
}

int hhe_seal_keys_unpack(const hhe_seal_ring *ring, const uint8_t *in, size_t len, uint64_t *index, uint64_t *ksk, size_t cap_keys,
                         size_t *n_keys, size_t *consumed) {
  return guarded([&] {
    const sealio::Ring r = ring_of(ring);
    const size_t per = static_cast<size_t>(r.L()) * 2 * r.K() * r.N;
    size_t n = 0;
    const size_t used = sealio::keys_walk(r, in, len, [&](uint64_t idx, const uint64_t *k) {
      if (ksk) {
        if (n >= cap_keys) throw std::invalid_argument("key buffer too small");
        std::memcpy(ksk + n * per, k, per * 8);
        if (index) index[n] = idx;
      }
      ++n;
    });
    if (n_keys) *n_keys = n;
    if (consumed) *consumed = used;
  });
}

int hhe_load_seal_keys(hhe_ctx *ctx, int kind, const uint8_t *in, size_t len, size_t *n_keys, size_t *consumed) {
  return guarded([&] {
    Engine &e = E(ctx);
    if (kind < 0 || kind > 2) throw std::invalid_argument("kind must be HHE_KEYSET_0, HHE_KEYSET_1 or HHE_RELIN");
    size_t n = 0;
    const size_t used = sealio::keys_walk(ring_from_params(e.params()), in, len, [&](uint64_t idx, const uint64_t *k) {
      if (kind == 2 && idx != 0) throw std::logic_error("key data is invalid: RelinKeys with more than one key");
      e.load_ksk(kind, kind == 2 ? 0u : static_cast<u32>(2 * idx + 1), k);
      ++n;
    });
    if (n_keys) *n_keys = n;
    if (consumed) *consumed = used;
  });
}

int hhe_pasta3_decompose_serialized(hhe_ctx *ctx, const uint8_t *enc_key_bytes, size_t enc_key_len, const uint64_t *sym_ct,
                                    size_t n_words, uint64_t nonce, uint64_t first_counter, int use_bsgs, int compr_mode, uint8_t *out,
                                    size_t cap, size_t *out_sizes, size_t *written) {
  return guarded([&] {
    Engine &e = E(ctx);
    const sealio::Ring r = ring_from_params(e.params());
    const size_t ctw = e.ct_words();
    std::vector<u64> key(ctw);
    int ksize = 0;
    sealio::ct_load(r, enc_key_bytes, enc_key_len, key.data(), ctw, &ksize);
    if (ksize != 2) throw std::invalid_argument("encrypted symmetric key must be a size-2 ciphertext");
    const size_t nblocks = (n_words + kPastaT - 1) / kPastaT;
    std::vector<u64> cts(nblocks * ctw);
    const int rc = hhe_pasta3_decompose(ctx, key.data(), sym_ct, n_words, nonce, first_counter, use_bsgs, cts.data());
    if (rc == HHE_ERR_INVALID) throw std::invalid_argument(g_error);  // keep the inner call's error class
    if (rc == HHE_ERR_LOGIC) throw std::logic_error(g_error);
    if (rc != HHE_OK) throw std::runtime_error(g_error);
    size_t o = 0;
    for (size_t b = 0; b < nblocks; ++b) {
      const size_t n = sealio::ct_save(r, cts.data() + b * ctw, 2, compr_mode, out + o, cap - o);
      if (out_sizes) out_sizes[b] = n;
      o += n;
    }
    if (written) *written = o;
  });
}

}  // extern "C"
