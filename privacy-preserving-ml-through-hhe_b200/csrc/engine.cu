// Engine implementation: orchestration of the sm_100a kernels in kernels.h / keccak.h.
#include "engine.h"

#include <algorithm>
#include <map>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "keccak.h"

namespace hhe {

namespace {

constexpr int kEwThreads = 256;

inline size_t ew_grid(size_t total) { return (total + kEwThreads - 1) / kEwThreads; }

inline int ntt_threads(int logS) {
  int groups = 1 << (logS - kRadixLog);
  return std::max(32, std::min(HHE_MAX_THREADS, groups));
}

bool getenv_flag(const char *name) {
  const char *v = std::getenv(name);
  return v && *v && *v != '0';
}

TabMap map_mod(int limbs, int period, int base) {
  TabMap m{};
  for (int l = 0; l < limbs && l < kMaxMapLimbs; ++l) m.id[l] = static_cast<unsigned char>(base + (l % period));
  return m;
}

// mod-2N inverse of an odd Galois element
u32 inv_mod_2n(u32 elt, u64 two_n) {
  u64 inv = 1;
  for (int i = 0; i < 6; ++i) inv = (inv * (2 - static_cast<u64>(elt) * inv)) & (two_n - 1);
  return static_cast<u32>(inv);
}

}  // namespace

#define HHE_DISPATCH_LOG(value, ...)                                                       \
  switch (value) {                                                                         \
    case 8: { constexpr int LOGV = 8; __VA_ARGS__; } break;                                \
    case 9: { constexpr int LOGV = 9; __VA_ARGS__; } break;                                \
    case 10: { constexpr int LOGV = 10; __VA_ARGS__; } break;                              \
    case 11: { constexpr int LOGV = 11; __VA_ARGS__; } break;                              \
    case 12: { constexpr int LOGV = 12; __VA_ARGS__; } break;                              \
    case 13: { constexpr int LOGV = 13; __VA_ARGS__; } break;                              \
    case 14: { constexpr int LOGV = 14; __VA_ARGS__; } break;                              \
    default: throw std::invalid_argument("poly_modulus_degree not supported by the shared-memory NTT (512..16384)"); \
  }

Engine::Engine(const Params &p, int device, void *stream) : P_(p), device_(device) {
  if (P_.logn < 9 || P_.logn > 15) throw std::invalid_argument("poly_modulus_degree must be in [512, 32768]");
  // N = 32768: a limb does not fit one SM's shared memory -> split transforms (kernels.h NttSplitBody / KsDigitsQuadBody).
  // HHE_FORCE_SPLIT=1 selects the same code path at smaller N (tests).
  split_ = P_.logn == 15 || getenv_flag("HHE_FORCE_SPLIT");
  if (split_ && table_is_f64(P_, 0)) throw std::invalid_argument("HHE_FORCE_SPLIT needs integer-path moduli (> 2^49)");
#ifdef HHE_CUDA
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("NO_DEVICE: no CUDA device is visible (there is no CPU path)");
  if (device < 0 || device >= ndev) throw std::runtime_error("NO_DEVICE: CUDA device ordinal out of range");
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  cudaDeviceProp prop{};
  cuda_check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  if (prop.major < 10) throw std::runtime_error("NO_DEVICE: libhhe_b200 is built for sm_100a (Blackwell) only");
  dev_.sm_count = prop.multiProcessorCount;
  if (stream) {
    dev_.stream = static_cast<cudaStream_t>(stream);
  } else {
    cuda_check(cudaStreamCreateWithFlags(&dev_.stream, cudaStreamNonBlocking), "cudaStreamCreate");
    dev_.own_stream = true;
  }
  batch_ = 2 * dev_.sm_count;
#else
  (void)stream;
  dev_.sm_count = 1;
  batch_ = 4;
#endif
  const DevConsts hc = make_devconsts(P_);
  dC_ = static_cast<DevConsts *>(dev_.dmalloc(sizeof(DevConsts)));
  dev_.h2d(dC_, &hc, sizeof(DevConsts));
  const size_t ntab = P_.tab.size();
  f64_gmin_ = P_.logn % 3 ? P_.logn % 3 : 3;
  compact_keys_ = true;
  for (int k = 0; k < P_.K; ++k) compact_keys_ = compact_keys_ && table_is_f64(P_, k);
  tmem_ks_ = compact_keys_ && !split_ && !getenv_flag("HHE_NO_TMEM");
  half_fwd_ = compact_keys_ && !split_ && !getenv_flag("HHE_NO_HALF");
  if (const char *v = std::getenv("HHE_KS_THREADS")) ks_threads_ = std::atoi(v);
  cluster_inv_ = half_fwd_ && !getenv_flag("HHE_NO_CLUSTER");
  std::vector<W2> tw(ntab * 2 * P_.N);
  for (size_t t = 0; t < ntab; ++t) {
    if (!P_.tab[t].q) continue;
    const bool f64 = table_is_f64(P_, static_cast<int>(t));
    if (f64) {
      // FP64-pipe tables: two planes of N doubles per direction (index-major, component-major; see ntt_core.h)
      for (int dir = 0; dir < 2; ++dir) {
        double *plane = reinterpret_cast<double *>(&tw[(t * 2 + dir) * P_.N]);
        const std::vector<Twiddle> &src = dir ? P_.tab[t].inv : P_.tab[t].fwd;
        for (u64 k = 0; k < P_.N; ++k) plane[k] = static_cast<double>(src[k].w);
        double *cm = plane + P_.N;
        for (int g0 = f64_gmin_; g0 + 3 <= P_.logn; g0 += 3) {
          double *T = cm + f64tw_offset(g0, f64_gmin_);
          for (int d = 0; d < 3; ++d)
            for (int j = 0; j < (1 << d); ++j)
              for (u64 H = 0; H < (1ULL << g0); ++H)
                T[(static_cast<size_t>((1 << d) - 1 + j) << g0) + H] = static_cast<double>(src[(1ULL << (g0 + d)) + (H << d) + j].w);
        }
      }
      continue;
    }
    for (u64 k = 0; k < P_.N; ++k) {
      tw[(t * 2) * P_.N + k] = w2(P_.tab[t].fwd[k]);
      tw[(t * 2 + 1) * P_.N + k] = w2(P_.tab[t].inv[k]);
    }
  }
  dTw_ = static_cast<W2 *>(dev_.dmalloc(tw.size() * sizeof(W2)));
  dev_.h2d(dTw_, tw.data(), tw.size() * sizeof(W2));
  dIndex_ = static_cast<u32 *>(dev_.dmalloc(P_.N * sizeof(u32)));
  dev_.h2d(dIndex_, P_.index_map.data(), P_.N * sizeof(u32));
  dev_.sync();
}

Engine::~Engine() {
#ifdef HHE_CUDA
  cudaSetDevice(device_);
  cudaStreamSynchronize(dev_.stream);
#endif
  for (auto &kv : keys_) dev_.dfree(kv.second);
  for (auto &kv : perms_) dev_.dfree(kv.second);
  for (auto &c : chunks_) dev_.dfree(c.ptr);
  dev_.dfree(dC_);
  dev_.dfree(dTw_);
  dev_.dfree(dIndex_);
  dev_.dfree(dFeistel_);
#ifdef HHE_CUDA
  if (dev_.own_stream) cudaStreamDestroy(dev_.stream);
#endif
}

// ------------------------------------------------------------------------------------------------ arena
u64 *Engine::scratch(size_t words) {
  words = (words + 31) & ~static_cast<size_t>(31);  // 256-byte granularity
  for (size_t c = cur_chunk_; c < chunks_.size(); ++c) {
    if (chunks_[c].used + words <= chunks_[c].words) {
      u64 *p = chunks_[c].ptr + chunks_[c].used;
      chunks_[c].used += words;
      cur_chunk_ = c;
      return p;
    }
  }
#ifdef HHE_CUDA
  const size_t min_words = (static_cast<size_t>(256) << 20) / 8;
#else
  const size_t min_words = (static_cast<size_t>(1) << 20) / 8;
#endif
  size_t sz = std::max(words, min_words);
  if (!chunks_.empty()) sz = std::max(sz, chunks_.back().words);
  Chunk ch{static_cast<u64 *>(dev_.dmalloc(sz * 8)), sz, words};
  chunks_.push_back(ch);
  cur_chunk_ = chunks_.size() - 1;
  return ch.ptr;
}

void Engine::arena_restore(size_t chunk, size_t used) {
  if (chunks_.empty()) return;
  for (size_t c = chunk + 1; c < chunks_.size(); ++c) chunks_[c].used = 0;
  chunks_[chunk].used = used;
  cur_chunk_ = chunk;
}

// ------------------------------------------------------------------------------------------------ keys
void Engine::load_ksk(int kind, u32 elt, const u64 *host_ksk) {
  if (kind < 0 || kind > 2) throw std::invalid_argument("key kind must be 0, 1 (galois keysets) or 2 (relin)");
  if (kind == 2) elt = 0;
  const size_t words = static_cast<size_t>(P_.L) * 2 * P_.K * P_.N;
  Scope sc(*this);
  u64 *raw = scratch(words);
  dev_.h2d(raw, host_ksk, words * 8);
  W2 *dst = static_cast<W2 *>(dev_.dmalloc(words * (compact_keys_ ? sizeof(double) : sizeof(W2))));
  ShoupifyBody body{raw, dst, dC_, words, compact_keys_ ? 1 : 0, tmem_ks_ ? 1 : 0};
  dev_.launch(body, ew_grid(words), kEwThreads, 0);
  dev_.sync();  // host_ksk may be released by the caller; raw scratch is recycled
  auto key = std::make_pair(kind, elt);
  auto it = keys_.find(key);
  if (it != keys_.end()) dev_.dfree(it->second);
  keys_[key] = dst;
}

void Engine::clear_keyset(int kind) {
  if (kind < 0 || kind > 2) throw std::invalid_argument("key kind must be 0, 1 (galois keysets) or 2 (relin)");
  dev_.sync();  // no kernel may still read the keys
  for (auto it = keys_.begin(); it != keys_.end();) {
    if (it->first.first == kind) {
      dev_.dfree(it->second);
      it = keys_.erase(it);
    } else {
      ++it;
    }
  }
}

const W2 *Engine::find_key(int kind, u32 elt) const {
  auto it = keys_.find(std::make_pair(kind, kind == 2 ? 0u : elt));
  return it == keys_.end() ? nullptr : it->second;
}

const W2 *Engine::need_key(int kind, u32 elt) const {
  const W2 *k = find_key(kind, elt);
  if (!k) throw std::invalid_argument(kind == 2 ? "relinearization key not loaded" : "Galois key not present");
  return k;
}

// ------------------------------------------------------------------------------------------------ primitives
void Engine::ntt(const u64 *in, u64 *out, size_t items, int limbs, const TabMap &map, bool inverse, size_t item_stride,
                 size_t limb_stride) {
  if (split_) {
    if (limb_stride) throw std::invalid_argument("strided limbs are not supported by the split transforms");
    const size_t stride = item_stride ? item_stride : static_cast<size_t>(limbs) * P_.N;
    Scope sc(*this);
    if (!inverse && in == out) {
      // both half-CTAs of a limb read the whole limb: an in-place forward transform needs a private copy of the input
      const size_t words = (items - 1) * stride + static_cast<size_t>(limbs) * P_.N;
      u64 *copy = scratch(words);
      dev_.d2d(copy, in, words * 8);
      in = copy;
    }
    HHE_DISPATCH_LOG(P_.logn - 1, {
      NttSplitBody<LOGV> body{in, out, dC_, twref(), map, limbs, inverse ? 1 : 0, stride};
      dev_.launch(body, items * limbs * 2, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
    });
    if (inverse) {
      const size_t total = items * limbs * (P_.N / 2);
      InvFinalBody fin{out, dC_, twref(), map, limbs, stride, total};
      dev_.launch(fin, ew_grid(total), kEwThreads, 0);
    }
    return;
  }
  if (cluster_inv_) {
    bool all_f64 = true;
    for (int l = 0; l < limbs; ++l) all_f64 = all_f64 && table_is_f64(P_, map.id[l]);
    if (all_f64 && !inverse && !getenv_flag("HHE_NO_FWD_CLUSTER")) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        NttFwdClusterBody<LOGV> body{in, out, dC_, twref(), map, limbs, item_stride ? item_stride : static_cast<size_t>(limbs) << (LOGV + 1),
                                     limb_stride ? limb_stride : static_cast<size_t>(2) << LOGV};
        dev_.launch_cluster2(body, items * limbs * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
      return;
    }
    if (all_f64 && inverse) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        using Body = InvClusterBody<LOGV, PlanScaled>;
        Body body{PlanScaled{in, out, map, limbs, item_stride ? item_stride : static_cast<size_t>(limbs) << (LOGV + 1),
                             limb_stride ? limb_stride : static_cast<size_t>(2) << LOGV},
                  dC_, twref()};
        dev_.launch_cluster2(body, items * limbs * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
      return;
    }
  }
  HHE_DISPATCH_LOG(P_.logn, {
    NttBody<LOGV> body{in, out, dC_, twref(), map, limbs, inverse ? 1 : 0, item_stride ? item_stride : static_cast<size_t>(limbs) << LOGV,
                        limb_stride ? limb_stride : static_cast<size_t>(1) << LOGV};
    dev_.launch(body, items * limbs, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::require_whole_limb(const char *what) const {
  if (split_) throw std::invalid_argument(std::string(what) + " is not available at poly_modulus_degree 32768 in this build (NTT, rotate, relinearize, multiply are)");
}

void Engine::add(const u64 *a, const u64 *b, u64 *out, size_t items, int size) {
  const size_t total = items * ct_words(size);
  AddBody body{a, b, out, dC_, size * P_.L, total};
  dev_.launch(body, ew_grid(total), kEwThreads, 0);
}

void Engine::negate(const u64 *a, u64 *out, size_t items) {
  const size_t total = items * ct_words();
  NegateBody body{a, out, dC_, total};
  dev_.launch(body, ew_grid(total), kEwThreads, 0);
}

void Engine::add_plain(const u64 *a, const u64 *pt, size_t pstride, u64 *out, size_t items, bool negate_first, const u32 *ptidx) {
  const size_t total = items * ct_words();
  AddPlainBody body{a, pt, pstride, out, dC_, negate_first ? 1 : 0, total, ptidx};
  dev_.launch(body, ew_grid(total), kEwThreads, 0);
}

void Engine::broadcast(const u64 *src, u64 *out, size_t words, size_t items) {
  BroadcastBody body{src, out, words, words * items};
  dev_.launch(body, ew_grid(words * items), kEwThreads, 0);
}

void Engine::encode_slots(const u64 *slots, size_t sstride, const u32 *lens, u32 n, u64 *pt, size_t items) {
  if (n > P_.N) throw std::invalid_argument("values_matrix size exceeds slot count");
  require_whole_limb("encode");
  HHE_DISPATCH_LOG(P_.logn, {
    EncodeBody<LOGV> body{slots, sstride, lens, n, nullptr, nullptr, dIndex_, pt, dC_, twref(), kSlots, 0, 0};
    dev_.launch(body, items, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::encode_material(const u32 *material, const u32 *mat_index, int mode, int layer, int diag, u64 *pt, size_t items) {
  require_whole_limb("PASTA transciphering");
  HHE_DISPATCH_LOG(P_.logn, {
    EncodeBody<LOGV> body{nullptr, 0, nullptr, 0, material, mat_index, dIndex_, pt, dC_, twref(), mode, layer, diag};
    dev_.launch(body, items, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::lift_ntt(const u64 *pt, u64 *D, size_t items) {
  require_whole_limb("multiply_plain");
  if (half_fwd_) {
    HHE_DISPATCH_LOG(P_.logn - 1, {
      LiftNttHalfBody<LOGV> body{pt, D, dC_, twref()};
      dev_.launch(body, items * P_.L * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
    });
    return;
  }
  HHE_DISPATCH_LOG(P_.logn, {
    LiftNttBody<LOGV> body{pt, D, dC_, twref()};
    dev_.launch(body, items * P_.L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::ntt_mac(const u64 *ct, const u64 *D, size_t dstride, u64 *sum, size_t items, bool first, int comps, size_t sum_off,
                     u64 *ntt_out, const u32 *didx) {
  require_whole_limb("multiply_plain");
  if (half_fwd_) {
    HHE_DISPATCH_LOG(P_.logn - 1, {
      NttMacHalfBody<LOGV> body{ct, D, dstride, sum, dC_, twref(), first ? 1 : 0, comps, ct_words(), sum_off, ntt_out, didx};
      dev_.launch(body, items * comps * P_.L * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
    });
    return;
  }
  HHE_DISPATCH_LOG(P_.logn, {
    NttMacBody<LOGV> body{ct, D, dstride, sum, dC_, twref(), first ? 1 : 0, comps, ct_words(), sum_off, ntt_out, didx};
    dev_.launch(body, items * comps * P_.L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::strided_copy(const u64 *src, size_t sstride, u64 *dst, size_t dstride, size_t words, size_t rows) {
  StridedCopyBody body{src, dst, sstride, dstride, words, words * rows};
  dev_.launch(body, ew_grid(words * rows), kEwThreads, 0);
}

// Permutation of NTT slots induced by X -> X^elt: slot i holds a(psi^(2*bitrev(i)+1)), so galois(a) at slot i is a at
// the slot whose exponent is (2*bitrev(i)+1)*elt mod 2N  (cf. GaloisTool::apply_galois_ntt, seal/util/galois.h:78).
const u32 *Engine::ntt_perm(u32 elt) {
  auto it = perms_.find(elt);
  if (it != perms_.end()) return it->second;
  const u64 N = P_.N, m = 2 * N;
  auto brev = [&](u64 x) {
    u64 r = 0;
    for (int i = 0; i < P_.logn; ++i, x >>= 1) r = (r << 1) | (x & 1);
    return r;
  };
  std::vector<u32> table(N);
  for (u64 i = 0; i < N; ++i) {
    const u64 e = ((2 * brev(i) + 1) * elt) % m;
    table[i] = static_cast<u32>(brev((e - 1) >> 1));
  }
  u32 *d = static_cast<u32 *>(dev_.dmalloc(N * sizeof(u32)));
  dev_.h2d(d, table.data(), N * sizeof(u32));
  dev_.sync();
  perms_[elt] = d;
  return d;
}

void Engine::ct_intt(u64 *ct, size_t items, int size) {
  ntt(ct, ct, items, size * P_.L, map_mod(size * P_.L, P_.L, 0), true);
}

void Engine::multiply_plain(const u64 *a, const u64 *pt, size_t pstride, u64 *out, size_t items) {
  Scope sc(*this);
  const size_t ditems = pstride ? items : 1;
  u64 *D = scratch(ditems * P_.L * P_.N);
  lift_ntt(pt, D, ditems);
  ntt_mac(a, D, pstride ? static_cast<size_t>(P_.L) * P_.N : 0, out, items, true);
  ct_intt(out, items);
}

void Engine::galois(const u64 *a, u32 elt, u64 *out, size_t items) {
  const size_t total = items * ct_words();
  GaloisBody body{a, out, dC_, inv_mod_2n(elt, 2 * P_.N), total};
  dev_.launch(body, ew_grid(total), kEwThreads, 0);
}

void Engine::key_switch(const u64 *target, size_t tstride, const W2 *key, const u64 *base0, const u64 *base1,
                        size_t bstride, u64 *out, size_t items) {
  Scope sc(*this);
  const int K = P_.K;
  u64 *acc = scratch(items * 2 * K * P_.N);
  if (split_) {
    HHE_DISPATCH_LOG(P_.logn - 2, {
      constexpr int S = 1 << LOGV;
      KsDigitsQuadBody<LOGV> body{target, tstride, key, acc, dC_, twref(), static_cast<int>(items)};
      dev_.launch(body, items * K * 4, ntt_threads(LOGV), (ntt_smem_words(S) + 2 * S) * 8);
    });
  } else {
    launch_ks_digits(target, tstride, key, acc, items, nullptr, 0, nullptr);
  }
  if (compact_keys_ && !split_) {
    // FP64 path: the two special limbs first, then the data limbs with ModDown + add fused into the transform's store
    TabMap msp2{};
    msp2.id[0] = msp2.id[1] = static_cast<unsigned char>(K - 1);
    const size_t N = P_.N;
    ntt(acc + static_cast<size_t>(K - 1) * N, acc + static_cast<size_t>(K - 1) * N, items, 2, msp2, true, static_cast<size_t>(2) * K * N,
        static_cast<size_t>(K) * N);
    if (cluster_inv_) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        using Body = InvClusterBody<LOGV, PlanModDownAdd>;
        Body body{PlanModDownAdd{acc, base0, base1, bstride, out}, dC_, twref()};
        dev_.launch_cluster2(body, items * 2 * P_.L * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
      return;
    }
    HHE_DISPATCH_LOG(P_.logn, {
      InttModDownAddBody<LOGV> body{acc, base0, base1, bstride, out, dC_, twref()};
      dev_.launch(body, items * 2 * P_.L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
    });
    return;
  }
  ntt(acc, acc, items, 2 * K, map_mod(2 * K, K, 0), true);
  ModDownBody md{acc, base0, base1, bstride, out, dC_, items * P_.N};
  dev_.launch(md, ew_grid(items * P_.N), kEwThreads, 0);
}

void Engine::launch_ks_digits(const u64 *target, size_t tstride, const W2 *key, u64 *acc, size_t items, const u64 *reuse,
                              size_t reuse_stride, const u32 *perm) {
  const int K = P_.K;
  if (tmem_ks_) {
#ifdef HHE_CUDA
    const bool emulate = false;
#else
    const bool emulate = true;
#endif
    HHE_DISPATCH_LOG(P_.logn - 1, {
      constexpr int G = (1 << LOGV) / 8;
      if (ks_threads_ == 256) {
        const int nt = std::max(32, std::min(256, G));
        KsDigitsTmemBody<LOGV, 256> body{target, tstride, reinterpret_cast<const double *>(key), acc, dC_, twref(), static_cast<int>(items),
                                         reuse, reuse_stride, perm};
        dev_.launch(body, items * K * 2, nt, KsDigitsTmemBody<LOGV, 256>::smem_bytes(nt, emulate));
      } else {
        const int nt = std::max(32, std::min(512, G));
        KsDigitsTmemBody<LOGV> body{target, tstride, reinterpret_cast<const double *>(key), acc, dC_, twref(), static_cast<int>(items),
                                    reuse, reuse_stride, perm};
        dev_.launch(body, items * K * 2, nt, KsDigitsTmemBody<LOGV>::smem_bytes(nt, emulate));
      }
    });
    return;
  }
  HHE_DISPATCH_LOG(P_.logn - 1, {
    constexpr int S = 1 << LOGV;
    KsDigitsBody<LOGV> body{target, tstride, key, acc, dC_, twref(), static_cast<int>(items), reuse, reuse_stride, perm};
    dev_.launch(body, items * K * 2, ntt_threads(LOGV), (ntt_smem_words(S) + 2 * S) * 8);
  });
}

void Engine::apply_galois(const u64 *a, u32 elt, const W2 *key, u64 *out, size_t items) {
  Scope sc(*this);
  u64 *g = scratch(items * ct_words());
  galois(a, elt, g, items);
  key_switch(g + static_cast<size_t>(P_.L) * P_.N, ct_words(), key, g, nullptr, ct_words(), out, items);
}

void Engine::rotate_rows(const u64 *a, int steps, int keyset, u64 *out, size_t items) {
  if (keyset < 0 || keyset > 1) throw std::invalid_argument("keyset must be 0 or 1");
  if (steps == 0) {
    if (out != a) dev_.d2d(out, a, items * ct_words() * 8);
    return;
  }
  const u32 elt = P_.galois_elt_from_step(steps);
  if (!elt) throw std::invalid_argument("step count too large");
  if (const W2 *key = find_key(keyset, elt)) {
    apply_galois(a, elt, key, out, items);
    return;
  }
  // Evaluator::rotate_internal: fall back to the NAF terms, least-significant first
  std::vector<int> terms = naf_steps(steps);
  if (terms.size() == 1) throw std::invalid_argument("Galois key not present");
  Scope sc(*this);
  u64 *tmp = scratch(items * ct_words());
  const u64 *cur = a;
  // ping-pong between out and tmp so the last term lands in out
  std::vector<int> eff;
  for (int s : terms)
    if (static_cast<u64>(s < 0 ? -s : s) != P_.N / 2) eff.push_back(s);
  if (eff.empty()) {
    if (out != a) dev_.d2d(out, a, items * ct_words() * 8);
    return;
  }
  u64 *bufs[2] = {out, tmp};
  int which = (eff.size() & 1) ? 0 : 1;
  // in-place use (a == out) is safe: apply_galois gathers its whole input into scratch before anything is written
  for (size_t i = 0; i < eff.size(); ++i) {
    const u32 e = P_.galois_elt_from_step(eff[i]);
    const W2 *key = e ? find_key(keyset, e) : nullptr;
    if (!key) throw std::invalid_argument("Galois key not present");
    apply_galois(cur, e, key, bufs[which], items);
    cur = bufs[which];
    which ^= 1;
  }
}

void Engine::rotate_columns(const u64 *a, int keyset, u64 *out, size_t items) {
  const u32 elt = static_cast<u32>(2 * P_.N - 1);
  apply_galois(a, elt, need_key(keyset, elt), out, items);
}

void Engine::relinearize(const u64 *a3, u64 *out, size_t items) {
  const W2 *key = need_key(2, 0);
  const size_t poly = static_cast<size_t>(P_.L) * P_.N;
  if (out == a3) throw std::invalid_argument("relinearize: output must not alias the size-3 input");
  key_switch(a3 + 2 * poly, 3 * poly, key, a3, a3 + poly, 3 * poly, out, items);
}

void Engine::multiply(const u64 *a, const u64 *b, u64 *out3, size_t items) {
  Scope sc(*this);
  const int L = P_.L, K = P_.K, Lb = L + 1;
  const size_t N = P_.N;
  const bool sq = (a == b);
  u64 *aq = scratch(items * 2 * L * N), *ab = scratch(items * 2 * Lb * N);
  u64 *bq = sq ? aq : scratch(items * 2 * L * N), *bb = sq ? ab : scratch(items * 2 * Lb * N);
  u64 *dq = scratch(items * 3 * L * N), *db = scratch(items * 3 * Lb * N);
  const TabMap mq = map_mod(3 * L, L, 0), mb = map_mod(3 * Lb, Lb, K);
  for (int op = 0; op < (sq ? 1 : 2); ++op) {
    const u64 *x = op ? b : a;
    u64 *xq = op ? bq : aq, *xb = op ? bb : ab;
    BehzExtendBody ext{x, xb, dC_, items * 2 * N};
    dev_.launch(ext, ew_grid(items * 2 * N), kEwThreads, 0);
    ntt(x, xq, items, 2 * L, mq, false);
    ntt(xb, xb, items, 2 * Lb, mb, false);
  }
  TensorBody tq{aq, bq, dq, dC_, L, 0, items * L * N};
  dev_.launch(tq, ew_grid(items * L * N), kEwThreads, 0);
  TensorBody tb{ab, bb, db, dC_, Lb, K, items * Lb * N};
  dev_.launch(tb, ew_grid(items * Lb * N), kEwThreads, 0);
  ntt(dq, dq, items, 3 * L, mq, true);
  ntt(db, db, items, 3 * Lb, mb, true);
  BehzScaleRoundBody sr{dq, db, out3, dC_, items * 3 * N};
  dev_.launch(sr, ew_grid(items * 3 * N), kEwThreads, 0);
}

void Engine::exponentiate3(const u64 *a, u64 *out, size_t items) {
  Scope sc(*this);
  u64 *t3 = scratch(items * ct_words(3)), *sq = scratch(items * ct_words());
  multiply(a, a, t3, items);
  relinearize(t3, sq, items);
  multiply(sq, a, t3, items);
  relinearize(t3, out, items);
}

// ------------------------------------------------------------------------------------------------ PASTA-3
void Engine::material(const u64 *d_counters, size_t nblocks, u64 nonce, u32 *d_out) {
  MaterialBody body{d_counters, nonce, d_out, P_.t};
  dev_.launch(body, nblocks, 256, kMaterialSmem);
}

const u64 *Engine::feistel_mask_ntt() {
  if (!dFeistel_) {
    Scope sc(*this);
    u64 *pt = scratch(P_.N);
    encode_material(nullptr, nullptr, kFeistel, 0, 0, pt, 1);
    dFeistel_ = static_cast<u64 *>(dev_.dmalloc(static_cast<size_t>(P_.L) * P_.N * 8));
    lift_ntt(pt, dFeistel_, 1);
  }
  return dFeistel_;
}

// PASTA_SEAL::diagonal (src/pasta/pasta_3_seal.cpp:370-413); the 128 products are summed in the NTT domain.
void Engine::affine_diagonal(u64 *state, const u32 *mat, int layer, size_t nb, size_t nd, const u32 *didx) {
  Scope sc(*this);
  const size_t ctw = ct_words(), N = P_.N, dw = static_cast<size_t>(P_.L) * N;
  const size_t ds = dw;  // diagonals: one set per distinct counter (nd of them), blocks find theirs through didx
  u64 *tmp = scratch(nb * ctw), *sum = scratch(nb * ctw), *pt = scratch(nd * N), *D = scratch(nd * dw);
  if (N / 2 != kPastaT) {
    rotate_rows(state, kPastaT, 0, tmp, nb);
    add(state, tmp, state, nb);
  }
  const u32 e1 = P_.galois_elt_from_step(-1);
  const W2 *k1 = need_key(0, e1);
  u64 *cur = state, *nxt = tmp;
  for (int i = 0; i < kPastaT; ++i) {
    if (i) {
      apply_galois(cur, e1, k1, nxt, nb);
      std::swap(cur, nxt);
    }
    encode_material(mat, nullptr, kDiag, layer, i, pt, nd);
    lift_ntt(pt, D, nd);
    ntt_mac(cur, D, ds, sum, nb, i == 0, 2, 0, nullptr, didx);
  }
  ntt(sum, state, nb, 2 * P_.L, map_mod(2 * P_.L, P_.L, 0), true);
}

// Same computation with the rotating state kept NTT-resident (see kernels.h "NTT-resident rotation chain"): per
// rotation 64 + 10 + 8 + 8 limb transforms instead of 72 + 18 + 16. Used when every coefficient prime is on the FP64
// path; bit-identical to affine_diagonal (tests/test_engine_parity.py compares both rings with the oracle).
void Engine::affine_diagonal_resident(u64 *state, const u32 *mat, int layer, size_t nb, size_t nd, const u32 *didx) {
  Scope sc(*this);
  const int L = P_.L, K = P_.K;
  const size_t ctw = ct_words(), N = P_.N, dw = static_cast<size_t>(L) * N;
  const size_t ds = dw;
  u64 *tmp = scratch(nb * ctw), *sum = scratch(nb * ctw), *pt = scratch(nd * N), *D = scratch(nd * dw);
  u64 *stn = scratch(nb * ctw), *c0a = scratch(nb * dw), *c0b = scratch(nb * dw), *c1c = scratch(nb * dw), *c1n = scratch(nb * dw),
      *g1 = scratch(nb * dw), *acc = scratch(nb * 2 * K * N);
  if (N / 2 != kPastaT) {
    rotate_rows(state, kPastaT, 0, tmp, nb);
    add(state, tmp, state, nb);
  }
  const u32 e1 = P_.galois_elt_from_step(-1);
  const W2 *k1 = need_key(0, e1);
  const u32 *perm = ntt_perm(e1);
  const u32 e1_inv = inv_mod_2n(e1, 2 * N);
  // step 0: sum = NTT(state) * D_0, keeping NTT(state)
  encode_material(mat, nullptr, kDiag, layer, 0, pt, nd);
  lift_ntt(pt, D, nd);
  ntt_mac(state, D, ds, sum, nb, true, 2, 0, stn, didx);
  strided_copy(stn, ctw, c0a, dw, dw, nb);
  strided_copy(stn + dw, ctw, c1n, dw, dw, nb);
  strided_copy(state + dw, ctw, c1c, dw, dw, nb);
  u64 *c0_in = c0a, *c0_out = c0b;
  TabMap msp2{};
  msp2.id[0] = msp2.id[1] = static_cast<unsigned char>(K - 1);
  {  // g1 = galois(c1) in coefficient form: the digits of the first key switch (later ones come out of intt_moddown)
    GaloisBody gb{c1c, g1, dC_, e1_inv, nb * dw};
    dev_.launch(gb, ew_grid(nb * dw), kEwThreads, 0);
  }
  for (int i = 1; i < kPastaT; ++i) {
    launch_ks_digits(g1, dw, k1, acc, nb, c1n, dw, perm);
    // inverse NTT of the two special limbs acc[0][K-1], acc[1][K-1] (K*N words apart inside an item), then of acc[1][i<L]
    // with the ModDown and the next rotation's Galois map fused into the store
    ntt(acc + static_cast<size_t>(K - 1) * N, acc + static_cast<size_t>(K - 1) * N, nb, 2, msp2, true, static_cast<size_t>(2) * K * N,
        static_cast<size_t>(K) * N);
    if (cluster_inv_) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        using Body = InvClusterBody<LOGV, PlanModDownGalois>;
        Body body{PlanModDownGalois{acc, c1c, g1, e1, P_.logn}, dC_, twref()};
        dev_.launch_cluster2(body, nb * L * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
    } else {
      HHE_DISPATCH_LOG(P_.logn, {
        InttModDownBody<LOGV> body{acc, c1c, g1, dC_, twref(), e1};
        dev_.launch(body, nb * L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
    }
    encode_material(mat, nullptr, kDiag, layer, i, pt, nd);
    lift_ntt(pt, D, nd);
    if (half_fwd_) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        Corr0MacHalfBody<LOGV> body{acc, c0_in, c0_out, perm, D, sum, dC_, twref(), ds, didx};
        dev_.launch(body, nb * L * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
    } else {
      HHE_DISPATCH_LOG(P_.logn, {
        Corr0MacBody<LOGV> body{acc, c0_in, c0_out, perm, D, sum, dC_, twref(), ds, didx};
        dev_.launch(body, nb * L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
    }
    std::swap(c0_in, c0_out);
    ntt_mac(c1c, D, ds, sum, nb, false, 1, dw, c1n, didx);
  }
  ntt(sum, state, nb, 2 * L, map_mod(2 * L, L, 0), true);
}

// PASTA_SEAL::babystep_giantstep (src/pasta/pasta_3_seal.cpp:267-366), N1 = 16, N2 = 8
void Engine::affine_bsgs(u64 *state, const u32 *mat, int layer, size_t nb, size_t nd, const u32 *didx) {
  constexpr int N1 = 16, N2 = 8;
  Scope sc(*this);
  const size_t ctw = ct_words(), N = P_.N, dw = static_cast<size_t>(P_.L) * N;
  u64 *tmp = scratch(nb * ctw), *inner = scratch(nb * ctw), *outer = scratch(nb * ctw), *pt = scratch(nd * N1 * N),
      *D = scratch(nd * N1 * dw), *rot = scratch(nb * ctw * N1);
  if (N / 2 != kPastaT) {
    rotate_rows(state, kPastaT, 0, tmp, nb);
    add(state, tmp, state, nb);
  }
  dev_.d2d(rot, state, nb * ctw * 8);
  for (int j = 1; j < N1; ++j) rotate_rows(rot + (j - 1) * nb * ctw, -1, 0, rot + j * nb * ctw, nb);
  const TabMap mq = map_mod(2 * P_.L, P_.L, 0);
  // every baby rotation is multiplied with 8 diagonals: transform each once (in place), then the products are element-wise
  ntt(rot, rot, nb * N1, 2 * P_.L, mq, false);
  for (int k = 0; k < N2; ++k) {
    // the 16 diagonals of this giant step are encoded, lifted and transformed as one batch ([j][nd] items), then one pass over
    // the baby rotations forms the inner sum: every residue of `inner` is written once
    for (int j = 0; j < N1; ++j) encode_material(mat, nullptr, kDiagBsgs, layer, k * N1 + j, pt + static_cast<size_t>(j) * nd * N, nd);
    lift_ntt(pt, D, nd * N1);
    DyadicMacNBody mac{rot, D, inner, dC_, N1, nb * ctw, nd * dw, dw, didx, nb * ctw};
    dev_.launch(mac, ew_grid(nb * ctw), kEwThreads, 0);
    if (k == 0) {
      ntt(inner, outer, nb, 2 * P_.L, mq, true);
    } else {
      ntt(inner, inner, nb, 2 * P_.L, mq, true);
      rotate_rows(inner, -k * N1, 0, tmp, nb);
      add(outer, tmp, outer, nb);
    }
  }
  dev_.d2d(state, outer, nb * ctw * 8);
}

// PASTA_SEAL::sbox_feistel (src/pasta/pasta_3_seal.cpp:222-247)
void Engine::feistel(u64 *state, size_t nb) {
  Scope sc(*this);
  const size_t ctw = ct_words();
  u64 *rot = scratch(nb * ctw), *masked = scratch(nb * ctw), *t3 = scratch(nb * ct_words(3));
  rotate_rows(state, -1, 0, rot, nb);
  ntt_mac(rot, feistel_mask_ntt(), 0, masked, nb, true);
  ct_intt(masked, nb);
  multiply(masked, masked, t3, nb);
  relinearize(t3, rot, nb);
  add(state, rot, state, nb);
}

void Engine::pasta_batch(const u64 *d_enc_key, const u64 *d_sym, const u32 *d_lens, const u64 *d_counters, size_t nb, size_t nd,
                         const u32 *didx, u64 nonce, bool use_bsgs, u64 *d_out) {
  Scope sc(*this);
  const size_t ctw = ct_words(), N = P_.N;
  u64 *state = scratch(nb * ctw), *tmp = scratch(nb * ctw), *pt = scratch(nb * N);
  u32 *mat = reinterpret_cast<u32 *>(scratch((nd * kMaterialWords + 1) / 2));
  feistel_mask_ntt();
  material(d_counters, nd, nonce, mat);
  broadcast(d_enc_key, state, ctw, nb);
  for (int layer = 0; layer < 4; ++layer) {
    if (use_bsgs)
      affine_bsgs(state, mat, layer, nb, nd, didx);
    else if (compact_keys_ && !getenv_flag("HHE_NO_RESIDENT"))
      affine_diagonal_resident(state, mat, layer, nb, nd, didx);
    else
      affine_diagonal(state, mat, layer, nb, nd, didx);
    encode_material(mat, nullptr, kRc, layer, 0, pt, nd);  // add_rc (:205-211)
    add_plain(state, pt, N, state, nb, false, didx);
    rotate_columns(state, 0, tmp, nb);  // mix (:417-423)
    add(tmp, state, tmp, nb);
    add(state, tmp, state, nb);
    if (layer < 2) {
      feistel(state, nb);
    } else if (layer == 2) {
      exponentiate3(state, tmp, nb);  // sbox_cube (:215-218)
      std::swap(state, tmp);
    }
  }
  encode_slots(d_sym, kPastaT, d_lens, kPastaT, pt, nb);
  add_plain(state, pt, N, d_out, nb, true);  // negate_inplace; add_plain (:168-169)
}

void Engine::pasta_decompose(const u64 *d_enc_key, const u64 *d_sym, const u32 *d_lens, const std::vector<u64> &counters,
                             u64 nonce, bool use_bsgs, u64 *d_out) {
  const size_t nblocks = counters.size();
  if (2 * kPastaT != P_.N && 4 * kPastaT > P_.N) throw std::runtime_error("too little slots for matmul implementation!");
  const u32 e1 = P_.galois_elt_from_step(-1), ec = static_cast<u32>(2 * P_.N - 1);
  need_key(0, e1);
  need_key(0, ec);
  need_key(2, 0);
  if (P_.N / 2 != kPastaT) need_key(0, P_.galois_elt_from_step(kPastaT));
  Scope sc(*this);
  const size_t step = static_cast<size_t>(std::max(1, batch_)), ctw = ct_words();
  u64 *d_ctr = scratch(std::min(step, nblocks));
  // Blocks with equal SHAKE counters (records restart at counter 0: CSP.cpp:247-252) have identical round
  // matrices and constants: per lock-step batch the round material, the encoded diagonals and their lifted transforms are computed
  // once per DISTINCT counter, and every block reads its counter's copy through an index (didx).
  const bool share = !getenv_flag("HHE_NO_SHARED_MATERIAL");
  u32 *d_idx = reinterpret_cast<u32 *>(scratch((std::min(step, nblocks) + 1) / 2));
  std::vector<u64> uniq;
  std::vector<u32> idx;
  for (size_t off = 0; off < nblocks; off += step) {
    const size_t nb = std::min(step, nblocks - off);
    uniq.clear();
    idx.resize(nb);
    std::map<u64, u32> seen;
    for (size_t b = 0; b < nb; ++b) {
      const u64 c = counters[off + b];
      auto it = share ? seen.find(c) : seen.end();
      if (it == seen.end()) {
        idx[b] = static_cast<u32>(uniq.size());
        if (share) seen.emplace(c, idx[b]);
        uniq.push_back(c);
      } else {
        idx[b] = it->second;
      }
    }
    const size_t nd = uniq.size();
    if (nd == nb) {  // all distinct: the caller's array is the list of counters
      dev_.h2d(d_ctr, counters.data() + off, nb * 8);
    } else {
      dev_.h2d(d_ctr, uniq.data(), nd * 8);
      dev_.h2d(d_idx, idx.data(), nb * 4);
      dev_.sync();  // uniq / idx are reused by the next batch
    }
    pasta_batch(d_enc_key, d_sym + off * kPastaT, d_lens + off, d_ctr, nb, nd, nd < nb ? d_idx : nullptr, nonce, use_bsgs,
                d_out + off * ctw);
  }
}

// pasta::PASTA::encrypt / decrypt (src/pasta/pasta_3_plain.cpp:9-47): SHAKE material per block, then the keyed permutation
void Engine::pasta_plain(const u64 *d_key256, const u64 *d_in, size_t n_words, u64 nonce, u64 first_counter, bool decrypt, u64 *d_out) {
  const size_t nblocks = (n_words + kPastaT - 1) / kPastaT;
  const size_t step = 1024;  // 512 KiB of round material per block in flight
  Scope sc(*this);
  u64 *d_ctr = scratch(std::min(step, nblocks));
  u32 *mat = reinterpret_cast<u32 *>(scratch((std::min(step, nblocks) * kMaterialWords + 1) / 2));
  std::vector<u64> ctr(std::min(step, nblocks));
  for (size_t off = 0; off < nblocks; off += step) {
    const size_t nb = std::min(step, nblocks - off);
    for (size_t b = 0; b < nb; ++b) ctr[b] = first_counter + off + b;
    dev_.h2d(d_ctr, ctr.data(), nb * 8);
    dev_.sync();  // ctr is reused by the next chunk
    material(d_ctr, nb, nonce, mat);
    PastaPlainBody body{mat, d_key256, d_in, d_out, off * kPastaT, n_words, P_.t, decrypt ? 1 : 0};
    dev_.launch(body, nb, 256, kPastaPlainSmem);
  }
}

// SEALZpCipher::mask (src/pasta/SEAL_Cipher.cpp:161-166)
void Engine::mask(const u64 *a, const u64 *d_mask_slots, u32 n, u64 *out, size_t items) {
  Scope sc(*this);
  u64 *pt = scratch(P_.N);
  encode_slots(d_mask_slots, 0, nullptr, n, pt, 1);
  multiply_plain(a, pt, 0, out, items);
}

// SEALZpCipher::flatten (src/pasta/SEAL_Cipher.cpp:170-181): out = in[0] + sum_i rot(in[i], -128 i)
void Engine::flatten(const u64 *in, size_t per, int keyset, u64 *out, size_t items) {
  Scope sc(*this);
  const size_t ctw = ct_words();
  u64 *gath = scratch(items * ctw), *rot = scratch(items * ctw);
  for (size_t i = 0; i < per; ++i) {
    for (size_t g = 0; g < items; ++g) dev_.d2d((i ? gath : out) + g * ctw, in + (g * per + i) * ctw, ctw * 8);
    if (i) {
      rotate_rows(gath, -static_cast<int>(i * kPastaT), keyset, rot, items);
      add(out, rot, out, items);
    }
  }
}

// sealhelper::encrypted_vec_sum (src/util/sealhelper.cpp:379-392): out = sum_{i<n} rot(a, -i), every rotation from
// the input. Rotations that SEAL expands into NAF chains share their leading steps; those prefixes are computed once
// (bit-exact: the same sequence of key switches is applied to the same operand).
void Engine::vec_sum(const u64 *a, size_t n, int keyset, u64 *out, size_t items) {
  if (keyset < 0 || keyset > 1) throw std::invalid_argument("keyset must be 0 or 1");
  std::vector<std::vector<int>> seqs;
  for (size_t i = 1; i < n; ++i) {
    const int step = -static_cast<int>(i);
    const u32 elt = P_.galois_elt_from_step(step);
    if (!elt) throw std::invalid_argument("step count too large");
    if (find_key(keyset, elt)) {
      seqs.push_back({step});
      continue;
    }
    std::vector<int> terms = naf_steps(step), eff;
    if (terms.size() == 1) throw std::invalid_argument("Galois key not present");
    for (int s : terms)
      if (static_cast<u64>(s < 0 ? -s : s) != P_.N / 2) eff.push_back(s);
    for (int s : eff)
      if (!find_key(keyset, P_.galois_elt_from_step(s))) throw std::invalid_argument("Galois key not present");
    seqs.push_back(eff);
  }
  std::sort(seqs.begin(), seqs.end());
  size_t depth = 0;
  for (auto &s : seqs) depth = std::max(depth, s.size());
  Scope sc(*this);
  const size_t ctw = ct_words();
  std::vector<u64 *> level(depth + 1);
  for (size_t d = 1; d <= depth; ++d) level[d] = scratch(items * ctw);
  if (out != a) dev_.d2d(out, a, items * ctw * 8);
  u64 *acc = out;
  const u64 *src0 = a;
  u64 *acopy = nullptr;
  if (out == a) {  // keep the operand intact while accumulating
    acopy = scratch(items * ctw);
    dev_.d2d(acopy, a, items * ctw * 8);
    src0 = acopy;
  }
  std::vector<int> path;
  for (auto &s : seqs) {
    size_t common = 0;
    while (common < path.size() && common < s.size() && path[common] == s[common]) ++common;
    path.resize(common);
    for (size_t d = common; d < s.size(); ++d) {
      const u32 e = P_.galois_elt_from_step(s[d]);
      apply_galois(d == 0 ? src0 : level[d], e, need_key(keyset, e), level[d + 1], items);
      path.push_back(s[d]);
    }
    if (s.empty())
      add(acc, src0, acc, items);
    else
      add(acc, level[s.size()], acc, items);
  }
}

}  // namespace hhe
