// Engine implementation (1/4): construction, scratch arena, keys. Orchestration of the sm_100a kernels in kernels.h / keccak.h.
#include "engine_impl.h"

namespace hhe {

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
  batch_ = auto_batch();
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
  if (const char *v = std::getenv("HHE_KS_SPLIT_MAX")) ks_split_max_ = std::atoi(v);
  cluster_inv_ = half_fwd_ && !getenv_flag("HHE_NO_CLUSTER");
  dev_.strict_cluster = getenv_flag("HHE_STRICT_CLUSTER");
#ifdef HHE_CUDA
  dev_.pdl = !getenv_flag("HHE_NO_PDL");
#endif
  dev_.ordinal = device;
  // L2 prefetch (cp.async.bulk.prefetch.L2 by one thread per CTA pair) of the operands of the pair that starts one full wave
  // later. Measured on B200 (profiles/r2_ab_prefetch.txt): the plain transforms gain 3-8 % (their load phase is a chain of DRAM
  // round trips), but corr0_mac / ntt_mac, which already move 4 TB/s, lose 20 % (the prefetch bursts queue in front of the
  // demand loads) and ks_digits (L2-resident operands) does not change. So only Engine::ntt uses it by default;
  // HHE_PF_ALL=1 switches it on everywhere, HHE_NO_PREFETCH=1 off.
  if (!getenv_flag("HHE_NO_PREFETCH")) {
    pf_ntt_ = dev_.sm_count;  // two CTAs per SM in flight = sm_count limbs
    if (getenv_flag("HHE_PF_ALL")) {
      pf_limbs_ = dev_.sm_count;
      pf_items_ = (2 * dev_.sm_count + 2 * P_.K - 1) / (2 * P_.K) + 1;
    }
  }
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
  dev_.d2h_release();
#ifdef HHE_CUDA
  if (dev_.own_stream) cudaStreamDestroy(dev_.stream);
#endif
}

int Engine::auto_batch() {
  int b = 2 * dev_.sm_count;
#ifdef HHE_CUDA
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    // per block: BSGS keeps 16 baby rotations + 16 lifted diagonals + state/temporaries ~ 22 ciphertexts + 16 L-limb plaintexts
    const size_t per_block = (22 * ct_words() + 16 * static_cast<size_t>(P_.L) * P_.N + 2 * static_cast<size_t>(P_.K) * P_.N) * 8;
    size_t already = 0;
    for (auto &c : chunks_) already += c.words * 8;  // the arena is reused
    const size_t budget = (free_b + already) / 10 * 8;
    const size_t fit = budget / per_block;
    if (fit < static_cast<size_t>(b)) b = static_cast<int>(fit < 1 ? 1 : fit);
  }
#endif
  return b;
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

}  // namespace hhe
