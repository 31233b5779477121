// Engine implementation (3/4): key switching (rotations, relinearisation) and the BEHZ ciphertext product.
#include "engine_impl.h"

namespace hhe {

void Engine::galois(const u64 *a, u32 elt, u64 *out, size_t items) {
  const size_t total = items * ct_words();
  GaloisBody body{a, out, dC_, inv_mod_2n(elt, 2 * P_.N), total};
  dev_.launch(body, ew_grid(total), kEwThreads, 0);
}

void Engine::key_switch(const u64 *target, size_t tstride, const W2 *key, const u64 *base0, const u64 *base1,
                        size_t bstride, u64 *out, size_t items, const u64 *accum) {
  if (accum && !(compact_keys_ && !split_ && cluster_inv_)) {
    // only the FP64 cluster path folds the accumulation into its store: elsewhere key switch first, then add
    Scope tmp_scope(*this);
    u64 *tmp = scratch(items * ct_words());
    key_switch(target, tstride, key, base0, base1, bstride, tmp, items, nullptr);
    add(accum, tmp, out, items);
    return;
  }
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
        Body body{PlanModDownAdd{acc, base0, base1, bstride, out, accum}, dC_, twref(), pf_limbs_, static_cast<int>(items * 2 * P_.L)};
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
  if (tmem_ks_ && P_.L <= kKsSplitCluster && items <= static_cast<size_t>(ks_split_max_)) {
    // service-sized request: the digit loop spread over an eight-CTA cluster (one digit's latency instead of L)
    HHE_DISPATCH_LOG(P_.logn - 1, {
      KsDigitsSplitBody<LOGV> body{target, tstride, reinterpret_cast<const double *>(key), acc, dC_, twref(), reuse, reuse_stride, perm};
      dev_.launch_cluster8(body, items * K * 2 * kKsSplitCluster, half_threads(LOGV), KsDigitsSplitBody<LOGV>::smem_bytes());
    });
    return;
  }
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
                                         reuse, reuse_stride, perm, pf_items_};
        dev_.launch(body, items * K * 2, nt, KsDigitsTmemBody<LOGV, 256>::smem_bytes(nt, emulate));
      } else {
        const int nt = std::max(32, std::min(512, G));
        KsDigitsTmemBody<LOGV> body{target, tstride, reinterpret_cast<const double *>(key), acc, dC_, twref(), static_cast<int>(items),
                                    reuse, reuse_stride, perm, pf_items_};
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

void Engine::apply_galois(const u64 *a, u32 elt, const W2 *key, u64 *out, size_t items, const u64 *accum) {
  Scope sc(*this);
  u64 *g = scratch(items * ct_words());
  galois(a, elt, g, items);
  key_switch(g + static_cast<size_t>(P_.L) * P_.N, ct_words(), key, g, nullptr, ct_words(), out, items, accum);
}

// acc += rotate_rows(a, steps): with the step's own key the addition is folded into the key switch's last store (no separate
// 6 MiB element-wise pass); a NAF chain rotates first and adds afterwards
void Engine::rotate_rows_add(const u64 *a, int steps, int keyset, u64 *acc, size_t items) {
  const u32 elt = steps ? P_.galois_elt_from_step(steps) : 0;
  const W2 *key = elt ? find_key(keyset, elt) : nullptr;
  if (key) {
    apply_galois(a, elt, key, acc, items, acc);
    return;
  }
  Scope sc(*this);
  u64 *tmp = scratch(items * ct_words());
  rotate_rows(a, steps, keyset, tmp, items);
  add(acc, tmp, acc, items);
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

}  // namespace hhe
