// Engine implementation (2/4): transforms, element-wise operations, encode, multiply_plain.
#include "engine_impl.h"

namespace hhe {

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
                                     limb_stride ? limb_stride : static_cast<size_t>(2) << LOGV, pf_ntt_, static_cast<int>(items * limbs)};
        dev_.launch_cluster2(body, items * limbs * 2, half_threads(LOGV), half_smem(LOGV));
      });
      return;
    }
    if (all_f64 && inverse) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        using Body = InvClusterBody<LOGV, PlanScaled>;
        Body body{PlanScaled{in, out, map, limbs, item_stride ? item_stride : static_cast<size_t>(limbs) << (LOGV + 1),
                             limb_stride ? limb_stride : static_cast<size_t>(2) << LOGV},
                  dC_, twref(), pf_ntt_, static_cast<int>(items * limbs)};
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

void Engine::add_plain(const u64 *a, const u64 *pt, size_t pstride, u64 *out, size_t items, bool negate_first, const u32 *ptidx,
                       const u32 *aidx) {
  const size_t total = items * ct_words();
  AddPlainBody body{a, pt, pstride, out, dC_, negate_first ? 1 : 0, total, ptidx, aidx};
  dev_.launch(body, ew_grid(total), kEwThreads, 0);
}

void Engine::broadcast(const u64 *src, u64 *out, size_t words, size_t items) {
  BroadcastBody body{src, out, words, words * items};
  dev_.launch(body, ew_grid(words * items), kEwThreads, 0);
}

void Engine::encode_slots(const u64 *slots, size_t sstride, const u32 *lens, u32 n, u64 *pt, size_t items) {
  if (n > P_.N) throw std::invalid_argument("values_matrix size exceeds slot count");
  if (split_ || table_is_f64(P_, P_.tab_plain())) {
    // slot vector in global memory, then the inverse transform mod t as a batched NTT: split transforms at N = 32768, the FP64
    // two-CTA cluster kernels when the ring is on the FP64 path (half the time of the whole-limb integer encode kernel)
    EncodeScatterBody body{slots, sstride, lens, n, nullptr, nullptr, dIndex_, pt, dC_, kSlots, 0, 0, 0};
    dev_.launch(body, items, 256, 0);
    TabMap mt{};
    mt.id[0] = static_cast<unsigned char>(P_.tab_plain());
    ntt(pt, pt, items, 1, mt, true);
    return;
  }
  HHE_DISPATCH_LOG(P_.logn, {
    EncodeBody<LOGV> body{slots, sstride, lens, n, nullptr, nullptr, dIndex_, pt, dC_, twref(), kSlots, 0, 0, 0};
    dev_.launch(body, items, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::encode_material(const u32 *material, const u32 *mat_index, int mode, int layer, int diag, u64 *pt, size_t items, int ndiag) {
  if (split_ || table_is_f64(P_, P_.tab_plain())) {
    const size_t all = items * static_cast<size_t>(ndiag > 1 ? ndiag : 1);
    EncodeScatterBody body{nullptr, 0, nullptr, 0, material, mat_index, dIndex_, pt, dC_, mode, layer, diag, ndiag > 1 ? static_cast<int>(items) : 0};
    dev_.launch(body, all, 256, 0);
    TabMap mt{};
    mt.id[0] = static_cast<unsigned char>(P_.tab_plain());
    ntt(pt, pt, all, 1, mt, true);
    return;
  }
  HHE_DISPATCH_LOG(P_.logn, {
    EncodeBody<LOGV> body{nullptr, 0, nullptr, 0, material, mat_index, dIndex_, pt, dC_, twref(), mode, layer, diag,
                          ndiag > 1 ? static_cast<int>(items) : 0};
    dev_.launch(body, items * static_cast<size_t>(ndiag > 1 ? ndiag : 1), ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::lift_ntt(const u64 *pt, u64 *D, size_t items, const u32 *nolift) {
  if (split_) {  // N = 32768: element-wise lift, then split forward transforms (out of place: D is written by the lift first)
    Scope sc(*this);
    const size_t total = items * P_.L * P_.N;
    u64 *lifted = scratch(total);
    LiftBody body{pt, lifted, dC_, nolift, total};
    dev_.launch(body, ew_grid(total), kEwThreads, 0);
    ntt(lifted, D, items, P_.L, map_mod(P_.L, P_.L, 0), false);
    return;
  }
  if (half_fwd_) {
    HHE_DISPATCH_LOG(P_.logn - 1, {
      LiftNttHalfBody<LOGV> body{pt, D, dC_, twref(), nolift};
      dev_.launch(body, items * P_.L * 2, half_threads(LOGV), half_smem(LOGV));
    });
    return;
  }
  HHE_DISPATCH_LOG(P_.logn, {
    LiftNttBody<LOGV> body{pt, D, dC_, twref(), nolift};
    dev_.launch(body, items * P_.L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
  });
}

void Engine::ntt_mac(const u64 *ct, const u64 *D, size_t dstride, u64 *sum, size_t items, bool first, int comps, size_t sum_off,
                     u64 *ntt_out, const u32 *didx) {
  if (split_) {  // N = 32768: split forward transforms of the ciphertext limbs, then the element-wise product and sum
    Scope sc(*this);
    const size_t total = items * comps * P_.L * P_.N;
    u64 *xn = ntt_out ? ntt_out : scratch(total);
    ntt(ct, xn, items, comps * P_.L, map_mod(comps * P_.L, P_.L, 0), false);
    DyadicMacBody body{xn, D, dstride, sum, dC_, first ? 1 : 0, comps, ct_words(), sum_off, didx, total};
    dev_.launch(body, ew_grid(total), kEwThreads, 0);
    return;
  }
  if (half_fwd_) {
    HHE_DISPATCH_LOG(P_.logn - 1, {
      NttMacHalfBody<LOGV> body{ct, D, dstride, sum, dC_, twref(), first ? 1 : 0, comps, ct_words(), sum_off, ntt_out, didx,
                                pf_limbs_, static_cast<int>(items * comps * P_.L)};
      dev_.launch(body, items * comps * P_.L * 2, half_threads(LOGV), half_smem(LOGV));
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
  // [0, N): the permutation (gather form); [N, 2N): its inverse (scatter form, ntt_perm_inv)
  std::vector<u32> both(2 * N);
  for (u64 i = 0; i < N; ++i) {
    both[i] = table[i];
    both[N + table[i]] = static_cast<u32>(i);
  }
  u32 *d = static_cast<u32 *>(dev_.dmalloc(2 * N * sizeof(u32)));
  dev_.h2d(d, both.data(), 2 * N * sizeof(u32));
  dev_.sync();
  perms_[elt] = d;
  return d;
}

void Engine::ct_intt(u64 *ct, size_t items, int size) {
  ntt(ct, ct, items, size * P_.L, map_mod(size * P_.L, P_.L, 0), true);
}

void Engine::multiply_plain(const u64 *a, const u64 *pt, size_t pstride, u64 *out, size_t items, const u32 *nolift) {
  Scope sc(*this);
  const size_t ditems = pstride ? items : 1;
  u64 *D = scratch(ditems * P_.L * P_.N);
  lift_ntt(pt, D, ditems, nolift);
  ntt_mac(a, D, pstride ? static_cast<size_t>(P_.L) * P_.N : 0, out, items, true);
  ct_intt(out, items);
}

}  // namespace hhe
