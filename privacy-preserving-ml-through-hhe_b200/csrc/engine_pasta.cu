// Engine implementation (4/4): PASTA-3 transciphering, mask / flatten, encrypted_vec_sum.
#include "engine_impl.h"
#include "keccak.h"

namespace hhe {

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
  // the diagonals are encoded and lifted G at a time (one launch each instead of G): G = the largest power of two <= 128 whose lifted
  // transforms (nd x G MiB at N = 16384) stay within about 2.5 GiB -- all 128 of a layer for a service-sized request, 8 for the
  // full 296-block batch
  int G = 128;
  while (G > 1 && static_cast<size_t>(G) * nd * dw * 8 > (static_cast<size_t>(5) << 29)) G >>= 1;
  if (const char *v = std::getenv("HHE_DIAG_GROUP")) G = std::max(1, std::min(128, std::atoi(v)));
  while (kPastaT % G) --G;
  u64 *tmp = scratch(nb * ctw), *sum = scratch(nb * ctw), *pt = scratch(static_cast<size_t>(G) * nd * N), *Dg = scratch(static_cast<size_t>(G) * nd * dw);
  u64 *D = Dg;
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
  encode_material(mat, nullptr, kDiag, layer, 0, pt, nd, G);
  lift_ntt(pt, Dg, nd * G);
  ntt_mac(state, D, ds, sum, nb, true, 2, 0, stn, didx);
  // half-limb FP64 kernels: both components of the rotated ciphertext stay NTT-resident (Corr0MacHalfBody, comps = 2), the
  // coefficient form of c1 is never stored (only its Galois image g1, the next key switch's digits), and the two NTT-resident
  // arrays are kept in the order their next reader wants (permuted by the rotation's slot permutation: corr_mac scatters its outputs
  // through the inverse permutation, so neither it nor ks_digits has a dependent gather on its critical path)
  const bool pair = half_fwd_ && cluster_inv_ && !getenv_flag("HHE_NO_CORR_PAIR");
  const bool fuse_tail = !getenv_flag("HHE_NO_ROT_TAIL");
  const bool scatter = pair && fuse_tail && !getenv_flag("HHE_NO_SCATTER");
  if (scatter) {
    PermCopyBody p0{stn, c0a, perm, ctw, dw, dC_, L, nb * dw}, p1{stn + dw, c1n, perm, ctw, dw, dC_, L, nb * dw};
    dev_.launch(p0, ew_grid(nb * dw), kEwThreads, 0);
    dev_.launch(p1, ew_grid(nb * dw), kEwThreads, 0);
  } else {
    strided_copy(stn, ctw, c0a, dw, dw, nb);
    strided_copy(stn + dw, ctw, c1n, dw, dw, nb);
  }
  strided_copy(state + dw, ctw, c1c, dw, dw, nb);
  u64 *c0_in = c0a, *c0_out = c0b;
  TabMap msp2{};
  msp2.id[0] = msp2.id[1] = static_cast<unsigned char>(K - 1);
  {  // g1 = galois(c1) in coefficient form: the digits of the first key switch (later ones come out of intt_moddown)
    GaloisBody gb{c1c, g1, dC_, e1_inv, nb * dw};
    dev_.launch(gb, ew_grid(nb * dw), kEwThreads, 0);
  }
  const u32 *pinv = scatter ? ntt_perm_inv(e1) : nullptr;
  for (int i = 1; i < kPastaT; ++i) {
    launch_ks_digits(g1, dw, k1, acc, nb, c1n, dw, scatter ? nullptr : perm);
    // inverse NTT of the two special limbs acc[0][K-1], acc[1][K-1] (K*N words apart inside an item), then of acc[1][i<L]
    // with the ModDown and the next rotation's Galois map fused into the store
    ntt(acc + static_cast<size_t>(K - 1) * N, acc + static_cast<size_t>(K - 1) * N, nb, 2, msp2, true, static_cast<size_t>(2) * K * N,
        static_cast<size_t>(K) * N);
    if (i % G == 0) {
      encode_material(mat, nullptr, kDiag, layer, i, pt, nd, G);
      lift_ntt(pt, Dg, nd * G);
    }
    D = Dg + static_cast<size_t>(i % G) * nd * dw;
    if (pair && fuse_tail) {
      // corr_mac and intt_moddown in one launch (RotTailBody): they are independent and overlap
      HHE_DISPATCH_LOG(P_.logn - 1, {
        RotTailBody<LOGV> body{
            Corr0MacHalfBody<LOGV>{acc, c0_in, c0_out, perm, D, sum, dC_, twref(), ds, didx, pf_limbs_, static_cast<int>(nb * L), 2, c1n, dw, dw,
                                   pinv},
            InvClusterBody<LOGV, PlanModDownGalois>{PlanModDownGalois{acc, nullptr, g1, e1, P_.logn}, dC_, twref(), pf_limbs_, static_cast<int>(nb * L)},
            static_cast<int>(nb * L * 4)};
        dev_.launch_cluster2(body, nb * L * 6, half_threads(LOGV), half_smem(LOGV));
      });
      std::swap(c0_in, c0_out);
      continue;
    }
    if (cluster_inv_) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        using Body = InvClusterBody<LOGV, PlanModDownGalois>;
        Body body{PlanModDownGalois{acc, pair ? nullptr : c1c, g1, e1, P_.logn}, dC_, twref(), pf_limbs_, static_cast<int>(nb * L)};
        dev_.launch_cluster2(body, nb * L * 2, half_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
    } else {
      HHE_DISPATCH_LOG(P_.logn, {
        InttModDownBody<LOGV> body{acc, c1c, g1, dC_, twref(), e1};
        dev_.launch(body, nb * L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
    }
    if (half_fwd_) {
      HHE_DISPATCH_LOG(P_.logn - 1, {
        Corr0MacHalfBody<LOGV> body{acc, c0_in, c0_out, perm, D, sum, dC_, twref(), ds, didx, pf_limbs_, static_cast<int>(nb * L), pair ? 2 : 1, c1n, dw, dw,
                                    nullptr};
        dev_.launch(body, nb * L * (pair ? 4 : 2), half_threads(LOGV), half_smem(LOGV));
      });
    } else {
      HHE_DISPATCH_LOG(P_.logn, {
        Corr0MacBody<LOGV> body{acc, c0_in, c0_out, perm, D, sum, dC_, twref(), ds, didx};
        dev_.launch(body, nb * L, ntt_threads(LOGV), ntt_smem_words(1 << LOGV) * 8);
      });
    }
    std::swap(c0_in, c0_out);
    if (!pair) ntt_mac(c1c, D, ds, sum, nb, false, 1, dw, c1n, didx);
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
  const TabMap mq = map_mod(2 * P_.L, P_.L, 0);
  if (compact_keys_ && half_fwd_ && cluster_inv_ && !getenv_flag("HHE_NO_RESIDENT") && !getenv_flag("HHE_NO_BSGS_RESIDENT")) {
    // The 15 chained baby rotations on the NTT-resident chain (as the diagonal layer's, without the plaintext product): every baby
    // rotation is wanted in NTT form anyway (it is multiplied with 8 diagonals element-wise), so rot[j] is written directly as
    // (c0n, c1n) by rot_tail; per rotation 64 + 2 + 8 + 16 limb transforms instead of 72 + 18 for the generic rotation plus 16 for
    // the forward transform afterwards. Bit-identical (same identities as affine_diagonal_resident).
    const int L = P_.L, K = P_.K;
    const u32 e1 = P_.galois_elt_from_step(-1);
    const W2 *k1 = need_key(0, e1);
    const u32 *perm = ntt_perm(e1);
    u64 *g1 = scratch(nb * dw), *acc = scratch(nb * 2 * K * N);
    TabMap msp2{};
    msp2.id[0] = msp2.id[1] = static_cast<unsigned char>(K - 1);
    ntt(state, rot, nb, 2 * L, mq, false);  // rot[0] = NTT(state)
    {  // digits of the first key switch: galois(c1) in coefficient form, items strided by a whole ciphertext
      Scope inner_sc(*this);
      u64 *c1c = scratch(nb * dw);
      strided_copy(state + dw, ctw, c1c, dw, dw, nb);
      GaloisBody gb{c1c, g1, dC_, inv_mod_2n(e1, 2 * N), nb * dw};
      dev_.launch(gb, ew_grid(nb * dw), kEwThreads, 0);
    }
    for (int j = 1; j < N1; ++j) {
      u64 *prev = rot + static_cast<size_t>(j - 1) * nb * ctw, *cur = rot + static_cast<size_t>(j) * nb * ctw;
      launch_ks_digits(g1, dw, k1, acc, nb, prev + dw, ctw, perm);
      ntt(acc + static_cast<size_t>(K - 1) * N, acc + static_cast<size_t>(K - 1) * N, nb, 2, msp2, true, static_cast<size_t>(2) * K * N,
          static_cast<size_t>(K) * N);
      HHE_DISPATCH_LOG(P_.logn - 1, {
        RotTailBody<LOGV> body{
            Corr0MacHalfBody<LOGV>{acc, prev, cur, perm, nullptr, nullptr, dC_, twref(), 0, nullptr, 0, static_cast<int>(nb * L), 2, cur + dw, ctw, ctw,
                                   nullptr},  // natural order: the baby rotations are multiplied element-wise afterwards
            InvClusterBody<LOGV, PlanModDownGalois>{PlanModDownGalois{acc, nullptr, g1, e1, P_.logn}, dC_, twref(), 0, static_cast<int>(nb * L)},
            static_cast<int>(nb * L * 4)};
        dev_.launch_cluster2(body, nb * L * 6, half_threads(LOGV), half_smem(LOGV));
      });
    }
  } else {
    dev_.d2d(rot, state, nb * ctw * 8);
    for (int j = 1; j < N1; ++j) rotate_rows(rot + (j - 1) * nb * ctw, -1, 0, rot + j * nb * ctw, nb);
    // every baby rotation is multiplied with 8 diagonals: transform each once (in place), then the products are element-wise
    ntt(rot, rot, nb * N1, 2 * P_.L, mq, false);
  }
  for (int k = 0; k < N2; ++k) {
    // the 16 diagonals of this giant step are encoded, lifted and transformed as one batch ([j][nd] items), then one pass over
    // the baby rotations forms the inner sum: every residue of `inner` is written once
    encode_material(mat, nullptr, kDiagBsgs, layer, k * N1, pt, nd, N1);  // [j][nd][N]
    lift_ntt(pt, D, nd * N1);
    DyadicMacNBody mac{rot, D, inner, dC_, N1, nb * ctw, nd * dw, dw, didx, nb * ctw};
    dev_.launch(mac, ew_grid(nb * ctw), kEwThreads, 0);
    if (k == 0) {
      ntt(inner, outer, nb, 2 * P_.L, mq, true);
    } else {
      ntt(inner, inner, nb, 2 * P_.L, mq, true);
      rotate_rows_add(inner, -k * N1, 0, outer, nb);
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
  if (!d_sym) {  // keystream only: the caller finishes every block that uses it (pasta_decompose, shared keystreams)
    dev_.d2d(d_out, state, nb * ctw * 8);
    return;
  }
  encode_slots(d_sym, kPastaT, d_lens, kPastaT, pt, nb);
  add_plain(state, pt, N, d_out, nb, true);  // negate_inplace; add_plain (:168-169)
}

void Engine::pasta_check_keys() {
  if (2 * kPastaT != P_.N && 4 * kPastaT > P_.N) throw std::runtime_error("too little slots for matmul implementation!");
  const u32 e1 = P_.galois_elt_from_step(-1), ec = static_cast<u32>(2 * P_.N - 1);
  need_key(0, e1);
  need_key(0, ec);
  need_key(2, 0);
  if (P_.N / 2 != kPastaT) need_key(0, P_.galois_elt_from_step(kPastaT));
}

bool Engine::share_keystreams() const { return !getenv_flag("HHE_NO_SHARED_KEYSTREAM"); }

// The keystream ciphertext of a block depends on (encrypted key, nonce, counter) only -- not on the data -- and the evaluation is
// deterministic. Records restart their counters (CSP.cpp:247-252), so within one call every block with the same
// counter has the SAME keystream ciphertext, bit for bit: callers with repeated counters compute it once per distinct counter here
// (d_ks[i] = the state of pasta_3_seal.cpp:106-167 for counters[i]) and finish every block with pasta_finish. A stream of distinct
// counters (the headline workload) never takes this path.
void Engine::pasta_keystreams(const u64 *d_enc_key, const std::vector<u64> &counters, u64 nonce, bool use_bsgs, u64 *d_ks) {
  pasta_check_keys();
  Scope sc(*this);
  const size_t n = counters.size(), step = static_cast<size_t>(std::max(1, batch_)), ctw = ct_words();
  u64 *d_ctr = scratch(std::min(step, std::max<size_t>(1, n)));
  for (size_t off = 0; off < n; off += step) {
    const size_t nb = std::min(step, n - off);
    dev_.h2d(d_ctr, counters.data() + off, nb * 8);
    pasta_batch(d_enc_key, nullptr, nullptr, d_ctr, nb, nb, nullptr, nonce, use_bsgs, d_ks + off * ctw);
  }
  dev_.sync();  // `counters` is the caller's host vector
}

// What is a block's own: res[b] = encode(c_b) - ks[idx[b]]  (negate_inplace + add_plain, pasta_3_seal.cpp:168-169)
void Engine::pasta_finish(const u64 *d_ks, const u32 *d_idx, const u64 *d_sym, const u32 *d_lens, size_t nblocks, u64 *d_out) {
  Scope sc(*this);
  const size_t step = static_cast<size_t>(std::max(1, batch_)), ctw = ct_words(), N = P_.N;
  u64 *pt = scratch(std::min(step, std::max<size_t>(1, nblocks)) * N);
  for (size_t off = 0; off < nblocks; off += step) {
    const size_t nb = std::min(step, nblocks - off);
    encode_slots(d_sym + off * kPastaT, kPastaT, d_lens + off, kPastaT, pt, nb);
    add_plain(d_ks, pt, N, d_out + off * ctw, nb, true, nullptr, d_idx + off);
  }
}

void Engine::pasta_decompose(const u64 *d_enc_key, const u64 *d_sym, const u32 *d_lens, const std::vector<u64> &counters,
                             u64 nonce, bool use_bsgs, u64 *d_out) {
  const size_t nblocks = counters.size();
  pasta_check_keys();
  Scope sc(*this);
  const size_t step = static_cast<size_t>(std::max(1, batch_)), ctw = ct_words();
  u64 *d_ctr = scratch(std::min(step, nblocks));
  if (share_keystreams() && nblocks > 1) {  // see pasta_keystreams
    std::map<u64, u32> seen;
    std::vector<u64> uniq;
    std::vector<u32> idx(nblocks);
    for (size_t b = 0; b < nblocks; ++b) {
      auto ins = seen.emplace(counters[b], static_cast<u32>(uniq.size()));
      if (ins.second) uniq.push_back(counters[b]);
      idx[b] = ins.first->second;
    }
    if (uniq.size() < nblocks) {
      u64 *ks = scratch(uniq.size() * ctw);
      pasta_keystreams(d_enc_key, uniq, nonce, use_bsgs, ks);
      u32 *d_idx = reinterpret_cast<u32 *>(scratch((nblocks + 1) / 2));
      dev_.h2d(d_idx, idx.data(), nblocks * 4);
      pasta_finish(ks, d_idx, d_sym, d_lens, nblocks, d_out);
      dev_.sync();  // idx is a host vector the copy above reads
      return;
    }
  }
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
  u64 *gath = scratch(items * ctw);
  for (size_t i = 0; i < per; ++i) {
    strided_copy(in + i * ctw, per * ctw, i ? gath : out, ctw, ctw, items);  // block i of every group in one gather launch
    if (i) rotate_rows_add(gath, -static_cast<int>(i * kPastaT), keyset, out, items);
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
  for (size_t si = 0; si < seqs.size(); ++si) {
    const auto &s = seqs[si];
    size_t common = 0;
    while (common < path.size() && common < s.size() && path[common] == s[common]) ++common;
    path.resize(common);
    // a sequence that no later one extends is a leaf: its last key switch adds straight into the running sum (the addition is folded
    // into the key switch's store); a sequence that is also the prefix of the next one keeps its result for it
    const bool leaf = !s.empty() && !(si + 1 < seqs.size() && seqs[si + 1].size() > s.size() &&
                                      std::equal(s.begin(), s.end(), seqs[si + 1].begin()));
    for (size_t d = common; d < s.size(); ++d) {
      const u32 e = P_.galois_elt_from_step(s[d]);
      const bool fold = leaf && d + 1 == s.size();
      apply_galois(d == 0 ? src0 : level[d], e, need_key(keyset, e), fold ? acc : level[d + 1], items, fold ? acc : nullptr);
      path.push_back(s[d]);
    }
    if (s.empty())
      add(acc, src0, acc, items);
    else if (!leaf)
      add(acc, level[s.size()], acc, items);
    else
      path.pop_back();  // the leaf's last level was never materialised
  }
}

}  // namespace hhe
