// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A flat C ABI over the UNMODIFIED reference hot path so that Python tests and
// bench.py's cpu_baseline / --impl reference leg can drive it through ctypes.
// It is compiled by oracle/Makefile together with the reference's own sources
// *where they lie* under /root/reference (src/pasta/*.cpp, src/util/sealhelper.cpp,
// libs/keccak/*.c) and linked against the vendored libs/seal/lib/libseal-4.0.a;
// the output goes to oracle/_ref/libhhe_ref.so (git-ignored, ships via gpurun).
// No reference source is copied into this repository.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load
// this library. The product (libhhe_b200.so) never links or calls it.
//
// Reference entry points exercised:
//   pasta::PASTA_SEAL::decomposition      src/pasta/pasta_3_seal.cpp:106-172
//   pasta::SEALZpCipher::mask / flatten   src/pasta/SEAL_Cipher.cpp:161-181
//   pasta::PASTA::encrypt / Pasta::*      src/pasta/pasta_3_plain.cpp:9-26,56-129,286-295
//   sealhelper::packed_enc_multiply       src/util/sealhelper.cpp:268-274
//   sealhelper::encrypted_vec_sum         src/util/sealhelper.cpp:379-392
//   seal::Evaluator / BatchEncoder        libs/seal/include/SEAL-4.0/seal/evaluator.h
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "seal/seal.h"
#include "pasta_3_plain.h"
#include "pasta_3_seal.h"
#include "sealhelper.h"

using namespace seal;

namespace {

struct Ref {
  std::shared_ptr<SEALContext> ctx;
  std::unique_ptr<KeyGenerator> keygen;
  SecretKey sk;
  PublicKey pk;
  RelinKeys rk;
  GaloisKeys gk[2];  // keyset 0: explicit step list, keyset 1: SEAL default set
  bool have_gk[2] = {false, false};
  std::unique_ptr<Encryptor> enc;
  std::unique_ptr<Evaluator> eval;
  std::unique_ptr<Decryptor> dec;
  std::unique_ptr<BatchEncoder> benc;
  size_t N = 0, L = 0, K = 0;
  uint64_t t = 0;
  std::string err;
};

thread_local std::string g_err;

Ciphertext make_ct(Ref *r, const uint64_t *data, int size) {
  Ciphertext ct(*r->ctx);
  ct.resize(*r->ctx, r->ctx->first_parms_id(), static_cast<size_t>(size));
  std::memcpy(ct.data(), data, sizeof(uint64_t) * size * r->L * r->N);
  ct.is_ntt_form() = false;
  return ct;
}

void dump_ct(Ref *r, const Ciphertext &ct, uint64_t *out) {
  std::memcpy(out, ct.data(), sizeof(uint64_t) * ct.size() * r->L * r->N);
}

Plaintext make_pt(Ref *r, const uint64_t *coeffs) {
  Plaintext pt(r->N);
  std::memcpy(pt.data(), coeffs, sizeof(uint64_t) * r->N);
  // SEAL keeps plaintexts trimmed of leading zero coefficients
  size_t n = r->N;
  while (n > 0 && coeffs[n - 1] == 0) n--;
  pt.resize(n);
  return pt;
}

const GaloisKeys &keyset(Ref *r, int which) {
  if (which < 0 || which > 1 || !r->have_gk[which]) throw std::invalid_argument("galois keyset not generated");
  return r->gk[which];
}

template <typename F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::invalid_argument &e) {
    g_err = std::string("invalid_argument: ") + e.what();
    return 1;
  } catch (const std::logic_error &e) {
    g_err = std::string("logic_error: ") + e.what();
    return 2;
  } catch (const std::exception &e) {
    g_err = std::string("runtime_error: ") + e.what();
    return 3;
  }
}

}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// nq == 0 -> CoeffModulus::BFVDefault(N) (the reference's create_context, SEAL_Cipher.cpp:38-68);
// otherwise the given primes with sec_level none (small-N test rings).
// steps/n_steps: rotation steps for keyset 0 (0 = column swap, as in add_gk_indices).
// want_default_gk != 0 additionally generates SEAL's default power-of-two set (keyset 1).
void *ref_create(uint64_t N, uint64_t t, const uint64_t *q, int nq, uint64_t seed, const int *steps, int n_steps,
                 int want_default_gk) {
  Ref *r = new Ref();
  int rc = guarded([&] {
    EncryptionParameters parms(scheme_type::bfv);
    parms.set_poly_modulus_degree(N);
    sec_level_type sec = sec_level_type::tc128;
    if (nq == 0) {
      parms.set_coeff_modulus(CoeffModulus::BFVDefault(N));
    } else {
      std::vector<Modulus> mods;
      for (int i = 0; i < nq; i++) mods.emplace_back(q[i]);
      parms.set_coeff_modulus(mods);
      sec = sec_level_type::none;
    }
    parms.set_plain_modulus(t);
    prng_seed_type s{};
    for (size_t i = 0; i < s.size(); i++) s[i] = seed * 0x9E3779B97F4A7C15ULL + i;
    parms.set_random_generator(std::make_shared<Blake2xbPRNGFactory>(s));
    r->ctx = std::make_shared<SEALContext>(parms, true, sec);
    if (!r->ctx->parameters_set()) throw std::invalid_argument(r->ctx->parameter_error_message());
    r->keygen = std::make_unique<KeyGenerator>(*r->ctx);
    r->sk = r->keygen->secret_key();
    r->keygen->create_public_key(r->pk);
    r->keygen->create_relin_keys(r->rk);
    if (n_steps > 0) {
      std::vector<int> st(steps, steps + n_steps);
      r->keygen->create_galois_keys(st, r->gk[0]);
      r->have_gk[0] = true;
    }
    if (want_default_gk) {
      r->keygen->create_galois_keys(r->gk[1]);
      r->have_gk[1] = true;
    }
    r->enc = std::make_unique<Encryptor>(*r->ctx, r->pk);
    r->eval = std::make_unique<Evaluator>(*r->ctx);
    r->dec = std::make_unique<Decryptor>(*r->ctx, r->sk);
    r->benc = std::make_unique<BatchEncoder>(*r->ctx);
    r->N = N;
    r->t = t;
    r->K = r->ctx->key_context_data()->parms().coeff_modulus().size();
    r->L = r->ctx->first_context_data()->parms().coeff_modulus().size();
  });
  if (rc) {
    delete r;
    return nullptr;
  }
  return r;
}

void ref_destroy(void *h) { delete static_cast<Ref *>(h); }

// info[0..3] = N, L, K, t
void ref_info(void *h, uint64_t *info) {
  Ref *r = static_cast<Ref *>(h);
  info[0] = r->N;
  info[1] = r->L;
  info[2] = r->K;
  info[3] = r->t;
}

void ref_moduli(void *h, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  auto &m = r->ctx->key_context_data()->parms().coeff_modulus();
  for (size_t i = 0; i < m.size(); i++) out[i] = m[i].value();
}

// psi for each key-level prime (K values), then the plain-modulus root.
void ref_ntt_roots(void *h, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  auto tabs = r->ctx->key_context_data()->small_ntt_tables();
  for (size_t i = 0; i < r->K; i++) out[i] = tabs[i].get_root();
  out[r->K] = r->ctx->first_context_data()->plain_ntt_tables()->get_root();
}

// out = m_sk, gamma, m_tilde, base_B[0..L)
void ref_behz(void *h, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  auto rt = r->ctx->first_context_data()->rns_tool();
  out[0] = rt->m_sk().value();
  out[1] = rt->gamma().value();
  out[2] = rt->m_tilde().value();
  for (size_t i = 0; i < r->L; i++) out[3 + i] = (*rt->base_B())[i].value();
  // psi of each Bsk prime
  auto bt = rt->base_Bsk_ntt_tables();
  for (size_t i = 0; i < r->L + 1; i++) out[3 + r->L + i] = bt[i].get_root();
}

uint32_t ref_galois_elt(void *h, int step) {
  Ref *r = static_cast<Ref *>(h);
  return r->ctx->key_context_data()->galois_tool()->get_elt_from_step(step);
}

// kind 0/1: galois keyset 0/1 (elt = galois element); kind 2: relin key (elt ignored).
// out layout [L digits][2][K][N] (SEAL's own, NTT form). Returns 0 on success, -1 if the key is absent.
int ref_get_ksk(void *h, int kind, uint32_t elt, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  const std::vector<PublicKey> *key = nullptr;
  if (kind == 2) {
    key = &r->rk.key(2);
  } else {
    if (!r->have_gk[kind] || !r->gk[kind].has_key(elt)) return -1;
    key = &r->gk[kind].key(elt);
  }
  size_t per = 2 * r->K * r->N;
  for (size_t j = 0; j < key->size(); j++) std::memcpy(out + j * per, (*key)[j].data().data(), per * sizeof(uint64_t));
  return 0;
}

// Number of keys and their galois elements in keyset `kind` (elts may be NULL).
int ref_list_galois(void *h, int kind, uint32_t *elts) {
  Ref *r = static_cast<Ref *>(h);
  if (!r->have_gk[kind]) return 0;
  int n = 0;
  for (uint32_t e = 1; e < 2 * r->N; e += 2)
    if (r->gk[kind].has_key(e)) {
      if (elts) elts[n] = e;
      n++;
    }
  return n;
}

void ref_public_key(void *h, uint64_t *out) {  // [2][K][N], NTT form, key level (seal::PublicKey::data())
  Ref *r = static_cast<Ref *>(h);
  std::memcpy(out, r->pk.data().data(), sizeof(uint64_t) * 2 * r->K * r->N);
}

void ref_secret_key(void *h, uint64_t *out) {  // [K][N], NTT form (SEAL's storage)
  Ref *r = static_cast<Ref *>(h);
  std::memcpy(out, r->sk.data().data(), sizeof(uint64_t) * r->K * r->N);
}

int ref_encode(void *h, const uint64_t *slots, size_t n, uint64_t *pt_out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    std::vector<uint64_t> v(slots, slots + n);
    Plaintext p;
    r->benc->encode(v, p);
    std::memset(pt_out, 0, sizeof(uint64_t) * r->N);
    std::memcpy(pt_out, p.data(), sizeof(uint64_t) * p.coeff_count());
  });
}

int ref_encrypt(void *h, const uint64_t *slots, size_t n, uint64_t *ct_out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    std::vector<uint64_t> v(slots, slots + n);
    Plaintext p;
    r->benc->encode(v, p);
    Ciphertext ct;
    r->enc->encrypt(p, ct);
    dump_ct(r, ct, ct_out);
  });
}

// slots_out[N]; returns noise budget (bits) in *budget if non-NULL.
int ref_decrypt(void *h, const uint64_t *ct, int size, uint64_t *slots_out, int *budget) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext c = make_ct(r, ct, size);
    Plaintext p;
    r->dec->decrypt(c, p);
    std::vector<uint64_t> v;
    r->benc->decode(p, v);
    std::memcpy(slots_out, v.data(), sizeof(uint64_t) * r->N);
    if (budget) *budget = r->dec->invariant_noise_budget(c);
  });
}

int ref_ntt(void *h, int limb, int inverse, uint64_t *data) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    auto tabs = r->ctx->key_context_data()->small_ntt_tables();
    if (inverse)
      util::inverse_ntt_negacyclic_harvey(data, tabs[limb]);
    else
      util::ntt_negacyclic_harvey(data, tabs[limb]);
  });
}

// NTT over the BEHZ auxiliary primes (index 0..L-1 = base_B, L = m_sk).
int ref_ntt_bsk(void *h, int idx, int inverse, uint64_t *data) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    auto tabs = r->ctx->first_context_data()->rns_tool()->base_Bsk_ntt_tables();
    if (inverse)
      util::inverse_ntt_negacyclic_harvey(data, tabs[idx]);
    else
      util::ntt_negacyclic_harvey(data, tabs[idx]);
  });
}

int ref_add(void *h, const uint64_t *a, const uint64_t *b, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2), y = make_ct(r, b, 2), z;
    r->eval->add(x, y, z);
    dump_ct(r, z, out);
  });
}

int ref_negate(void *h, const uint64_t *a, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2);
    r->eval->negate_inplace(x);
    dump_ct(r, x, out);
  });
}

int ref_add_plain(void *h, const uint64_t *a, const uint64_t *pt, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2);
    Plaintext p = make_pt(r, pt);
    r->eval->add_plain_inplace(x, p);
    dump_ct(r, x, out);
  });
}

int ref_multiply_plain(void *h, const uint64_t *a, const uint64_t *pt, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2);
    Plaintext p = make_pt(r, pt);
    r->eval->multiply_plain_inplace(x, p);
    dump_ct(r, x, out);
  });
}

int ref_rotate_rows(void *h, const uint64_t *a, int steps, int which_keys, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2);
    r->eval->rotate_rows_inplace(x, steps, keyset(r, which_keys));
    dump_ct(r, x, out);
  });
}

int ref_rotate_columns(void *h, const uint64_t *a, int which_keys, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2);
    r->eval->rotate_columns_inplace(x, keyset(r, which_keys));
    dump_ct(r, x, out);
  });
}

// sealhelper::packed_enc_multiply (Evaluator::multiply): out is a size-3 ciphertext.
int ref_multiply(void *h, const uint64_t *a, const uint64_t *b, uint64_t *out3) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2), y = make_ct(r, b, 2), z;
    sealhelper::packed_enc_multiply(x, y, z, *r->eval);
    dump_ct(r, z, out3);
  });
}

int ref_square(void *h, const uint64_t *a, uint64_t *out3) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2);
    r->eval->square_inplace(x);
    dump_ct(r, x, out3);
  });
}

int ref_relinearize(void *h, const uint64_t *a3, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a3, 3);
    r->eval->relinearize_inplace(x, r->rk);
    dump_ct(r, x, out);
  });
}

int ref_exponentiate3(void *h, const uint64_t *a, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2);
    r->eval->exponentiate_inplace(x, 3, r->rk);
    dump_ct(r, x, out);
  });
}

// sealhelper::encrypted_vec_sum with galois keyset `which_keys`.
int ref_vec_sum(void *h, const uint64_t *a, size_t n, int which_keys, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    Ciphertext x = make_ct(r, a, 2), z;
    sealhelper::encrypted_vec_sum(x, z, *r->eval, keyset(r, which_keys), n);
    dump_ct(r, z, out);
  });
}

// pasta::PASTA_SEAL::decomposition. enc_key: size-2 ct. out: ceil(n/128) size-2 cts.
int ref_pasta_decompose(void *h, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n, int use_bsgs,
                        uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    pasta::PASTA_SEAL hhe(r->ctx, r->pk, r->sk, r->rk, keyset(r, 0));
    hhe.activate_bsgs(use_bsgs != 0);
    std::vector<uint64_t> c(sym_ct, sym_ct + n);
    std::vector<Ciphertext> k{make_ct(r, enc_key, 2)};
    std::vector<Ciphertext> res = hhe.decomposition(c, k, true);
    size_t per = 2 * r->L * r->N;
    for (size_t b = 0; b < res.size(); b++) dump_ct(r, res[b], out + b * per);
  });
}

int ref_mask(void *h, const uint64_t *a, const uint64_t *mask, size_t n, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    pasta::PASTA_SEAL hhe(r->ctx, r->pk, r->sk, r->rk, keyset(r, 0));
    Ciphertext x = make_ct(r, a, 2);
    std::vector<uint64_t> m(mask, mask + n);
    hhe.mask(x, m);
    dump_ct(r, x, out);
  });
}

int ref_flatten(void *h, const uint64_t *cts, size_t count, int which_keys, uint64_t *out) {
  Ref *r = static_cast<Ref *>(h);
  return guarded([&] {
    pasta::PASTA_SEAL hhe(r->ctx, r->pk, r->sk, r->rk, keyset(r, 0));
    size_t per = 2 * r->L * r->N;
    std::vector<Ciphertext> in;
    for (size_t i = 0; i < count; i++) in.push_back(make_ct(r, cts + i * per, 2));
    Ciphertext z;
    hhe.flatten(in, z, keyset(r, which_keys));
    dump_ct(r, z, out);
  });
}

// ---- plain PASTA-3 (symmetric side) -------------------------------------------------------------

int ref_pasta_plain(const uint64_t *key256, uint64_t p, const uint64_t *in, size_t n, int decrypt, uint64_t *out) {
  return guarded([&] {
    std::vector<uint64_t> k(key256, key256 + 256);
    pasta::PASTA c(k, p);
    std::vector<uint64_t> v(in, in + n);
    std::vector<uint64_t> o = decrypt ? c.decrypt(v) : c.encrypt(v);
    std::memcpy(out, o.data(), sizeof(uint64_t) * n);
  });
}

// One affine layer's SHAKE-derived material for (nonce, counter), layer index `layer` (0..3):
// mat1[128*128], mat2[128*128], rc[256] (rc1 then rc2), consuming the stream in the reference's order.
int ref_pasta_layer_material(uint64_t p, uint64_t nonce, uint64_t counter, int layer, uint64_t *mat1, uint64_t *mat2,
                             uint64_t *rc) {
  return guarded([&] {
    pasta::Pasta ps(p);
    ps.init_shake(nonce, counter);
    for (int l = 0; l <= layer; l++) {
      auto m1 = ps.get_random_matrix();
      auto m2 = ps.get_random_matrix();
      auto v = ps.get_rc_vec(4096);
      if (l == layer) {
        for (size_t i = 0; i < 128; i++)
          for (size_t j = 0; j < 128; j++) {
            mat1[i * 128 + j] = m1[i][j];
            mat2[i * 128 + j] = m2[i][j];
          }
        for (size_t i = 0; i < 128; i++) {
          rc[i] = v[i];
          rc[128 + i] = v[4096 + i];
        }
      }
    }
  });
}

// ---- CPU baseline: the reference's decomposition on `threads` host threads ----------------------
// Every thread owns a PASTA_SEAL (as BaseCSP::decompose does, CSP.cpp:238-242) and transciphers
// `blocks_per_thread` blocks. Returns wall seconds (max over threads) or <0 on error.
double ref_bench_decompose(void *h, const uint64_t *enc_key, int threads, int blocks_per_thread, int use_bsgs) {
  Ref *r = static_cast<Ref *>(h);
  std::atomic<int> failed{0};
  std::vector<std::thread> pool;
  std::vector<std::unique_ptr<pasta::PASTA_SEAL>> hhe(threads);
  for (int i = 0; i < threads; i++) {
    hhe[i] = std::make_unique<pasta::PASTA_SEAL>(r->ctx, r->pk, r->sk, r->rk, keyset(r, 0));
    hhe[i]->activate_bsgs(use_bsgs != 0);
  }
  Ciphertext key = make_ct(r, enc_key, 2);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < threads; i++) {
    pool.emplace_back([&, i] {
      try {
        std::vector<uint64_t> c(static_cast<size_t>(blocks_per_thread) * 128);
        for (size_t j = 0; j < c.size(); j++) c[j] = (j * 2654435761ULL + i) % r->t;
        std::vector<Ciphertext> k{key};
        auto res = hhe[i]->decomposition(c, k, true);
        if (res.size() != static_cast<size_t>(blocks_per_thread)) failed++;
      } catch (...) {
        failed++;
      }
    });
  }
  for (auto &th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  if (failed) return -1.0;
  return std::chrono::duration<double>(t1 - t0).count();
}

// Time `reps` calls of one primitive on one thread; returns seconds per call.
// op: 0 ntt fwd, 1 ntt inv (limb 0), 2 rotate_rows(-1), 3 relinearize(of a square), 4 multiply_plain, 5 multiply
double ref_bench_primitive(void *h, const uint64_t *ct, int op, int reps) {
  Ref *r = static_cast<Ref *>(h);
  try {
    Ciphertext x = make_ct(r, ct, 2);
    std::vector<uint64_t> limb(ct, ct + r->N);
    auto tabs = r->ctx->key_context_data()->small_ntt_tables();
    Ciphertext sq;
    r->eval->square(x, sq);
    Plaintext p;
    std::vector<uint64_t> v(r->N, 3);
    r->benc->encode(v, p);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) {
      Ciphertext y;
      switch (op) {
        case 0: util::ntt_negacyclic_harvey(limb.data(), tabs[0]); break;
        case 1: util::inverse_ntt_negacyclic_harvey(limb.data(), tabs[0]); break;
        case 2: r->eval->rotate_rows(x, -1, keyset(r, 0), y); break;
        case 3: r->eval->relinearize(sq, r->rk, y); break;
        case 4: r->eval->multiply_plain(x, p, y); break;
        case 5: r->eval->multiply(x, x, y); break;
        default: return -1.0;
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count() / reps;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1.0;
  }
}

// ---- SEAL wire format (Ciphertext/GaloisKeys/RelinKeys::save/load, serialization.h:49-91): checker for the codec ----
// level 0 = first (data) parms_id, 1 = key parms_id
void ref_parms_id(void *h, int level, uint64_t *out4) {
  Ref *r = static_cast<Ref *>(h);
  const parms_id_type &id = level ? r->ctx->key_parms_id() : r->ctx->first_parms_id();
  for (size_t i = 0; i < 4; i++) out4[i] = id[i];
}

// compr: 0 none, 1 zlib, 2 zstd. Returns bytes written (<= cap) or -1.
long long ref_ct_save(void *h, const uint64_t *ct, int size, int compr, uint8_t *out, size_t cap) {
  Ref *r = static_cast<Ref *>(h);
  long long n = -1;
  guarded([&] {
    Ciphertext c = make_ct(r, ct, size);
    n = static_cast<long long>(c.save(reinterpret_cast<seal_byte *>(out), cap, static_cast<compr_mode_type>(compr)));
  });
  return n;
}

// Ciphertext::load(context, bytes): returns bytes consumed or -1; *size_out = ciphertext size, ct_out [size][L][N]
long long ref_ct_load(void *h, const uint8_t *in, size_t len, uint64_t *ct_out, int *size_out) {
  Ref *r = static_cast<Ref *>(h);
  long long n = -1;
  guarded([&] {
    Ciphertext c;
    n = static_cast<long long>(c.load(*r->ctx, reinterpret_cast<const seal_byte *>(in), len));
    *size_out = static_cast<int>(c.size());
    dump_ct(r, c, ct_out);
  });
  return n;
}

// kind 0/1: GaloisKeys keyset, 2: RelinKeys. out == NULL: returns the upper bound save_size(compr).
long long ref_keys_save(void *h, int kind, int compr, uint8_t *out, size_t cap) {
  Ref *r = static_cast<Ref *>(h);
  long long n = -1;
  guarded([&] {
    const compr_mode_type cm = static_cast<compr_mode_type>(compr);
    const KSwitchKeys &k = kind == 2 ? static_cast<const KSwitchKeys &>(r->rk) : static_cast<const KSwitchKeys &>(keyset(r, kind));
    if (!out)
      n = static_cast<long long>(k.save_size(cm));
    else
      n = static_cast<long long>(k.save(reinterpret_cast<seal_byte *>(out), cap, cm));
  });
  return n;
}

// GaloisKeys::load(context, bytes) round trip check: loads, then returns the number of keys present (or -1)
int ref_galois_load_count(void *h, const uint8_t *in, size_t len) {
  Ref *r = static_cast<Ref *>(h);
  int n = -1;
  guarded([&] {
    GaloisKeys g;
    g.load(*r->ctx, reinterpret_cast<const seal_byte *>(in), len);
    n = 0;
    for (auto &v : g.data()) n += !v.empty();
  });
  return n;
}

}  // extern "C"
