// Drop-in C++ shim: the reference's class / helper names on top of the libhhe_b200 C ABI.
//
// A maintainer of the reference replaces
//     #include "pasta_3_seal.h"      ->  #include "hhe_seal_shim.h"
//     pasta::PASTA_SEAL               ->  pasta_b200::PASTA_SEAL            (src/pasta/pasta_3_seal.h:9-54)
//     sealhelper::packed_enc_multiply ->  sealhelper_b200::packed_enc_multiply   (src/util/sealhelper.h:84-87)
//     sealhelper::encrypted_vec_sum   ->  sealhelper_b200::encrypted_vec_sum     (src/util/sealhelper.h:125-129)
// and links libhhe_b200.so; call sites in src/examples/CSP/CSP.cpp:235-323 compile unchanged.
// Needs the SEAL 4.0 headers of the reference (libs/seal/include/SEAL-4.0); it only uses SEAL types as containers
// (Ciphertext::data()/resize, KSwitchKeys::data()) -- every arithmetic instruction runs on the GPU.
//
// Error behaviour mirrors the reference: std::invalid_argument (SEAL: missing Galois key, bad sizes),
// std::logic_error (transparent result), std::runtime_error ("too little slots for matmul implementation!", CUDA).
#pragma once
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "seal/seal.h"

#include "hhe_b200.h"

namespace hhe_shim {

inline void check(int rc) {
  if (rc == HHE_OK) return;
  const std::string msg = hhe_last_error();
  if (rc == HHE_ERR_INVALID) throw std::invalid_argument(msg);
  if (rc == HHE_ERR_LOGIC) throw std::logic_error(msg);
  throw std::runtime_error(msg);
}

// One engine context per SEALContext (lives as long as the shim objects that share it).
class Engine {
 public:
  explicit Engine(const seal::SEALContext &context, int device = 0) : context_(context) {
    auto &kp = context.key_context_data()->parms();
    std::vector<uint64_t> q;
    for (auto &m : kp.coeff_modulus()) q.push_back(m.value());
    check(hhe_ctx_create(&ctx_, kp.poly_modulus_degree(), kp.plain_modulus().value(), q.data(), static_cast<int>(q.size()), device, nullptr));
    N_ = kp.poly_modulus_degree();
    L_ = q.size() - 1;
  }
  ~Engine() { hhe_ctx_destroy(ctx_); }
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;

  // The engine shared by every shim object built on the same encryption parameters and device (pasta_b200::PASTA_SEAL,
  // hhe_shim::Evaluator): keys uploaded through one of them are visible to the others, as with one seal::SEALContext.
  static std::shared_ptr<Engine> shared(const seal::SEALContext &context, int device = 0) {
    static std::mutex mu;
    static std::map<std::pair<seal::parms_id_type, int>, std::weak_ptr<Engine>> registry;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(context.key_parms_id(), device);
    if (auto e = registry[key].lock()) return e;
    auto e = std::make_shared<Engine>(context, device);
    registry[key] = e;
    return e;
  }

  hhe_ctx *ctx() const { return ctx_; }
  size_t ct_words(size_t size = 2) const { return size * L_ * N_; }
  const seal::SEALContext &context() const { return context_; }

  // seal::KSwitchKeys (GaloisKeys / RelinKeys) -> engine keyset. ksk layout [digit][2][K][N] = back-to-back PublicKeys.
  // The object replaces whatever the keyset held before (a key it lacks is then missing, as in SEAL).
  void load(const seal::KSwitchKeys &keys, int kind) {
    check(hhe_clear_keyset(ctx_, kind));
    has_[kind] = false;
    const auto &data = keys.data();
    for (size_t index = 0; index < data.size(); ++index) {
      if (data[index].empty()) continue;
      std::vector<uint64_t> flat;
      for (auto &pk : data[index]) flat.insert(flat.end(), pk.data().data(), pk.data().data() + pk.data().dyn_array().size());
      const uint32_t elt = kind == HHE_RELIN ? 0u : static_cast<uint32_t>(2 * index + 1);  // GaloisKeys::get_index inverse
      check(hhe_load_ksk(ctx_, kind, elt, flat.data()));
    }
    fp_[kind] = fingerprint(keys);
    has_[kind] = true;
  }

  // Upload only when the keyset does not already hold these keys (the 26-key default set is ~0.5 GB of PCIe traffic).
  // Identity is by CONTENT: the reference hands its GaloisKeys / RelinKeys around by value (CSP.h:74,121 return copies, the
  // PASTA_SEAL ctor takes copies), so addresses say nothing -- a later temporary with other keys can sit at the same address.
  // The fingerprint covers parms_id, which indices are present and 16 words spread over every key polynomial pair.
  void ensure_loaded(const seal::KSwitchKeys &keys, int kind) {
    if (!has_[kind] || fp_[kind] != fingerprint(keys)) load(keys, kind);
  }

  // The same keys as they arrive over gRPC (GaloisKeys/RelinKeys::save bytes, src/examples/Analyst/Analyst.cpp:273-318): parsed by
  // the engine's SEAL wire-format codec and uploaded key by key, without materialising a seal::GaloisKeys on the host.
  size_t load_serialized(const std::string &bytes, int kind) {
    check(hhe_clear_keyset(ctx_, kind));
    has_[kind] = false;
    size_t n = 0;
    check(hhe_load_seal_keys(ctx_, kind, reinterpret_cast<const uint8_t *>(bytes.data()), bytes.size(), &n, nullptr));
    return n;
  }

  seal::Ciphertext wrap(const uint64_t *words, size_t size = 2) const {
    seal::Ciphertext ct(context_);
    ct.resize(context_, context_.first_parms_id(), size);
    std::memcpy(ct.data(), words, sizeof(uint64_t) * ct_words(size));
    ct.is_ntt_form() = false;
    return ct;
  }
  void require_fresh_level(const seal::Ciphertext &ct, size_t size) const {
    if (ct.parms_id() != context_.first_parms_id() || ct.is_ntt_form() || ct.size() != size)
      throw std::invalid_argument("encrypted is not valid for encryption parameters");
  }

  static uint64_t fingerprint(const seal::KSwitchKeys &keys) {
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mix = [&h](uint64_t v) {
      for (int i = 0; i < 8; ++i) {
        h ^= (v >> (8 * i)) & 0xff;
        h *= 0x100000001b3ULL;
      }
    };
    for (uint64_t w : keys.parms_id()) mix(w);
    const auto &data = keys.data();
    mix(data.size());
    for (size_t index = 0; index < data.size(); ++index) {
      if (data[index].empty()) continue;
      mix(index);
      mix(data[index].size());
      for (auto &pk : data[index]) {
        const uint64_t *w = pk.data().data();
        const size_t n = pk.data().dyn_array().size();
        mix(n);
        for (size_t s = 0; s < 16 && n; ++s) mix(w[(n - 1) * s / 15]);
      }
    }
    return h;
  }

 private:
  seal::SEALContext context_;  // by value (a SEALContext is a handle on shared context data): no reference to outlive
  hhe_ctx *ctx_ = nullptr;
  size_t N_ = 0, L_ = 0;
  uint64_t fp_[3] = {0, 0, 0};
  bool has_[3] = {false, false, false};
};

// seal::Evaluator-shaped facade over the engine: the members the reference's path calls on its Evaluator
// (src/examples/CSP/CSP.cpp:298,306,314; src/util/sealhelper.cpp:268-274,379-392; src/examples/hhe_pktnn_examples.cpp:653-673,
// 957-992) with seal::Evaluator's signatures, so a call site keeps its text when the object's type is switched:
//     Evaluator *csp_he_eval = new Evaluator(*context);                    // CSP.cpp:19, with `using hhe_shim::Evaluator;`
//     getEvaluator()->relinearize_inplace(record, getCSPHERelinKeysMapValue(analystId));                      // CSP.cpp:306
//     sealhelper::packed_enc_multiply(record, w, tmp, *getEvaluator());   // CSP.cpp:295-298, with sealhelper = sealhelper_b200
// Every call runs on the GPU through the C ABI; key objects are uploaded on first use and recognised by content afterwards.
class Evaluator {
 public:
  explicit Evaluator(const seal::SEALContext &context, int device = 0) : engine_(Engine::shared(context, device)) {}
  explicit Evaluator(std::shared_ptr<Engine> engine) : engine_(std::move(engine)) {}
  Engine &engine() const { return *engine_; }

  void multiply(const seal::Ciphertext &a, const seal::Ciphertext &b, seal::Ciphertext &destination) const {
    engine_->require_fresh_level(a, 2);
    engine_->require_fresh_level(b, 2);
    std::vector<uint64_t> out(engine_->ct_words(3));
    check(hhe_multiply(engine_->ctx(), a.data(), b.data(), out.data(), 1));
    destination = engine_->wrap(out.data(), 3);
  }
  void multiply_inplace(seal::Ciphertext &a, const seal::Ciphertext &b) const { multiply(a, b, a); }
  void square(const seal::Ciphertext &a, seal::Ciphertext &destination) const { multiply(a, a, destination); }
  void square_inplace(seal::Ciphertext &a) const { multiply(a, a, a); }
  void relinearize(const seal::Ciphertext &a, const seal::RelinKeys &rk, seal::Ciphertext &destination) const {
    if (a.size() == 2) {  // SEAL: nothing to do
      destination = a;
      return;
    }
    engine_->require_fresh_level(a, 3);
    engine_->ensure_loaded(rk, HHE_RELIN);
    std::vector<uint64_t> out(engine_->ct_words());
    check(hhe_relinearize(engine_->ctx(), a.data(), out.data(), 1));
    destination = engine_->wrap(out.data());
  }
  void relinearize_inplace(seal::Ciphertext &a, const seal::RelinKeys &rk) const { relinearize(a, rk, a); }
  void rotate_rows(const seal::Ciphertext &a, int steps, const seal::GaloisKeys &gk, seal::Ciphertext &destination) const {
    engine_->require_fresh_level(a, 2);
    engine_->ensure_loaded(gk, HHE_KEYSET_1);
    std::vector<uint64_t> out(engine_->ct_words());
    check(hhe_rotate_rows(engine_->ctx(), a.data(), steps, HHE_KEYSET_1, out.data(), 1));
    destination = engine_->wrap(out.data());
  }
  void rotate_rows_inplace(seal::Ciphertext &a, int steps, const seal::GaloisKeys &gk) const { rotate_rows(a, steps, gk, a); }
  void rotate_columns(const seal::Ciphertext &a, const seal::GaloisKeys &gk, seal::Ciphertext &destination) const {
    engine_->require_fresh_level(a, 2);
    engine_->ensure_loaded(gk, HHE_KEYSET_1);
    std::vector<uint64_t> out(engine_->ct_words());
    check(hhe_rotate_columns(engine_->ctx(), a.data(), HHE_KEYSET_1, out.data(), 1));
    destination = engine_->wrap(out.data());
  }
  void rotate_columns_inplace(seal::Ciphertext &a, const seal::GaloisKeys &gk) const { rotate_columns(a, gk, a); }
  void add(const seal::Ciphertext &a, const seal::Ciphertext &b, seal::Ciphertext &destination) const {
    engine_->require_fresh_level(a, 2);
    engine_->require_fresh_level(b, 2);
    std::vector<uint64_t> out(engine_->ct_words());
    check(hhe_add(engine_->ctx(), a.data(), b.data(), out.data(), 1));
    destination = engine_->wrap(out.data());
  }
  void add_inplace(seal::Ciphertext &a, const seal::Ciphertext &b) const { add(a, b, a); }
  void negate_inplace(seal::Ciphertext &a) const {
    engine_->require_fresh_level(a, 2);
    std::vector<uint64_t> out(engine_->ct_words());
    check(hhe_negate(engine_->ctx(), a.data(), out.data(), 1));
    a = engine_->wrap(out.data());
  }
  void add_plain_inplace(seal::Ciphertext &a, const seal::Plaintext &p) const {
    engine_->require_fresh_level(a, 2);
    std::vector<uint64_t> pt = dense(p), out(engine_->ct_words());
    check(hhe_add_plain(engine_->ctx(), a.data(), pt.data(), out.data(), 1));
    a = engine_->wrap(out.data());
  }
  void multiply_plain(const seal::Ciphertext &a, const seal::Plaintext &p, seal::Ciphertext &destination) const {
    engine_->require_fresh_level(a, 2);
    std::vector<uint64_t> pt = dense(p), out(engine_->ct_words());
    check(hhe_multiply_plain(engine_->ctx(), a.data(), pt.data(), out.data(), 1));
    destination = engine_->wrap(out.data());
  }
  void multiply_plain_inplace(seal::Ciphertext &a, const seal::Plaintext &p) const { multiply_plain(a, p, a); }
  void exponentiate_inplace(seal::Ciphertext &a, uint64_t exponent, const seal::RelinKeys &rk) const {
    if (exponent != 3) throw std::invalid_argument("hhe_shim::Evaluator::exponentiate_inplace: only the exponent the path uses (3) is built");
    engine_->require_fresh_level(a, 2);
    engine_->ensure_loaded(rk, HHE_RELIN);
    std::vector<uint64_t> out(engine_->ct_words());
    check(hhe_exponentiate3(engine_->ctx(), a.data(), out.data(), 1));
    a = engine_->wrap(out.data());
  }

 private:
  // seal::Plaintext (coeff_count <= N coefficients) -> the dense u64[N] the C ABI takes
  std::vector<uint64_t> dense(const seal::Plaintext &p) const {
    const size_t n = engine_->context().key_context_data()->parms().poly_modulus_degree();
    std::vector<uint64_t> pt(n, 0);
    if (p.is_ntt_form() || p.coeff_count() > n) throw std::invalid_argument("plain is not valid for encryption parameters");
    std::memcpy(pt.data(), p.data(), sizeof(uint64_t) * p.coeff_count());
    return pt;
  }
  std::shared_ptr<Engine> engine_;
};

}  // namespace hhe_shim

namespace pasta_b200 {

// pasta::PASTA_SEAL (src/pasta/pasta_3_seal.h:9-54) + the SEALZpCipher members its callers use (SEAL_Cipher.h:79,87-88)
class PASTA_SEAL {
 public:
  PASTA_SEAL(std::shared_ptr<seal::SEALContext> con, seal::PublicKey, seal::SecretKey, seal::RelinKeys rk, seal::GaloisKeys gk,
             int device = 0)
      : context_(con), engine_(hhe_shim::Engine::shared(*con, device)) {
    engine_->load(rk, HHE_RELIN);
    engine_->load(gk, HHE_KEYSET_0);
  }
  virtual ~PASTA_SEAL() = default;
  virtual std::string get_cipher_name() const { return "PASTA-SEAL (n=128,r=3) [B200]"; }
  size_t get_plain_size() const { return 128; }
  size_t get_cipher_size() const { return 128; }
  size_t get_key_size() const { return 256; }
  void activate_bsgs(bool activate) { use_bsgs_ = activate; }
  std::shared_ptr<hhe_shim::Engine> engine() const { return engine_; }

  // src/pasta/pasta_3_seal.cpp:106-172 (nonce 123456789, counters restart at 0 for every call; batch_encoder ignored)
  virtual std::vector<seal::Ciphertext> decomposition(std::vector<uint64_t> &ciphertext, std::vector<seal::Ciphertext> enc_ssk,
                                                      bool batch_encoder = false) {
    (void)batch_encoder;
    if (enc_ssk.empty()) throw std::invalid_argument("encrypted symmetric key missing");
    engine_->require_fresh_level(enc_ssk[0], 2);
    const size_t blocks = (ciphertext.size() + 127) / 128;
    std::vector<uint64_t> out(blocks * engine_->ct_words());
    hhe_shim::check(hhe_pasta3_decompose(engine_->ctx(), enc_ssk[0].data(), ciphertext.data(), ciphertext.size(), 123456789ULL, 0,
                                         use_bsgs_ ? 1 : 0, out.data()));
    std::vector<seal::Ciphertext> res;
    for (size_t b = 0; b < blocks; ++b) res.push_back(engine_->wrap(out.data() + b * engine_->ct_words()));
    return res;
  }
  // src/pasta/pasta_3_seal.cpp:42-104: same computation on the member key set by set_encrypted_key
  void set_encrypted_key(const seal::Ciphertext &k) { secret_key_encrypted_ = {k}; }
  virtual std::vector<seal::Ciphertext> HE_decrypt(std::vector<uint64_t> &ciphertext, bool batch_encoder = false) {
    return decomposition(ciphertext, secret_key_encrypted_, batch_encoder);
  }
  // src/pasta/SEAL_Cipher.cpp:161-166
  void mask(seal::Ciphertext &cipher, std::vector<uint64_t> &mask) {
    engine_->require_fresh_level(cipher, 2);
    std::vector<uint64_t> out(engine_->ct_words());
    hhe_shim::check(hhe_mask(engine_->ctx(), cipher.data(), mask.data(), mask.size(), out.data(), 1));
    cipher = engine_->wrap(out.data());
  }
  // src/pasta/SEAL_Cipher.cpp:170-181 (the keys passed here go to keyset 1 so the PASTA keys stay loaded)
  void flatten(std::vector<seal::Ciphertext> &in, seal::Ciphertext &out, const seal::GaloisKeys &galois_keys) {
    if (in.empty()) throw std::invalid_argument("flatten: empty input");
    engine_->ensure_loaded(galois_keys, HHE_KEYSET_1);
    std::vector<uint64_t> flat(in.size() * engine_->ct_words()), res(engine_->ct_words());
    for (size_t i = 0; i < in.size(); ++i) {
      engine_->require_fresh_level(in[i], 2);
      std::memcpy(flat.data() + i * engine_->ct_words(), in[i].data(), sizeof(uint64_t) * engine_->ct_words());
    }
    hhe_shim::check(hhe_flatten(engine_->ctx(), flat.data(), in.size(), HHE_KEYSET_1, res.data(), 1));
    out = engine_->wrap(res.data());
  }

 private:
  std::shared_ptr<seal::SEALContext> context_;
  std::shared_ptr<hhe_shim::Engine> engine_;
  std::vector<seal::Ciphertext> secret_key_encrypted_;
  bool use_bsgs_ = false;
};

}  // namespace pasta_b200

namespace sealhelper_b200 {

// The reference passes a seal::Evaluator; the drop-in passes the engine that replaces it.
inline void packed_enc_multiply(const seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2, seal::Ciphertext &destination,
                                const hhe_shim::Engine &engine) {
  engine.require_fresh_level(encrypted1, 2);
  engine.require_fresh_level(encrypted2, 2);
  std::vector<uint64_t> out(engine.ct_words(3));
  hhe_shim::check(hhe_multiply(engine.ctx(), encrypted1.data(), encrypted2.data(), out.data(), 1));
  destination = engine.wrap(out.data(), 3);
}

// Evaluator::relinearize_inplace(ct, rk) at src/examples/CSP/CSP.cpp:306 (relin key loaded with Engine::load(rk, HHE_RELIN))
inline void relinearize_inplace(seal::Ciphertext &encrypted, const hhe_shim::Engine &engine) {
  engine.require_fresh_level(encrypted, 3);
  std::vector<uint64_t> out(engine.ct_words());
  hhe_shim::check(hhe_relinearize(engine.ctx(), encrypted.data(), out.data(), 1));
  encrypted = engine.wrap(out.data());
}

inline void encrypted_vec_sum(const seal::Ciphertext &encrypted_inp, seal::Ciphertext &destination, hhe_shim::Engine &engine,
                              const seal::GaloisKeys &gal_keys, const size_t vec_size);

// The reference's own signatures (src/util/sealhelper.h:84-87,125-129) with the facade in the Evaluator's place: the call sites
// CSP.cpp:295-298,311-315 and hhe_pktnn_examples.cpp:653-673,957-992 compile as they are written.
inline void packed_enc_multiply(const seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2, seal::Ciphertext &destination,
                                const hhe_shim::Evaluator &evaluator) {
  packed_enc_multiply(encrypted1, encrypted2, destination, static_cast<const hhe_shim::Engine &>(evaluator.engine()));
}
inline void encrypted_vec_sum(const seal::Ciphertext &encrypted_inp, seal::Ciphertext &destination, const hhe_shim::Evaluator &evaluator,
                              const seal::GaloisKeys &gal_keys, const size_t vec_size) {
  encrypted_vec_sum(encrypted_inp, destination, evaluator.engine(), gal_keys, vec_size);
}

inline void encrypted_vec_sum(const seal::Ciphertext &encrypted_inp, seal::Ciphertext &destination, hhe_shim::Engine &engine,
                              const seal::GaloisKeys &gal_keys, const size_t vec_size) {
  engine.require_fresh_level(encrypted_inp, 2);
  engine.ensure_loaded(gal_keys, HHE_KEYSET_1);
  std::vector<uint64_t> out(engine.ct_words());
  hhe_shim::check(hhe_vec_sum(engine.ctx(), encrypted_inp.data(), vec_size, HHE_KEYSET_1, out.data(), 1));
  destination = engine.wrap(out.data());
}

}  // namespace sealhelper_b200

// Service-level mirrors of the CSP request handlers (src/examples/CSP/CSP.cpp): one engine call per request, the
// intermediate ciphertexts stay in HBM. The keys are uploaded once per analyst (Engine::load), not per call as the
// reference's per-call PASTA_SEAL construction does (CSP.cpp:238-242).
namespace csp_b200 {

// BaseCSP::decompose (CSP.cpp:235-283): records -> one flattened ciphertext per record. `apply_mask` = false keeps the
// reference service's behaviour (its mask acts on a copy, CSP.cpp:262-269).
inline std::vector<seal::Ciphertext> decompose(hhe_shim::Engine &engine, const std::vector<std::vector<uint64_t>> &records,
                                               const seal::Ciphertext &enc_sym_key, const seal::GaloisKeys &csp_gk, bool use_bsgs = false,
                                               bool apply_mask = false) {
  if (records.empty()) return {};
  engine.require_fresh_level(enc_sym_key, 2);
  const size_t n = records[0].size();
  std::vector<uint64_t> flat_in(records.size() * n), out(records.size() * engine.ct_words());
  for (size_t r = 0; r < records.size(); ++r) {
    if (records[r].size() != n) throw std::invalid_argument("records must have equal length");
    std::memcpy(flat_in.data() + r * n, records[r].data(), sizeof(uint64_t) * n);
  }
  if (n > 128) engine.ensure_loaded(csp_gk, HHE_KEYSET_1);
  hhe_shim::check(hhe_csp_decompose(engine.ctx(), enc_sym_key.data(), flat_in.data(), n, records.size(), 123456789ULL, use_bsgs ? 1 : 0,
                                    apply_mask ? 1 : 0, HHE_KEYSET_1, out.data()));
  std::vector<seal::Ciphertext> res;
  for (size_t r = 0; r < records.size(); ++r) res.push_back(engine.wrap(out.data() + r * engine.ct_words()));
  return res;
}

// CSP_hhe_pktnn_1fc::evaluateModel (CSP.cpp:288-323): result[record][row]
inline std::vector<std::vector<seal::Ciphertext>> evaluate_model(hhe_shim::Engine &engine, const std::vector<seal::Ciphertext> &records,
                                                                 const std::vector<seal::Ciphertext> &enc_weights,
                                                                 const seal::GaloisKeys &analyst_gk, size_t input_len) {
  const size_t ctw = engine.ct_words();
  std::vector<uint64_t> x(records.size() * ctw), w(enc_weights.size() * ctw), out(records.size() * enc_weights.size() * ctw);
  for (size_t i = 0; i < records.size(); ++i) {
    engine.require_fresh_level(records[i], 2);
    std::memcpy(x.data() + i * ctw, records[i].data(), sizeof(uint64_t) * ctw);
  }
  for (size_t i = 0; i < enc_weights.size(); ++i) {
    engine.require_fresh_level(enc_weights[i], 2);
    std::memcpy(w.data() + i * ctw, enc_weights[i].data(), sizeof(uint64_t) * ctw);
  }
  engine.ensure_loaded(analyst_gk, HHE_KEYSET_1);
  hhe_shim::check(hhe_csp_evaluate_model(engine.ctx(), x.data(), records.size(), w.data(), enc_weights.size(), input_len, HHE_KEYSET_1,
                                         out.data()));
  std::vector<std::vector<seal::Ciphertext>> res(records.size());
  for (size_t r = 0; r < records.size(); ++r)
    for (size_t k = 0; k < enc_weights.size(); ++k) res[r].push_back(engine.wrap(out.data() + (r * enc_weights.size() + k) * ctw));
  return res;
}

}  // namespace csp_b200
