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
#include <memory>
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

  hhe_ctx *ctx() const { return ctx_; }
  size_t ct_words(size_t size = 2) const { return size * L_ * N_; }

  // seal::KSwitchKeys (GaloisKeys / RelinKeys) -> engine keyset. ksk layout [digit][2][K][N] = back-to-back PublicKeys.
  // The object replaces whatever the keyset held before (a key it lacks is then missing, as in SEAL).
  void load(const seal::KSwitchKeys &keys, int kind) {
    check(hhe_clear_keyset(ctx_, kind));
    loaded_[kind] = &keys;
    const auto &data = keys.data();
    for (size_t index = 0; index < data.size(); ++index) {
      if (data[index].empty()) continue;
      std::vector<uint64_t> flat;
      for (auto &pk : data[index]) flat.insert(flat.end(), pk.data().data(), pk.data().data() + pk.data().dyn_array().size());
      const uint32_t elt = kind == HHE_RELIN ? 0u : static_cast<uint32_t>(2 * index + 1);  // GaloisKeys::get_index inverse
      check(hhe_load_ksk(ctx_, kind, elt, flat.data()));
    }
  }

  // Upload only when a different key object is asked for than the one this keyset holds (the 26-key default set is ~0.5 GB
  // of PCIe traffic). Identity is by address: callers keep their seal::GaloisKeys alive and unmodified, as the reference does.
  void ensure_loaded(const seal::KSwitchKeys &keys, int kind) {
    if (loaded_[kind] != &keys) load(keys, kind);
  }

  // The same keys as they arrive over gRPC (GaloisKeys/RelinKeys::save bytes, src/examples/Analyst/Analyst.cpp:273-318): parsed by
  // the engine's SEAL wire-format codec and uploaded key by key, without materialising a seal::GaloisKeys on the host.
  size_t load_serialized(const std::string &bytes, int kind) {
    check(hhe_clear_keyset(ctx_, kind));
    loaded_[kind] = nullptr;
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

 private:
  const seal::SEALContext &context_;
  hhe_ctx *ctx_ = nullptr;
  size_t N_ = 0, L_ = 0;
  const seal::KSwitchKeys *loaded_[3] = {nullptr, nullptr, nullptr};
};

}  // namespace hhe_shim

namespace pasta_b200 {

// pasta::PASTA_SEAL (src/pasta/pasta_3_seal.h:9-54) + the SEALZpCipher members its callers use (SEAL_Cipher.h:79,87-88)
class PASTA_SEAL {
 public:
  PASTA_SEAL(std::shared_ptr<seal::SEALContext> con, seal::PublicKey, seal::SecretKey, seal::RelinKeys rk, seal::GaloisKeys gk,
             int device = 0)
      : context_(con), engine_(std::make_shared<hhe_shim::Engine>(*con, device)) {
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
