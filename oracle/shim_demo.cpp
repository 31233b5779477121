// oracle/shim_demo.cpp -- TEST INFRASTRUCTURE. Drop-in proof: the same call sequence as BaseCSP::decompose /
// CSP_hhe_pktnn_1fc::evaluateModel (src/examples/CSP/CSP.cpp:235-323) is run twice -- once with the reference's own
// pasta::PASTA_SEAL + seal::Evaluator (CPU, libseal-4.0.a) and once with the shim classes of
// privacy-preserving-ml-through-hhe_b200/host/hhe_seal_shim.h (GPU, libhhe_b200.so) -- and every seal::Ciphertext is
// compared word for word. Built by `make -C oracle ref` (needs /root/reference); the binary travels to the GPU box.
//   usage: shim_demo [poly_modulus_degree=16384] [input_len=300] [vec_sum_len=16]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "hhe_seal_shim.h"
#include "pasta_3_plain.h"
#include "pasta_3_seal.h"
#include "sealhelper.h"

using namespace seal;

static bool same(const Ciphertext &a, const Ciphertext &b) {
  return a.size() == b.size() && a.coeff_modulus_size() == b.coeff_modulus_size() &&
         !std::memcmp(a.data(), b.data(), sizeof(uint64_t) * a.size() * a.coeff_modulus_size() * a.poly_modulus_degree());
}

int main(int argc, char **argv) {
  const size_t N = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 16384;
  const size_t input_len = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 300;
  const size_t sum_len = argc > 3 ? std::strtoul(argv[3], nullptr, 10) : 16;
  auto context = pasta::SEALZpCipher::create_context(N, 65537, 128);
  KeyGenerator keygen(*context);
  SecretKey sk = keygen.secret_key();
  PublicKey pk;
  keygen.create_public_key(pk);
  RelinKeys rk;
  keygen.create_relin_keys(rk);
  GaloisKeys pasta_gk, flat_gk, sum_gk;
  keygen.create_galois_keys(std::vector<int>{0, -1, 128}, pasta_gk);
  std::vector<int> flat_steps;
  const size_t blocks = (input_len + 127) / 128;
  for (size_t i = 1; i < blocks; i++) flat_steps.push_back(-(int)(i * 128));
  if (!flat_steps.empty()) keygen.create_galois_keys(flat_steps, flat_gk);
  std::vector<int> pow2;
  for (int s = 1; s <= (int)sum_len; s <<= 1) pow2.push_back(-s), pow2.push_back(s);  // NAF terms have both signs
  keygen.create_galois_keys(pow2, sum_gk);
  BatchEncoder benc(*context);
  Encryptor enc(*context, pk);
  Evaluator eval(*context);
  Decryptor dec(*context, sk);

  std::vector<uint64_t> key(256), plain(input_len);
  for (size_t i = 0; i < 256; i++) key[i] = (i * 40503u + 12345u) % 65537;
  for (size_t i = 0; i < input_len; i++) plain[i] = (i * 7919u + 17u) % 256;
  pasta::PASTA cipher(key, 65537);
  std::vector<uint64_t> sym = cipher.encrypt(plain);
  std::vector<uint64_t> key_slots(N / 2 + 128, 0);
  for (size_t i = 0; i < 128; i++) key_slots[i] = key[i], key_slots[N / 2 + i] = key[128 + i];
  Plaintext kp;
  benc.encode(key_slots, kp);
  std::vector<Ciphertext> enc_key(1);
  enc.encrypt(kp, enc_key[0]);
  std::vector<int64_t> w(input_len);
  for (size_t i = 0; i < input_len; i++) w[i] = (int64_t)(i % 7) - 3;
  Plaintext wp;
  benc.encode(w, wp);
  Ciphertext enc_w;
  enc.encrypt(wp, enc_w);

  // ---- reference (CPU) ----
  auto t0 = std::chrono::steady_clock::now();
  pasta::PASTA_SEAL ref(context, pk, sk, rk, pasta_gk);
  std::vector<Ciphertext> r_blocks = ref.decomposition(sym, enc_key, true);
  Ciphertext r_flat, r_prod, r_sum;
  ref.flatten(r_blocks, r_flat, flat_gk);
  sealhelper::packed_enc_multiply(r_flat, enc_w, r_prod, eval);
  eval.relinearize_inplace(r_prod, rk);
  sealhelper::encrypted_vec_sum(r_prod, r_sum, eval, sum_gk, sum_len);
  auto t1 = std::chrono::steady_clock::now();

  // ---- drop-in (GPU) ----
  pasta_b200::PASTA_SEAL gpu(context, pk, sk, rk, pasta_gk);
  std::vector<Ciphertext> g_blocks = gpu.decomposition(sym, enc_key, true);
  Ciphertext g_flat, g_prod, g_sum;
  gpu.flatten(g_blocks, g_flat, flat_gk);
  hhe_shim::Engine &engine = *gpu.engine();
  sealhelper_b200::packed_enc_multiply(g_flat, enc_w, g_prod, engine);
  sealhelper_b200::relinearize_inplace(g_prod, engine);
  sealhelper_b200::encrypted_vec_sum(g_prod, g_sum, engine, sum_gk, sum_len);
  auto t2 = std::chrono::steady_clock::now();

  // ---- the reference's own lines (CSP.cpp:295-298,306,311-315) with the Evaluator facade in the seal::Evaluator's place ----
  bool facade_ok;
  {
    namespace sealhelper = sealhelper_b200;  // what a maintainer switches: the helper namespace and the evaluator's type
    using Evaluator = hhe_shim::Evaluator;
    Evaluator *csp_he_eval = new Evaluator(*context);  // CSP.cpp:19
    auto getEvaluator = [&] { return csp_he_eval; };
    Ciphertext tmp, tmp1;
    sealhelper::packed_enc_multiply(g_flat, enc_w, tmp, *getEvaluator());
    Ciphertext record = tmp;
    getEvaluator()->relinearize_inplace(record, RelinKeys(rk));  // a temporary copy, as CSP.h:121 returns by value
    sealhelper::encrypted_vec_sum(record, tmp1, *getEvaluator(), GaloisKeys(sum_gk), sum_len);
    Ciphertext r_prod3;  // the reference's size-3 product (r_prod was relinearized in place above)
    ::sealhelper::packed_enc_multiply(r_flat, enc_w, r_prod3, eval);
    facade_ok = same(tmp, r_prod3) && same(record, r_prod) && same(tmp1, r_sum);
    // a different key object at (possibly) the same address must be recognised by content: rotate with flat_gk, then sum_gk again
    if (!flat_steps.empty()) {
      Ciphertext a, b;
      getEvaluator()->rotate_rows(g_flat, flat_steps[0], GaloisKeys(flat_gk), a);
      eval.rotate_rows(r_flat, flat_steps[0], flat_gk, b);
      facade_ok = facade_ok && same(a, b);
      sealhelper::encrypted_vec_sum(record, tmp1, *getEvaluator(), GaloisKeys(sum_gk), sum_len);
      facade_ok = facade_ok && same(tmp1, r_sum);
    }
    delete csp_he_eval;
  }

  // ---- service-level calls (csp_b200::decompose / evaluate_model = BaseCSP::decompose / evaluateModel in one call each) ----
  std::vector<Ciphertext> s_flat = csp_b200::decompose(engine, {sym}, enc_key[0], flat_gk);
  std::vector<std::vector<Ciphertext>> s_res;
  bool svc_ok = s_flat.size() == 1 && same(s_flat[0], r_flat);
  if (sum_len == input_len) {  // evaluateModel sums over the whole record
    s_res = csp_b200::evaluate_model(engine, s_flat, {enc_w}, sum_gk, input_len);
    svc_ok = svc_ok && s_res.size() == 1 && s_res[0].size() == 1 && same(s_res[0][0], r_sum);
  }

  // ---- the same keys as serialized gRPC payloads (Analyst.cpp:273-318): parsed and uploaded by the engine's SEAL codec ----
  std::stringstream rk_ss, gk_ss;
  rk.save(rk_ss);        // SEAL's default compression (zstd)
  pasta_gk.save(gk_ss);
  const size_t n_rk = engine.load_serialized(rk_ss.str(), HHE_RELIN), n_gk = engine.load_serialized(gk_ss.str(), HHE_KEYSET_0);
  std::vector<uint64_t> first_block(sym.begin(), sym.begin() + std::min<size_t>(128, sym.size()));
  std::vector<Ciphertext> w_blocks = gpu.decomposition(first_block, enc_key, true);
  const bool wire_ok = n_rk == 1 && n_gk == 3 && w_blocks.size() == 1 && same(w_blocks[0], r_blocks[0]);
  svc_ok = svc_ok && wire_ok;

  bool ok = svc_ok && facade_ok && r_blocks.size() == g_blocks.size();
  for (size_t b = 0; ok && b < r_blocks.size(); b++) ok = same(r_blocks[b], g_blocks[b]);
  ok = ok && same(r_flat, g_flat) && same(r_prod, g_prod) && same(r_sum, g_sum);
  Plaintext p;
  dec.decrypt(g_flat, p);
  std::vector<uint64_t> slots;
  benc.decode(p, slots);
  bool dec_ok = true;
  for (size_t i = 0; i < input_len; i++) dec_ok = dec_ok && slots[i] == plain[i];
  std::printf("{\"N\": %zu, \"blocks\": %zu, \"ciphertexts_identical\": %s, \"decrypts_to_plaintext\": %s, \"noise_budget\": %d, "
              "\"serialized_keys_ok\": %s, \"evaluator_facade_ok\": %s, \"reference_cpu_s\": %.3f, \"b200_s\": %.3f}\n",
              N, blocks, ok ? "true" : "false", dec_ok ? "true" : "false", dec.invariant_noise_budget(g_sum), wire_ok ? "true" : "false", facade_ok ? "true" : "false",
              std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
  return ok && dec_ok ? 0 : 1;
}
