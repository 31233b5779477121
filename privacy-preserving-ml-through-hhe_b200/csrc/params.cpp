#include "params.h"

#include <stdexcept>
#include <string>

namespace hhe {

u64 mul_mod(u64 a, u64 b, u64 q) { return static_cast<u64>(static_cast<u128>(a) * b % q); }

u64 pow_mod(u64 a, u64 e, u64 q) {
  u64 r = 1 % q;
  a %= q;
  for (; e; e >>= 1) {
    if (e & 1) r = mul_mod(r, a, q);
    a = mul_mod(a, a, q);
  }
  return r;
}

u64 inv_mod_prime(u64 a, u64 q) { return pow_mod(a % q, q - 2, q); }

Twiddle shoup_pair(u64 w, u64 q) { return Twiddle{w, static_cast<u64>((static_cast<u128>(w) << 64) / q)}; }

namespace {

bool miller_rabin(u64 n) {
  if (n < 4) return n == 2 || n == 3;
  if (!(n & 1)) return false;
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) d >>= 1, ++s;
  // deterministic for 64-bit inputs
  for (u64 a : {2ULL, 325ULL, 9375ULL, 28178ULL, 450775ULL, 9780504ULL, 1795265022ULL}) {
    u64 x = pow_mod(a % n, d, n);
    if (x == 0 || x == 1 || x == n - 1) continue;
    bool witness = true;
    for (int r = 1; r < s && witness; ++r) {
      x = mul_mod(x, x, n);
      if (x == n - 1) witness = false;
    }
    if (witness) return false;
  }
  return true;
}

uint32_t reverse_bits(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i, x >>= 1) r = (r << 1) | (x & 1);
  return r;
}

// Smallest primitive 2N-th root of unity modulo the prime q.
u64 smallest_2n_root(u64 N, u64 q) {
  if ((q - 1) % (2 * N)) throw std::invalid_argument("modulus " + std::to_string(q) + " is not 1 mod 2N");
  u64 root = 0;
  for (u64 g = 2; !root; ++g) {
    u64 cand = pow_mod(g, (q - 1) / (2 * N), q);
    if (pow_mod(cand, N, q) == q - 1) root = cand;
  }
  // all primitive 2N-th roots are the odd powers of any one of them
  u64 step = mul_mod(root, root, q), cur = root, best = root;
  for (u64 i = 1; i < N; ++i) {
    cur = mul_mod(cur, step, q);
    if (cur < best) best = cur;
  }
  return best;
}

NttTable make_table(u64 N, int logn, u64 q) {
  NttTable t;
  t.q = q;
  t.psi = smallest_2n_root(N, q);
  t.fwd.resize(N);
  t.inv.resize(N);
  u64 psi_inv = inv_mod_prime(t.psi, q), a = 1, b = 1;
  for (u64 i = 0; i < N; ++i) {
    uint32_t k = reverse_bits(static_cast<uint32_t>(i), logn);
    t.fwd[k] = shoup_pair(a, q);
    t.inv[k] = shoup_pair(b, q);
    a = mul_mod(a, t.psi, q);
    b = mul_mod(b, psi_inv, q);
  }
  t.n_inv = shoup_pair(inv_mod_prime(N % q, q), q);
  return t;
}

// product of base[] except index `skip` (skip < 0: all), reduced mod m (m need not be prime)
u64 punctured(const u64 *base, int n, int skip, u64 m) {
  u64 r = 1 % m;
  for (int i = 0; i < n; ++i)
    if (i != skip) r = mul_mod(r, base[i] % m, m);
  return r;
}

}  // namespace

std::vector<int> naf_steps(int value) {
  std::vector<int> out;
  bool neg = value < 0;
  unsigned v = neg ? static_cast<unsigned>(-value) : static_cast<unsigned>(value);
  for (int bit = 0; v; ++bit) {
    int digit = 0;
    if (v & 1) digit = 2 - static_cast<int>(v & 3);
    v = (v - digit) >> 1;
    if (digit) out.push_back((neg ? -digit : digit) * (1 << bit));
  }
  return out;
}

uint32_t Params::galois_elt_from_step(int step) const {
  const u64 m = 2 * N;
  if (step == 0) return static_cast<uint32_t>(m - 1);
  u64 mag = step < 0 ? static_cast<u64>(-static_cast<int64_t>(step)) : static_cast<u64>(step);
  if (mag >= N / 2) return 0;
  u64 exponent = step < 0 ? N / 2 - mag : mag;
  return static_cast<uint32_t>(pow_mod(3, exponent, m));
}

Params Params::derive(u64 N, u64 t, const u64 *q, int nq) {
  if (N < 256 || N > 32768 || (N & (N - 1))) throw std::invalid_argument("poly_modulus_degree must be a power of two in [256, 32768]");
  if (nq < 2 || nq > kMaxLimbs - 2) throw std::invalid_argument("need 2..16 coefficient primes (data primes + special prime)");
  if (t < 2 || t >= (1ULL << 32)) throw std::invalid_argument("plain modulus out of range");
  Params p;
  p.N = N;
  p.t = t;
  p.K = nq;
  p.L = nq - 1;
  while ((1ULL << p.logn) < N) ++p.logn;
  p.q.assign(q, q + nq);
  for (int i = 0; i < nq; ++i) {
    if (q[i] >> 61) throw std::invalid_argument("coefficient primes must be below 2^61");
    if (!miller_rabin(q[i])) throw std::invalid_argument("coefficient modulus is not prime");
    for (int j = 0; j < i; ++j)
      if (q[i] == q[j]) throw std::invalid_argument("coefficient primes must be distinct");
  }
  if (!miller_rabin(t) || (t - 1) % (2 * N)) throw std::invalid_argument("plain modulus must be a prime = 1 mod 2N (batching)");
  const int L = p.L, K = p.K;

  // auxiliary BEHZ primes: 61-bit, = 1 mod 2N, scanned downwards; first two are m_sk and gamma
  std::vector<u64> aux;
  for (u64 v = ((1ULL << 61) - 1) / (2 * N) * (2 * N) + 1; static_cast<int>(aux.size()) < L + 2; v -= 2 * N)
    if (miller_rabin(v)) aux.push_back(v);
  p.m_sk = aux[0];
  p.gamma = aux[1];
  p.m_tilde = 1ULL << 32;
  for (int i = 0; i < L; ++i) p.bsk[i] = aux[2 + i];
  p.bsk[L] = p.m_sk;

  p.tab.reserve(2 * K + 1);
  for (int i = 0; i < K; ++i) p.tab.push_back(make_table(N, p.logn, q[i]));
  for (int i = 0; i <= L; ++i) p.tab.push_back(make_table(N, p.logn, p.bsk[i]));
  while (static_cast<int>(p.tab.size()) < 2 * K) p.tab.push_back(NttTable{});
  p.tab.push_back(make_table(N, p.logn, t));

  p.index_map.resize(N);
  u64 pos = 1;
  for (u64 i = 0; i < N / 2; ++i) {
    p.index_map[i] = reverse_bits(static_cast<uint32_t>((pos - 1) >> 1), p.logn);
    p.index_map[N / 2 + i] = reverse_bits(static_cast<uint32_t>((2 * N - pos - 1) >> 1), p.logn);
    pos = pos * 3 % (2 * N);
  }

  // Q = q_0 * ... * q_{L-1} as little-endian words; floor(Q / t), Q mod t
  std::vector<u64> big{1};
  for (int i = 0; i < L; ++i) {
    u64 carry = 0;
    for (auto &w : big) {
      u128 prod = static_cast<u128>(w) * q[i] + carry;
      w = static_cast<u64>(prod);
      carry = static_cast<u64>(prod >> 64);
    }
    if (carry) big.push_back(carry);
  }
  std::vector<u64> quot(big.size());
  u64 rem = 0;
  for (size_t w = big.size(); w-- > 0;) {
    u128 cur = (static_cast<u128>(rem) << 64) | big[w];
    quot[w] = static_cast<u64>(cur / t);
    rem = static_cast<u64>(cur % t);
  }
  p.q_mod_t = rem;
  p.half_t = (t + 1) >> 1;
  for (int j = 0; j < L; ++j) {
    u64 r = 0;
    for (size_t w = quot.size(); w-- > 0;) r = static_cast<u64>(((static_cast<u128>(r) << 64) | quot[w]) % q[j]);
    p.q_div_t_mod_q[j] = r;
  }

  const u64 q_sp = q[K - 1];
  p.half_sp = q_sp >> 1;
  for (int i = 0; i < L; ++i) {
    p.half_sp_mod_q[i] = p.half_sp % q[i];
    p.inv_sp_mod_q[i] = shoup_pair(inv_mod_prime(q_sp % q[i], q[i]), q[i]);
  }

  const u64 *qd = p.q.data();
  for (int i = 0; i < L; ++i) {
    u64 ip = inv_mod_prime(punctured(qd, L, i, q[i]), q[i]);
    p.inv_punct_q[i] = shoup_pair(ip, q[i]);
    p.mtilde_mod_q[i] = shoup_pair(mul_mod(p.m_tilde % q[i], ip, q[i]), q[i]);
    p.t_inv_punct_q[i] = shoup_pair(mul_mod(t % q[i], ip, q[i]), q[i]);
    for (int b = 0; b <= L; ++b) p.q2bsk[b][i] = punctured(qd, L, i, p.bsk[b]);
    p.q2mt[i] = static_cast<uint32_t>(punctured(qd, L, i, p.m_tilde));
    p.inv_punct_b[i] = shoup_pair(inv_mod_prime(punctured(p.bsk, L, i, p.bsk[i]), p.bsk[i]), p.bsk[i]);
    for (int j = 0; j < L; ++j) p.b2q[j][i] = punctured(p.bsk, L, i, q[j]);
    p.b2msk[i] = punctured(p.bsk, L, i, p.m_sk);
    p.pb_mod_q[i] = punctured(p.bsk, L, -1, q[i]);
  }
  uint32_t q32 = static_cast<uint32_t>(punctured(qd, L, -1, p.m_tilde)), inv = 1;
  for (int it = 0; it < 5; ++it) inv *= 2u - q32 * inv;  // Newton: inverse of an odd number mod 2^32
  p.neg_inv_q_mt = 0u - inv;
  for (int b = 0; b <= L; ++b) {
    u64 m = p.bsk[b];
    p.q_mod_bsk[b] = punctured(qd, L, -1, m);
    p.inv_mt_bsk[b] = shoup_pair(inv_mod_prime(p.m_tilde % m, m), m);
    p.t_mod_bsk[b] = shoup_pair(t % m, m);
    p.inv_q_bsk[b] = shoup_pair(inv_mod_prime(p.q_mod_bsk[b], m), m);
  }
  p.inv_pb_msk = shoup_pair(inv_mod_prime(punctured(p.bsk, L, -1, p.m_sk), p.m_sk), p.m_sk);
  return p;
}

}  // namespace hhe
