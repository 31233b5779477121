/* oracle/hhe_oracle.c -- CPU restatement of the reference hot path. TEST INFRASTRUCTURE, NOT PRODUCT.
 * See hhe_oracle.h for scope, pinning and who may load this. Citations are relative to /root/reference.
 * SEAL 4.0.0's .cpp sources are absent from the reference tree (headers + libseal-4.0.a only); the
 * SEAL-level functions below restate its published algorithms and are checked
 * limb-for-limb against the linked reference in tests/test_oracle_vs_ref.py.
 */
#include "hhe_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define PASTA_T 128
#define MAX_K 32
#define MAX_KEYS 64

/* ------------------------------------------------------------------ modular helpers */
static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
static inline u64 addmod(u64 a, u64 b, u64 q) {
  u64 s = a + b;
  return s >= q ? s - q : s;
}
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static u64 powmod(u64 a, u64 e, u64 q) {
  u64 r = 1 % q;
  a %= q;
  while (e) {
    if (e & 1) r = mulmod(r, a, q);
    a = mulmod(a, a, q);
    e >>= 1;
  }
  return r;
}
static u64 invmod_prime(u64 a, u64 q) { return powmod(a, q - 2, q); }
/* Shoup multiply: ws = floor(w * 2^64 / q) (MultiplyUIntModOperand, util/uintarithsmallmod.h:255-326) */
static inline u64 shoup_of(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
static inline u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) {
  u64 hi = (u64)(((u128)x * ws) >> 64);
  u64 r = x * w - hi * q;
  return r >= q ? r - q : r;
}

static int is_prime64(u64 n) {
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (size_t i = 0; i < 12; i++) {
    if (n % bases[i] == 0) return n == bases[i];
  }
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) {
    d >>= 1;
    s++;
  }
  for (size_t i = 0; i < 12; i++) {
    u64 x = powmod(bases[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int comp = 1;
    for (int r = 1; r < s; r++) {
      x = mulmod(x, x, n);
      if (x == n - 1) {
        comp = 0;
        break;
      }
    }
    if (comp) return 0;
  }
  return 1;
}

static unsigned bitrev(unsigned x, int bits) {
  unsigned r = 0;
  for (int i = 0; i < bits; i++) {
    r = (r << 1) | (x & 1);
    x >>= 1;
  }
  return r;
}

/* SEAL's NTT root: the numerically smallest primitive 2N-th root of unity mod q
 * (try_minimal_primitive_root, util/numth.h; NTTTables, util/ntt.h:69-93). */
static u64 minimal_primitive_root(u64 N, u64 q) {
  u64 two_n = 2 * N, root = 0;
  for (u64 g = 2;; g++) {
    u64 r = powmod(g, (q - 1) / two_n, q);
    if (powmod(r, N, q) == q - 1) {
      root = r;
      break;
    }
  }
  u64 sq = mulmod(root, root, q), cur = root, best = root;
  for (u64 i = 0; i < N; i++) {
    if (cur < best) best = cur;
    cur = mulmod(cur, sq, q);
  }
  return best;
}

/* ------------------------------------------------------------------ NTT tables */
typedef struct {
  u64 q, psi;
  u64 *w, *ws;   /* w[k] = psi^bitrev(k)   (SEAL root_powers layout, util/ntt.h) */
  u64 *iw, *iws; /* iw[k] = psi^-bitrev(k) */
  u64 ninv, ninvs;
} ntt_tab;

static void ntt_tab_init(ntt_tab *t, u64 N, int logn, u64 q) {
  t->q = q;
  t->psi = minimal_primitive_root(N, q);
  t->w = malloc(sizeof(u64) * N);
  t->ws = malloc(sizeof(u64) * N);
  t->iw = malloc(sizeof(u64) * N);
  t->iws = malloc(sizeof(u64) * N);
  u64 ipsi = invmod_prime(t->psi, q), p = 1, ip = 1;
  for (u64 i = 0; i < N; i++) {
    unsigned k = bitrev((unsigned)i, logn);
    t->w[k] = p;
    t->ws[k] = shoup_of(p, q);
    t->iw[k] = ip;
    t->iws[k] = shoup_of(ip, q);
    p = mulmod(p, t->psi, q);
    ip = mulmod(ip, ipsi, q);
  }
  t->ninv = invmod_prime(N % q, q);
  t->ninvs = shoup_of(t->ninv, q);
}
static void ntt_tab_free(ntt_tab *t) {
  free(t->w);
  free(t->ws);
  free(t->iw);
  free(t->iws);
}

/* Forward negacyclic NTT, natural order in, bit-reversed out: out[i] = a(psi^(2*bitrev(i)+1)).
 * Cooley-Tukey as in util/dwthandler.h:94-191 (transform_to_rev), canonical output. */
static void ntt_fwd(const ntt_tab *t, u64 N, u64 *a) {
  u64 q = t->q, gap = N;
  for (u64 m = 1; m < N; m <<= 1) {
    gap >>= 1;
    for (u64 i = 0; i < m; i++) {
      u64 w = t->w[m + i], ws = t->ws[m + i];
      u64 *x = a + 2 * i * gap, *y = x + gap;
      for (u64 j = 0; j < gap; j++) {
        u64 u = x[j], v = mul_shoup(y[j], w, ws, q);
        x[j] = addmod(u, v, q);
        y[j] = submod(u, v, q);
      }
    }
  }
}
/* Inverse: Gentleman-Sande (util/dwthandler.h:202-356, transform_from_rev), 1/N folded in, canonical. */
static void ntt_inv(const ntt_tab *t, u64 N, u64 *a) {
  u64 q = t->q, gap = 1;
  for (u64 m = N; m > 1; m >>= 1) {
    u64 h = m >> 1;
    for (u64 i = 0; i < h; i++) {
      u64 w = t->iw[h + i], ws = t->iws[h + i];
      u64 *x = a + 2 * i * gap, *y = x + gap;
      for (u64 j = 0; j < gap; j++) {
        u64 u = x[j], v = y[j];
        x[j] = addmod(u, v, q);
        y[j] = mul_shoup(submod(u, v, q), w, ws, q);
      }
    }
    gap <<= 1;
  }
  for (u64 j = 0; j < N; j++) a[j] = mul_shoup(a[j], t->ninv, t->ninvs, q);
}

/* ------------------------------------------------------------------ context */
typedef struct {
  uint32_t elt;
  u64 *data; /* [L][2][K][N] */
} ksk_t;

struct hor_ctx {
  u64 N, t;
  int logn, K, L;
  u64 q[MAX_K];
  ntt_tab *qt; /* K tables */
  ntt_tab pt;  /* mod t */
  uint32_t *index_map; /* BatchEncoder matrix_reps_index_map (batchencoder.h) */
  /* add_plain */
  u64 q_div_t_mod_q[MAX_K]; /* floor(Q/t) mod q_j, Q = prod of the L data primes */
  u64 q_mod_t, upper_half_threshold;
  u64 upper_half_increment[MAX_K]; /* q_j - t  (fast plain lift) */
  /* key switch */
  u64 inv_qsp[MAX_K], half, half_mod[MAX_K];
  /* BEHZ */
  u64 m_sk, gamma, m_tilde;
  u64 B[MAX_K];           /* base_B (L) */
  u64 Bsk[MAX_K + 1];     /* B + m_sk */
  ntt_tab *bt;            /* L+1 tables */
  u64 inv_punct_q[MAX_K]; /* (Q/q_i)^-1 mod q_i */
  u64 m_q2bsk[MAX_K + 1][MAX_K]; /* (Q/q_i) mod Bsk_p */
  u64 m_q2mt[MAX_K];      /* (Q/q_i) mod m_tilde */
  u64 neg_inv_q_mt;       /* -Q^-1 mod m_tilde */
  u64 q_mod_bsk[MAX_K + 1], inv_mt_bsk[MAX_K + 1], inv_q_bsk[MAX_K + 1];
  u64 inv_punct_B[MAX_K]; /* (P_B/b_i)^-1 mod b_i */
  u64 m_B2q[MAX_K][MAX_K]; /* (P_B/b_i) mod q_j : [j][i] */
  u64 m_B2msk[MAX_K];
  u64 inv_PB_msk;
  u64 PB_mod_q[MAX_K];
  /* keys */
  ksk_t keys[2][MAX_KEYS];
  int nkeys[2];
  u64 *relin;
};

static u64 prod_except(const u64 *base, int n, int skip, u64 mod) {
  u64 r = 1 % mod;
  for (int i = 0; i < n; i++)
    if (i != skip) r = mulmod(r, base[i] % mod, mod);
  return r;
}

hor_ctx *hor_create(uint64_t N, uint64_t t, const uint64_t *q, int K) {
  if (K < 2 || K > MAX_K - 2 || (N & (N - 1)) || N < 2 * PASTA_T) return NULL;
  hor_ctx *c = calloc(1, sizeof(*c));
  c->N = N;
  c->t = t;
  c->K = K;
  c->L = K - 1;
  int L = c->L;
  while (((u64)1 << c->logn) < N) c->logn++;
  memcpy(c->q, q, sizeof(u64) * K);
  c->qt = calloc(K, sizeof(ntt_tab));
  for (int i = 0; i < K; i++) ntt_tab_init(&c->qt[i], N, c->logn, q[i]);
  ntt_tab_init(&c->pt, N, c->logn, t);

  /* BatchEncoder index map (batchencoder.h:80-134): slot i <-> evaluation point 3^i, second row -3^i */
  c->index_map = malloc(sizeof(uint32_t) * N);
  u64 m = 2 * N, pos = 1, half = N / 2;
  for (u64 i = 0; i < half; i++) {
    c->index_map[i] = bitrev((unsigned)((pos - 1) >> 1), c->logn);
    c->index_map[half + i] = bitrev((unsigned)((m - pos - 1) >> 1), c->logn);
    pos = pos * 3 % m;
  }

  /* Q = prod_{i<L} q_i as a little-endian multiword; floor(Q/t) and Q mod t (util/scalingvariant.h) */
  u64 big[MAX_K + 1] = {1};
  int words = 1;
  for (int i = 0; i < L; i++) {
    u64 carry = 0;
    for (int wd = 0; wd < words; wd++) {
      u128 p = (u128)big[wd] * q[i] + carry;
      big[wd] = (u64)p;
      carry = (u64)(p >> 64);
    }
    if (carry) big[words++] = carry;
  }
  u64 quo[MAX_K + 1], rem = 0;
  for (int wd = words - 1; wd >= 0; wd--) {
    u128 cur = ((u128)rem << 64) | big[wd];
    quo[wd] = (u64)(cur / t);
    rem = (u64)(cur % t);
  }
  c->q_mod_t = rem;
  c->upper_half_threshold = (t + 1) >> 1;
  for (int j = 0; j < K; j++) {
    u64 r = 0;
    for (int wd = words - 1; wd >= 0; wd--) r = (u64)((((u128)r << 64) | quo[wd]) % q[j]);
    c->q_div_t_mod_q[j] = r;
    c->upper_half_increment[j] = q[j] - t;
  }

  /* key switching (Evaluator::switch_key_inplace, evaluator.h:1260) */
  u64 qsp = q[K - 1];
  c->half = qsp >> 1;
  for (int i = 0; i < L; i++) {
    c->inv_qsp[i] = invmod_prime(qsp % q[i], q[i]);
    c->half_mod[i] = c->half % q[i];
  }

  /* BEHZ auxiliary base (RNSTool::initialize, util/rns.h:190-400): 61-bit primes = 1 mod 2N, descending */
  u64 aux[MAX_K + 2];
  int found = 0;
  u64 factor = 2 * N, v = ((((u64)1 << 61) - 1) / factor) * factor + 1;
  while (found < L + 2) {
    if (is_prime64(v)) aux[found++] = v;
    v -= factor;
  }
  c->m_sk = aux[0];
  c->gamma = aux[1];
  c->m_tilde = (u64)1 << 32;
  for (int i = 0; i < L; i++) c->B[i] = c->Bsk[i] = aux[2 + i];
  c->Bsk[L] = c->m_sk;
  c->bt = calloc(L + 1, sizeof(ntt_tab));
  for (int i = 0; i <= L; i++) ntt_tab_init(&c->bt[i], N, c->logn, c->Bsk[i]);

  for (int i = 0; i < L; i++) {
    c->inv_punct_q[i] = invmod_prime(prod_except(q, L, i, q[i]), q[i]);
    for (int p = 0; p <= L; p++) c->m_q2bsk[p][i] = prod_except(q, L, i, c->Bsk[p]);
    c->m_q2mt[i] = prod_except(q, L, i, c->m_tilde);
    c->inv_punct_B[i] = invmod_prime(prod_except(c->B, L, i, c->B[i]), c->B[i]);
    for (int j = 0; j < L; j++) c->m_B2q[j][i] = prod_except(c->B, L, i, q[j]);
    c->m_B2msk[i] = prod_except(c->B, L, i, c->m_sk);
    c->PB_mod_q[i] = prod_except(c->B, L, -1, q[i]);
  }
  u64 q_mt = prod_except(q, L, -1, c->m_tilde), inv = 1; /* Newton inverse of an odd number mod 2^32 */
  for (int it = 0; it < 6; it++) inv = (inv * (2 - q_mt * inv)) & (c->m_tilde - 1);
  c->neg_inv_q_mt = (c->m_tilde - inv) & (c->m_tilde - 1);
  for (int p = 0; p <= L; p++) {
    u64 mod = c->Bsk[p];
    c->q_mod_bsk[p] = prod_except(q, L, -1, mod);
    c->inv_mt_bsk[p] = invmod_prime(c->m_tilde % mod, mod);
    c->inv_q_bsk[p] = invmod_prime(c->q_mod_bsk[p], mod);
  }
  c->inv_PB_msk = invmod_prime(prod_except(c->B, L, -1, c->m_sk), c->m_sk);
  return c;
}

void hor_destroy(hor_ctx *c) {
  if (!c) return;
  for (int i = 0; i < c->K; i++) ntt_tab_free(&c->qt[i]);
  for (int i = 0; i <= c->L; i++) ntt_tab_free(&c->bt[i]);
  ntt_tab_free(&c->pt);
  free(c->qt);
  free(c->bt);
  free(c->index_map);
  for (int s = 0; s < 2; s++)
    for (int i = 0; i < c->nkeys[s]; i++) free(c->keys[s][i].data);
  free(c->relin);
  free(c);
}

void hor_ntt_roots(const hor_ctx *c, uint64_t *out) {
  for (int i = 0; i < c->K; i++) out[i] = c->qt[i].psi;
  out[c->K] = c->pt.psi;
}

void hor_behz(const hor_ctx *c, uint64_t *out) {
  out[0] = c->m_sk;
  out[1] = c->gamma;
  out[2] = c->m_tilde;
  for (int i = 0; i < c->L; i++) out[3 + i] = c->B[i];
  for (int i = 0; i <= c->L; i++) out[3 + c->L + i] = c->bt[i].psi;
}

/* GaloisTool::get_elt_from_step (util/galois.h:124) */
uint32_t hor_galois_elt(const hor_ctx *c, int step) {
  u64 n = c->N, m = 2 * n;
  if (step == 0) return (uint32_t)(m - 1);
  u64 pos = step < 0 ? (u64)(-step) : (u64)step;
  if (pos >= (n >> 1)) return 0;
  if (step < 0) pos = (n >> 1) - pos;
  u64 e = 1;
  for (u64 i = 0; i < pos; i++) e = e * 3 % m;
  return (uint32_t)e;
}

/* util/numth.h:22-42 */
int hor_naf(int value, int *terms) {
  int sign = value < 0, n = 0;
  if (sign) value = -value;
  for (int i = 0; value; i++) {
    int zi = (value & 1) ? 2 - (value & 3) : 0;
    value = (value - zi) >> 1;
    if (zi) terms[n++] = (sign ? -zi : zi) * (1 << i);
  }
  return n;
}

int hor_load_ksk(hor_ctx *c, int kind, uint32_t elt, const uint64_t *data) {
  size_t words = (size_t)c->L * 2 * c->K * c->N;
  u64 *copy = malloc(words * sizeof(u64));
  memcpy(copy, data, words * sizeof(u64));
  if (kind == 2) {
    free(c->relin);
    c->relin = copy;
    return 0;
  }
  if (kind < 0 || kind > 1 || c->nkeys[kind] >= MAX_KEYS) {
    free(copy);
    return -1;
  }
  for (int i = 0; i < c->nkeys[kind]; i++)
    if (c->keys[kind][i].elt == elt) {
      free(c->keys[kind][i].data);
      c->keys[kind][i].data = copy;
      return 0;
    }
  c->keys[kind][c->nkeys[kind]].elt = elt;
  c->keys[kind][c->nkeys[kind]++].data = copy;
  return 0;
}

static const u64 *find_key(const hor_ctx *c, int ks, uint32_t elt) {
  if (ks < 0 || ks > 1) return NULL;
  for (int i = 0; i < c->nkeys[ks]; i++)
    if (c->keys[ks][i].elt == elt) return c->keys[ks][i].data;
  return NULL;
}

void hor_ntt(const hor_ctx *c, int limb, int inverse, uint64_t *data) {
  const ntt_tab *t = limb < c->K ? &c->qt[limb] : &c->bt[limb - c->K];
  if (inverse)
    ntt_inv(t, c->N, data);
  else
    ntt_fwd(t, c->N, data);
}

/* ------------------------------------------------------------------ SEAL-level ops */
#define CT_WORDS(c) ((size_t)2 * (c)->L * (c)->N)
#define LIMB(p, poly, i) ((p) + ((size_t)(poly) * c->L + (i)) * c->N)

/* BatchEncoder::encode (batchencoder.h:80): scatter through the index map, inverse NTT mod t */
void hor_encode(const hor_ctx *c, const uint64_t *slots, size_t n, uint64_t *pt) {
  memset(pt, 0, sizeof(u64) * c->N);
  for (size_t i = 0; i < n && i < c->N; i++) pt[c->index_map[i]] = slots[i] % c->t;
  ntt_inv(&c->pt, c->N, pt);
}

void hor_add(const hor_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out) {
  for (int p = 0; p < 2; p++)
    for (int i = 0; i < c->L; i++) {
      const u64 *x = LIMB(a, p, i), *y = LIMB(b, p, i);
      u64 *o = LIMB(out, p, i);
      for (u64 j = 0; j < c->N; j++) o[j] = addmod(x[j], y[j], c->q[i]);
    }
}

void hor_negate(const hor_ctx *c, const uint64_t *a, uint64_t *out) {
  for (int p = 0; p < 2; p++)
    for (int i = 0; i < c->L; i++) {
      const u64 *x = LIMB(a, p, i);
      u64 *o = LIMB(out, p, i);
      for (u64 j = 0; j < c->N; j++) o[j] = x[j] ? c->q[i] - x[j] : 0;
    }
}

/* Evaluator::add_plain, BFV (evaluator.h:665; util/scalingvariant.h multiply_add_plain_with_scaling_variant):
 * c0[j] += m_j*floor(Q/t) + floor((m_j*(Q mod t) + (t+1)/2) / t)   per limb */
void hor_add_plain(const hor_ctx *c, const uint64_t *a, const uint64_t *pt, uint64_t *out) {
  if (out != a) memcpy(out, a, CT_WORDS(c) * sizeof(u64));
  for (u64 j = 0; j < c->N; j++) {
    u64 mj = pt[j];
    u64 fix = (u64)(((u128)mj * c->q_mod_t + c->upper_half_threshold) / c->t);
    for (int i = 0; i < c->L; i++) {
      u64 *o = LIMB(out, 0, i);
      u64 scaled = (u64)(((u128)mj * c->q_div_t_mod_q[i] + fix) % c->q[i]);
      o[j] = addmod(o[j], scaled, c->q[i]);
    }
  }
}

/* centred ("fast plain") lift of a plaintext into limb i and forward NTT (Evaluator::multiply_plain_normal).
 * mono != 0: the plaintext has exactly one nonzero coefficient. SEAL then takes its monomial branch
 * (negacyclic_multiply_poly_mono_coeffmod) and, with fast plain lift (every q_i > t), multiplies by the coefficient AS IT IS,
 * also when it lies in the upper half: no centred lift. The product with m X^e equals the NTT product with the un-lifted
 * plaintext, so the same code serves; only the lift is skipped. (Verified limb-exact against the reference for t - 1.) */
static void lift_ntt(const hor_ctx *c, const u64 *pt, int i, int mono, u64 *dst) {
  for (u64 j = 0; j < c->N; j++) dst[j] = (!mono && pt[j] >= c->upper_half_threshold) ? pt[j] + c->upper_half_increment[i] : pt[j];
  ntt_fwd(&c->qt[i], c->N, dst);
}

/* Evaluator::multiply_plain on a coefficient-form ciphertext (evaluator.h:729) */
void hor_multiply_plain(const hor_ctx *c, const uint64_t *a, const uint64_t *pt, uint64_t *out) {
  u64 *m = malloc(sizeof(u64) * c->N), *x = malloc(sizeof(u64) * c->N);
  u64 nonzero = 0;
  for (u64 j = 0; j < c->N; j++) nonzero += pt[j] != 0;
  for (int i = 0; i < c->L; i++) {
    lift_ntt(c, pt, i, nonzero == 1, m);
    for (int p = 0; p < 2; p++) {
      memcpy(x, LIMB(a, p, i), sizeof(u64) * c->N);
      ntt_fwd(&c->qt[i], c->N, x);
      for (u64 j = 0; j < c->N; j++) x[j] = mulmod(x[j], m[j], c->q[i]);
      ntt_inv(&c->qt[i], c->N, x);
      memcpy(LIMB(out, p, i), x, sizeof(u64) * c->N);
    }
  }
  free(m);
  free(x);
}

/* GaloisTool::apply_galois on one coefficient-form limb (util/galois.h:32-33) */
static void galois_limb(const hor_ctx *c, const u64 *in, uint32_t elt, u64 q, u64 *out) {
  u64 raw = 0, mask = c->N - 1;
  for (u64 i = 0; i < c->N; i++) {
    u64 idx = raw & mask, v = in[i];
    if ((raw >> c->logn) & 1) v = v ? q - v : 0;
    out[idx] = v;
    raw += elt;
  }
}

/* Evaluator::switch_key_inplace, BFV branch (evaluator.h:1260): digit-wise re-reduction + NTT, inner product
 * with the key over all K limbs, then divide-and-round by the special prime. target: [L][N] coefficient form.
 * k0/k1: [L][N] results to be ADDED to the ciphertext components by the caller. */
static void key_switch(const hor_ctx *c, const u64 *target, const u64 *key, u64 *k0, u64 *k1) {
  int L = c->L, K = c->K;
  u64 N = c->N;
  u64 *acc = calloc((size_t)2 * K * N, sizeof(u64)), *tmp = malloc(sizeof(u64) * N);
  for (int J = 0; J < L; J++) {
    const u64 *dig = target + (size_t)J * N;
    for (int k = 0; k < K; k++) {
      u64 qk = c->q[k];
      for (u64 j = 0; j < N; j++) tmp[j] = dig[j] >= qk ? dig[j] % qk : dig[j];
      ntt_fwd(&c->qt[k], N, tmp);
      for (int comp = 0; comp < 2; comp++) {
        const u64 *kk = key + (((size_t)J * 2 + comp) * K + k) * N;
        u64 *ac = acc + ((size_t)comp * K + k) * N;
        for (u64 j = 0; j < N; j++) ac[j] = addmod(ac[j], mulmod(tmp[j], kk[j], qk), qk);
      }
    }
  }
  u64 qsp = c->q[K - 1];
  for (int comp = 0; comp < 2; comp++) {
    u64 *last = acc + ((size_t)comp * K + (K - 1)) * N;
    ntt_inv(&c->qt[K - 1], N, last);
    for (u64 j = 0; j < N; j++) last[j] = (last[j] + c->half) % qsp;
    u64 *dst = comp ? k1 : k0;
    for (int i = 0; i < L; i++) {
      u64 *ai = acc + ((size_t)comp * K + i) * N, qi = c->q[i];
      ntt_inv(&c->qt[i], N, ai);
      for (u64 j = 0; j < N; j++) {
        u64 v = submod(ai[j], last[j] % qi, qi);
        v = addmod(v, c->half_mod[i], qi);
        dst[(size_t)i * N + j] = mulmod(v, c->inv_qsp[i], qi);
      }
    }
  }
  free(acc);
  free(tmp);
}

static void apply_galois_with_key(const hor_ctx *c, const uint64_t *a, uint32_t elt, const u64 *key, uint64_t *out) {
  int L = c->L;
  u64 N = c->N;
  size_t poly = (size_t)L * N;
  u64 *g0 = calloc(poly, 8), *g1 = calloc(poly, 8), *k0 = malloc(poly * 8), *k1 = malloc(poly * 8);
  for (int i = 0; i < L; i++) {
    galois_limb(c, LIMB(a, 0, i), elt, c->q[i], g0 + (size_t)i * N);
    galois_limb(c, LIMB(a, 1, i), elt, c->q[i], g1 + (size_t)i * N);
  }
  key_switch(c, g1, key, k0, k1);
  for (int i = 0; i < L; i++)
    for (u64 j = 0; j < N; j++) {
      LIMB(out, 0, i)[j] = addmod(g0[(size_t)i * N + j], k0[(size_t)i * N + j], c->q[i]);
      LIMB(out, 1, i)[j] = k1[(size_t)i * N + j];
    }
  free(g0);
  free(g1);
  free(k0);
  free(k1);
}

int hor_apply_galois(const hor_ctx *c, const uint64_t *a, uint32_t elt, int ks, uint64_t *out) {
  const u64 *key = find_key(c, ks, elt);
  if (!key) return -1;
  apply_galois_with_key(c, a, elt, key, out);
  return 0;
}

/* Evaluator::rotate_internal (evaluator.h:955): direct key if present, else the NAF terms in emission order */
int hor_rotate_rows(const hor_ctx *c, const uint64_t *a, int steps, int ks, uint64_t *out) {
  if (steps == 0) {
    if (out != a) memcpy(out, a, CT_WORDS(c) * 8);
    return 0;
  }
  uint32_t elt = hor_galois_elt(c, steps);
  if (!elt) return -1;
  if (find_key(c, ks, elt)) return hor_apply_galois(c, a, elt, ks, out);
  int terms[40], n = hor_naf(steps, terms);
  if (n == 1) return -1;
  u64 *cur = malloc(CT_WORDS(c) * 8), *nxt = malloc(CT_WORDS(c) * 8);
  memcpy(cur, a, CT_WORDS(c) * 8);
  int rc = 0;
  for (int i = 0; i < n && !rc; i++) {
    int s = terms[i];
    if ((u64)(s < 0 ? -s : s) == (c->N >> 1)) continue;
    rc = hor_rotate_rows(c, cur, s, ks, nxt);
    u64 *sw = cur;
    cur = nxt;
    nxt = sw;
  }
  if (!rc) memcpy(out, cur, CT_WORDS(c) * 8);
  free(cur);
  free(nxt);
  return rc;
}

int hor_rotate_columns(const hor_ctx *c, const uint64_t *a, int ks, uint64_t *out) {
  return hor_apply_galois(c, a, (uint32_t)(2 * c->N - 1), ks, out);
}

int hor_relinearize(const hor_ctx *c, const uint64_t *a3, uint64_t *out) {
  if (!c->relin) return -1;
  int L = c->L;
  u64 N = c->N;
  size_t poly = (size_t)L * N;
  u64 *k0 = malloc(poly * 8), *k1 = malloc(poly * 8);
  key_switch(c, a3 + 2 * poly, c->relin, k0, k1);
  for (int i = 0; i < L; i++)
    for (u64 j = 0; j < N; j++) {
      out[(size_t)i * N + j] = addmod(a3[(size_t)i * N + j], k0[(size_t)i * N + j], c->q[i]);
      out[poly + (size_t)i * N + j] = addmod(a3[poly + (size_t)i * N + j], k1[(size_t)i * N + j], c->q[i]);
    }
  free(k0);
  free(k1);
  return 0;
}

/* ---- BEHZ multiplication (Evaluator::bfv_multiply, evaluator.h:214; RNSTool, util/rns.h:213-228) ---- */

/* steps (1)-(3): x (base q, [L][N]) -> NTT(x) in q ([L][N]) and NTT(x) in Bsk ([L+1][N]) */
static void behz_extend(const hor_ctx *c, const u64 *x, u64 *xq, u64 *xb) {
  int L = c->L;
  u64 N = c->N;
  u64 *z = malloc(sizeof(u64) * L * N);
  for (int i = 0; i < L; i++)
    for (u64 j = 0; j < N; j++) {
      u64 y = mulmod(x[(size_t)i * N + j], c->m_tilde % c->q[i], c->q[i]); /* fastbconv_m_tilde: x * m_tilde */
      z[(size_t)i * N + j] = mulmod(y, c->inv_punct_q[i], c->q[i]);
    }
  for (u64 j = 0; j < N; j++) {
    u64 mt = 0;
    for (int i = 0; i < L; i++) mt += z[(size_t)i * N + j] * c->m_q2mt[i];
    mt &= c->m_tilde - 1;
    u64 r = (mt * c->neg_inv_q_mt) & (c->m_tilde - 1); /* sm_mrq: r = -x/Q mod m_tilde */
    for (int p = 0; p <= L; p++) {
      u64 mod = c->Bsk[p];
      u128 s = 0;
      for (int i = 0; i < L; i++) s = (s + (u128)z[(size_t)i * N + j] * c->m_q2bsk[p][i]) % mod;
      u64 rp = r >= (c->m_tilde >> 1) ? r + mod - c->m_tilde : r;
      u64 v = (u64)(((u128)rp * c->q_mod_bsk[p] + s) % mod);
      xb[(size_t)p * N + j] = mulmod(v, c->inv_mt_bsk[p], mod);
    }
  }
  for (int i = 0; i < L; i++) {
    memcpy(xq + (size_t)i * N, x + (size_t)i * N, sizeof(u64) * N);
    ntt_fwd(&c->qt[i], N, xq + (size_t)i * N);
  }
  for (int p = 0; p <= L; p++) ntt_fwd(&c->bt[p], N, xb + (size_t)p * N);
  free(z);
}

/* steps (5)-(6): d (coefficient form in q [L][N] and Bsk [L+1][N]) -> round(t*d/Q) in base q */
static void behz_scale_round(const hor_ctx *c, const u64 *dq, const u64 *db, u64 *out) {
  int L = c->L;
  u64 N = c->N;
  for (u64 j = 0; j < N; j++) {
    u64 z[MAX_K], f[MAX_K + 1];
    for (int i = 0; i < L; i++) {
      u64 tq = mulmod(dq[(size_t)i * N + j], c->t % c->q[i], c->q[i]);
      z[i] = mulmod(tq, c->inv_punct_q[i], c->q[i]);
    }
    for (int p = 0; p <= L; p++) { /* fast_floor */
      u64 mod = c->Bsk[p];
      u128 s = 0;
      for (int i = 0; i < L; i++) s = (s + (u128)z[i] * c->m_q2bsk[p][i]) % mod;
      u64 tb = mulmod(db[(size_t)p * N + j], c->t % mod, mod);
      f[p] = mulmod(submod(tb, (u64)s, mod), c->inv_q_bsk[p], mod);
    }
    u64 zb[MAX_K]; /* fastbconv_sk */
    for (int i = 0; i < L; i++) zb[i] = mulmod(f[i], c->inv_punct_B[i], c->B[i]);
    u128 sk = 0;
    for (int i = 0; i < L; i++) sk = (sk + (u128)zb[i] * c->m_B2msk[i]) % c->m_sk;
    u64 alpha = mulmod(submod((u64)sk, f[L], c->m_sk), c->inv_PB_msk, c->m_sk);
    for (int jq = 0; jq < L; jq++) {
      u64 qj = c->q[jq];
      u128 s = 0;
      for (int i = 0; i < L; i++) s = (s + (u128)zb[i] * c->m_B2q[jq][i]) % qj;
      u64 v = (u64)s;
      if (alpha > (c->m_sk >> 1))
        v = (u64)(((u128)(c->m_sk - alpha) * c->PB_mod_q[jq] + v) % qj);
      else
        v = (u64)(((u128)alpha * (qj - c->PB_mod_q[jq]) + v) % qj);
      out[(size_t)jq * N + j] = v;
    }
  }
}

void hor_multiply(const hor_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out3) {
  int L = c->L;
  u64 N = c->N;
  size_t pq = (size_t)L * N, pb = (size_t)(L + 1) * N;
  u64 *xq[4], *xb[4];
  const u64 *src[4] = {a, a + pq, b, b + pq};
  for (int s = 0; s < 4; s++) {
    xq[s] = malloc(pq * 8);
    xb[s] = malloc(pb * 8);
    behz_extend(c, src[s], xq[s], xb[s]);
  }
  u64 *dq = malloc(pq * 8), *db = malloc(pb * 8);
  for (int d = 0; d < 3; d++) {
    for (int i = 0; i < L; i++) {
      u64 qi = c->q[i], *o = dq + (size_t)i * N;
      for (u64 j = 0; j < N; j++) {
        size_t ix = (size_t)i * N + j;
        if (d == 0)
          o[j] = mulmod(xq[0][ix] % qi, xq[2][ix] % qi, qi);
        else if (d == 2)
          o[j] = mulmod(xq[1][ix], xq[3][ix], qi);
        else
          o[j] = addmod(mulmod(xq[0][ix], xq[3][ix], qi), mulmod(xq[1][ix], xq[2][ix], qi), qi);
      }
      ntt_inv(&c->qt[i], N, o);
    }
    for (int p = 0; p <= L; p++) {
      u64 mod = c->Bsk[p], *o = db + (size_t)p * N;
      for (u64 j = 0; j < N; j++) {
        size_t ix = (size_t)p * N + j;
        if (d == 0)
          o[j] = mulmod(xb[0][ix], xb[2][ix], mod);
        else if (d == 2)
          o[j] = mulmod(xb[1][ix], xb[3][ix], mod);
        else
          o[j] = addmod(mulmod(xb[0][ix], xb[3][ix], mod), mulmod(xb[1][ix], xb[2][ix], mod), mod);
      }
      ntt_inv(&c->bt[p], N, o);
    }
    behz_scale_round(c, dq, db, out3 + (size_t)d * pq);
  }
  for (int s = 0; s < 4; s++) {
    free(xq[s]);
    free(xb[s]);
  }
  free(dq);
  free(db);
}

/* Evaluator::exponentiate_inplace(x, 3, rk) (evaluator.h:621) == relin(multiply(relin(multiply(x,x)), x)) */
int hor_exponentiate3(const hor_ctx *c, const uint64_t *a, uint64_t *out) {
  size_t w = CT_WORDS(c);
  u64 *t3 = malloc(w / 2 * 3 * 8), *sq = malloc(w * 8);
  hor_multiply(c, a, a, t3);
  int rc = hor_relinearize(c, t3, sq);
  if (!rc) {
    hor_multiply(c, sq, a, t3);
    rc = hor_relinearize(c, t3, out);
  }
  free(t3);
  free(sq);
  return rc;
}

/* sealhelper::encrypted_vec_sum (src/util/sealhelper.cpp:379-392): every rotation starts from the input */
int hor_vec_sum(const hor_ctx *c, const uint64_t *a, size_t n, int ks, uint64_t *out) {
  size_t w = CT_WORDS(c);
  u64 *rot = malloc(w * 8), *acc = malloc(w * 8);
  memcpy(acc, a, w * 8);
  int rc = 0;
  for (size_t i = 1; i < n && !rc; i++) {
    rc = hor_rotate_rows(c, a, -(int)i, ks, rot);
    if (!rc) hor_add(c, acc, rot, acc);
  }
  if (!rc) memcpy(out, acc, w * 8);
  free(rot);
  free(acc);
  return rc;
}

/* SEALZpCipher::mask (src/pasta/SEAL_Cipher.cpp:161-166) */
void hor_mask(const hor_ctx *c, const uint64_t *a, const uint64_t *mask, size_t n, uint64_t *out) {
  u64 *pt = malloc(sizeof(u64) * c->N);
  hor_encode(c, mask, n, pt);
  hor_multiply_plain(c, a, pt, out);
  free(pt);
}

/* SEALZpCipher::flatten (src/pasta/SEAL_Cipher.cpp:170-181) */
int hor_flatten(const hor_ctx *c, const uint64_t *cts, size_t count, int ks, uint64_t *out) {
  size_t w = CT_WORDS(c);
  u64 *tmp = malloc(w * 8), *acc = malloc(w * 8);
  memcpy(acc, cts, w * 8);
  int rc = 0;
  for (size_t i = 1; i < count && !rc; i++) {
    rc = hor_rotate_rows(c, cts + i * w, -(int)(i * PASTA_T), ks, tmp);
    if (!rc) hor_add(c, acc, tmp, acc);
  }
  if (!rc) memcpy(out, acc, w * 8);
  free(tmp);
  free(acc);
  return rc;
}

/* ------------------------------------------------------------------ Keccak-f[1600] / SHAKE128 */
/* Restates libs/keccak (XKCP KeccakHash.c / KeccakSponge.inc): rate 1344 bits, delimited suffix 0x1F. */
static const u64 keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int keccak_rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static inline u64 rol64(u64 x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static void keccak_f(u64 s[25]) {
  for (int round = 0; round < 24; round++) {
    u64 C[5], D[5], Bm[25];
    for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
    for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol64(C[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) s[i] ^= D[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) Bm[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y], keccak_rot[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) s[x + 5 * y] = Bm[x + 5 * y] ^ (~Bm[(x + 1) % 5 + 5 * y] & Bm[(x + 2) % 5 + 5 * y]);
    s[0] ^= keccak_rc[round];
  }
}

#define SHAKE128_RATE 168
typedef struct {
  u64 s[25];
  int pos; /* next unread byte of the current squeeze block */
} shake_t;

static void shake_absorb_final(shake_t *sh, const uint8_t *in, size_t inlen) {
  memset(sh, 0, sizeof(*sh));
  uint8_t *b = (uint8_t *)sh->s; /* little-endian host assumed (x86-64), as XKCP's opt64 does */
  size_t off = 0;
  while (inlen - off >= SHAKE128_RATE) {
    for (int i = 0; i < SHAKE128_RATE; i++) b[i] ^= in[off + i];
    keccak_f(sh->s);
    off += SHAKE128_RATE;
  }
  for (size_t i = 0; off + i < inlen; i++) b[i] ^= in[off + i];
  b[inlen - off] ^= 0x1F;
  b[SHAKE128_RATE - 1] ^= 0x80;
  keccak_f(sh->s);
  sh->pos = 0;
}
static void shake_squeeze(shake_t *sh, uint8_t *out, size_t n) {
  uint8_t *b = (uint8_t *)sh->s;
  for (size_t i = 0; i < n; i++) {
    if (sh->pos == SHAKE128_RATE) {
      keccak_f(sh->s);
      sh->pos = 0;
    }
    out[i] = b[sh->pos++];
  }
}
void hor_shake128(const uint8_t *in, size_t inlen, uint8_t *out, size_t outlen) {
  shake_t sh;
  shake_absorb_final(&sh, in, inlen);
  shake_squeeze(&sh, out, outlen);
}

/* ------------------------------------------------------------------ plain PASTA-3 */
typedef struct {
  shake_t sh;
  u64 p, mask;
} pasta_t;

/* Pasta::init_shake (src/pasta/pasta_3_plain.cpp:56-68): seed = BE64(nonce) || BE64(counter) */
static void pasta_init(pasta_t *ps, u64 p, u64 nonce, u64 counter) {
  uint8_t seed[16];
  for (int i = 0; i < 8; i++) {
    seed[i] = (uint8_t)(nonce >> (56 - 8 * i));
    seed[8 + i] = (uint8_t)(counter >> (56 - 8 * i));
  }
  shake_absorb_final(&ps->sh, seed, 16);
  ps->p = p;
  u64 bits = 0, v = p;
  while (v) {
    bits++;
    v >>= 1;
  }
  ps->mask = ((u64)1 << bits) - 1; /* max_prime_size (pasta_3_plain.cpp:145-151) */
}
/* Pasta::generate_random_field_element (pasta_3_plain.cpp:72-82) */
static u64 pasta_field_element(pasta_t *ps, int allow_zero) {
  for (;;) {
    uint8_t r[8];
    shake_squeeze(&ps->sh, r, 8);
    u64 be = 0;
    for (int i = 0; i < 8; i++) be = (be << 8) | r[i];
    u64 ele = be & ps->mask;
    if (!allow_zero && ele == 0) continue;
    if (ele < ps->p) return ele;
  }
}
/* Pasta::get_random_matrix / calculate_row (pasta_3_plain.cpp:86-119): row-major mat[128][128] */
static void pasta_matrix(pasta_t *ps, u64 *mat) {
  for (int j = 0; j < PASTA_T; j++) mat[j] = pasta_field_element(ps, 0);
  for (int i = 1; i < PASTA_T; i++) {
    const u64 *prev = mat + (size_t)(i - 1) * PASTA_T;
    u64 *row = mat + (size_t)i * PASTA_T;
    for (int j = 0; j < PASTA_T; j++) {
      u64 v = mulmod(mat[j], prev[PASTA_T - 1], ps->p);
      if (j) v = (v + prev[j - 1]) % ps->p;
      row[j] = v;
    }
  }
}

void hor_pasta_layer_material(uint64_t p, uint64_t nonce, uint64_t counter, int layer, uint64_t *mat1, uint64_t *mat2,
                              uint64_t *rc) {
  pasta_t ps;
  pasta_init(&ps, p, nonce, counter);
  for (int l = 0; l <= layer; l++) {
    pasta_matrix(&ps, mat1);
    pasta_matrix(&ps, mat2);
    for (int i = 0; i < 2 * PASTA_T; i++) rc[i] = pasta_field_element(&ps, 1); /* get_rc_vec (:286-295) */
  }
}

/* Pasta::gen_keystream (pasta_3_plain.cpp:156-171 and round functions :198-282). NB the plain cipher draws
 * matrix 1, matrix 2, rc 1, rc 2 per layer in the same stream order as the homomorphic side. */
void hor_pasta_keystream(const uint64_t *key256, uint64_t p, uint64_t nonce, uint64_t counter, uint64_t *ks128) {
  pasta_t ps;
  pasta_init(&ps, p, nonce, counter);
  u64 s1[PASTA_T], s2[PASTA_T], n1[PASTA_T], n2[PASTA_T];
  u64 *m = malloc(sizeof(u64) * PASTA_T * PASTA_T);
  memcpy(s1, key256, sizeof(s1));
  memcpy(s2, key256 + PASTA_T, sizeof(s2));
  for (int r = 0; r <= 3; r++) {
    for (int half = 0; half < 2; half++) { /* matmul(state1_), matmul(state2_) */
      u64 *st = half ? s2 : s1, *nw = half ? n2 : n1;
      pasta_matrix(&ps, m);
      for (int i = 0; i < PASTA_T; i++) {
        u64 acc = 0;
        for (int j = 0; j < PASTA_T; j++) acc = (acc + mulmod(m[(size_t)i * PASTA_T + j], st[j], p)) % p;
        nw[i] = acc;
      }
    }
    for (int i = 0; i < PASTA_T; i++) s1[i] = (n1[i] + pasta_field_element(&ps, 1)) % p;
    for (int i = 0; i < PASTA_T; i++) s2[i] = (n2[i] + pasta_field_element(&ps, 1)) % p;
    for (int i = 0; i < PASTA_T; i++) { /* mix */
      u64 sum = (s1[i] + s2[i]) % p;
      s1[i] = (s1[i] + sum) % p;
      s2[i] = (s2[i] + sum) % p;
    }
    if (r == 3) break;
    for (int half = 0; half < 2; half++) {
      u64 *st = half ? s2 : s1;
      if (r == 2) { /* cube */
        for (int i = 0; i < PASTA_T; i++) st[i] = mulmod(mulmod(st[i], st[i], p), st[i], p);
      } else { /* Feistel */
        u64 nw[PASTA_T];
        nw[0] = st[0];
        for (int i = 1; i < PASTA_T; i++) nw[i] = (mulmod(st[i - 1], st[i - 1], p) + st[i]) % p;
        memcpy(st, nw, sizeof(nw));
      }
    }
  }
  memcpy(ks128, s1, sizeof(s1));
  free(m);
}

/* PASTA::encrypt / decrypt (pasta_3_plain.cpp:9-47), nonce fixed by the caller */
void hor_pasta_plain(const uint64_t *key256, uint64_t p, const uint64_t *in, size_t n, int decrypt, uint64_t *out) {
  u64 ks[PASTA_T];
  for (size_t b = 0; b * PASTA_T < n; b++) {
    hor_pasta_keystream(key256, p, 123456789ULL, b, ks);
    for (size_t i = b * PASTA_T; i < (b + 1) * PASTA_T && i < n; i++)
      out[i] = decrypt ? submod(in[i] % p, ks[i - b * PASTA_T], p) : (in[i] + ks[i - b * PASTA_T]) % p;
  }
}

/* ------------------------------------------------------------------ homomorphic PASTA-3 */

/* PASTA_SEAL::diagonal (src/pasta/pasta_3_seal.cpp:370-413) */
static int he_diagonal(const hor_ctx *c, u64 *state, const u64 *mat1, const u64 *mat2) {
  size_t w = CT_WORDS(c), half = c->N / 2;
  u64 *tmp = malloc(w * 8), *sum = malloc(w * 8), *pt = malloc(c->N * 8), *diag = calloc(c->N, 8);
  int rc = 0;
  if (half != PASTA_T) {
    rc = hor_rotate_rows(c, state, PASTA_T, 0, tmp);
    if (!rc) hor_add(c, state, tmp, state);
  }
  for (int i = 0; i < PASTA_T && !rc; i++) {
    for (int j = 0; j < PASTA_T; j++) {
      diag[j] = mat1[(size_t)j * PASTA_T + (j + PASTA_T - i) % PASTA_T];
      diag[j + half] = mat2[(size_t)j * PASTA_T + (j + PASTA_T - i) % PASTA_T];
    }
    hor_encode(c, diag, half + PASTA_T, pt);
    if (i == 0) {
      hor_multiply_plain(c, state, pt, sum);
    } else {
      rc = hor_rotate_rows(c, state, -1, 0, tmp);
      if (rc) break;
      memcpy(state, tmp, w * 8);
      hor_multiply_plain(c, state, pt, tmp);
      hor_add(c, sum, tmp, sum);
    }
  }
  if (!rc) memcpy(state, sum, w * 8);
  free(tmp);
  free(sum);
  free(pt);
  free(diag);
  return rc;
}

/* PASTA_SEAL::babystep_giantstep (src/pasta/pasta_3_seal.cpp:267-366), N1 = 16, N2 = 8 */
static int he_bsgs(const hor_ctx *c, u64 *state, const u64 *mat1, const u64 *mat2) {
  enum { N1 = 16, N2 = 8 };
  size_t w = CT_WORDS(c), half = c->N / 2, slots = c->N;
  u64 *pts = malloc((size_t)PASTA_T * c->N * 8), *diag = malloc(slots * 8), *d1 = malloc(half * 8), *d2 = malloc(half * 8);
  for (int i = 0; i < PASTA_T; i++) {
    int k = i / N1;
    memset(d1, 0, half * 8);
    memset(d2, 0, half * 8);
    for (int j = 0; j < PASTA_T; j++) { /* diagonal i, then std::rotate left by k*N1 */
      int src = (j + k * N1) % PASTA_T;
      d1[j] = mat1[(size_t)src * PASTA_T + (src + PASTA_T - i) % PASTA_T];
      d2[j] = mat2[(size_t)src * PASTA_T + (src + PASTA_T - i) % PASTA_T];
    }
    if (half != PASTA_T) {
      for (int mm = 0; mm < k * N1; mm++) {
        size_t is = PASTA_T - 1 - mm, id = half - 1 - mm;
        d1[id] = d1[is];
        d1[is] = 0;
        d2[id] = d2[is];
        d2[is] = 0;
      }
    }
    memcpy(diag, d1, half * 8);
    memcpy(diag + half, d2, half * 8);
    hor_encode(c, diag, slots, pts + (size_t)i * c->N);
  }
  int rc = 0;
  u64 *tmp = malloc(w * 8), *rot = malloc(w * 8 * N1), *inner = malloc(w * 8), *outer = malloc(w * 8);
  if (half != PASTA_T) {
    rc = hor_rotate_rows(c, state, PASTA_T, 0, tmp);
    if (!rc) hor_add(c, state, tmp, state);
  }
  memcpy(rot, state, w * 8);
  for (int j = 1; j < N1 && !rc; j++) rc = hor_rotate_rows(c, rot + (size_t)(j - 1) * w, -1, 0, rot + (size_t)j * w);
  for (int k = 0; k < N2 && !rc; k++) {
    hor_multiply_plain(c, rot, pts + (size_t)(k * N1) * c->N, inner);
    for (int j = 1; j < N1; j++) {
      hor_multiply_plain(c, rot + (size_t)j * w, pts + (size_t)(k * N1 + j) * c->N, tmp);
      hor_add(c, inner, tmp, inner);
    }
    if (!k) {
      memcpy(outer, inner, w * 8);
    } else {
      rc = hor_rotate_rows(c, inner, -k * N1, 0, tmp);
      if (!rc) hor_add(c, outer, tmp, outer);
    }
  }
  if (!rc) memcpy(state, outer, w * 8);
  free(pts);
  free(diag);
  free(d1);
  free(d2);
  free(tmp);
  free(rot);
  free(inner);
  free(outer);
  return rc;
}

/* PASTA_SEAL::sbox_feistel (src/pasta/pasta_3_seal.cpp:222-247) */
static int he_feistel(const hor_ctx *c, u64 *state) {
  size_t w = CT_WORDS(c), half = c->N / 2;
  u64 *rot = malloc(w * 8), *t3 = malloc(w / 2 * 3 * 8), *mask = calloc(c->N, 8), *pt = malloc(c->N * 8);
  int rc = hor_rotate_rows(c, state, -1, 0, rot);
  if (!rc) {
    for (size_t i = 1; i < PASTA_T; i++) mask[i] = mask[half + i] = 1;
    hor_encode(c, mask, half + PASTA_T, pt);
    hor_multiply_plain(c, rot, pt, rot);
    hor_multiply(c, rot, rot, t3); /* square == multiply(x, x) limb-exactly */
    rc = hor_relinearize(c, t3, rot);
    if (!rc) hor_add(c, state, rot, state);
  }
  free(rot);
  free(t3);
  free(mask);
  free(pt);
  return rc;
}

/* PASTA_SEAL::decomposition (src/pasta/pasta_3_seal.cpp:106-172); == HE_decrypt (:42-104) */
int hor_pasta_decompose(const hor_ctx *c, const uint64_t *enc_key, const uint64_t *sym_ct, size_t n, uint64_t nonce,
                        uint64_t first_counter, int use_bsgs, uint64_t *out) {
  size_t w = CT_WORDS(c), half = c->N / 2;
  size_t nblk = (n + PASTA_T - 1) / PASTA_T;
  u64 *m1 = malloc(8 * PASTA_T * PASTA_T), *m2 = malloc(8 * PASTA_T * PASTA_T);
  u64 *state = malloc(w * 8), *tmp = malloc(w * 8), *pt = malloc(c->N * 8), *slots = calloc(c->N, 8);
  int rc = 0;
  for (size_t b = 0; b < nblk && !rc; b++) {
    pasta_t ps;
    pasta_init(&ps, c->t, nonce, first_counter + b);
    memcpy(state, enc_key, w * 8);
    for (int r = 1; r <= 4 && !rc; r++) {
      pasta_matrix(&ps, m1);
      pasta_matrix(&ps, m2);
      memset(slots, 0, c->N * 8);
      for (int i = 0; i < PASTA_T; i++) slots[i] = pasta_field_element(&ps, 1);
      for (int i = 0; i < PASTA_T; i++) slots[half + i] = pasta_field_element(&ps, 1);
      rc = use_bsgs ? he_bsgs(c, state, m1, m2) : he_diagonal(c, state, m1, m2);
      if (rc) break;
      hor_encode(c, slots, half + PASTA_T, pt); /* add_rc (:205-211) */
      hor_add_plain(c, state, pt, state);
      rc = hor_rotate_columns(c, state, 0, tmp); /* mix (:417-423) */
      if (rc) break;
      hor_add(c, tmp, state, tmp);
      hor_add(c, state, tmp, state);
      if (r == 3)
        rc = hor_exponentiate3(c, state, state); /* sbox_cube (:215-218) */
      else if (r < 3)
        rc = he_feistel(c, state);
    }
    if (rc) break;
    size_t cnt = n - b * PASTA_T < PASTA_T ? n - b * PASTA_T : PASTA_T;
    hor_encode(c, sym_ct + b * PASTA_T, cnt, pt);
    hor_negate(c, state, state);
    hor_add_plain(c, state, pt, out + b * w);
  }
  free(m1);
  free(m2);
  free(state);
  free(tmp);
  free(pt);
  free(slots);
  return rc;
}
