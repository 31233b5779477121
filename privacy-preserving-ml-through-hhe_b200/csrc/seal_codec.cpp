// SEAL 4.0 wire-format codec (see seal_codec.h). Host-side byte work only; no SEAL code is linked.
#include "seal_codec.h"

#include <dlfcn.h>
#include <zlib.h>

#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>

namespace hhe {
namespace sealio {
namespace {

// ------------------------------------------------------------------------------------------------ BLAKE2b (RFC 7693)
constexpr uint64_t kIv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                             0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
constexpr uint8_t kSigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

void compress(uint64_t h[8], const uint8_t block[128], uint64_t t, bool last) {
  uint64_t m[16], v[16];
  std::memcpy(m, block, 128);  // little-endian host
  for (int i = 0; i < 8; ++i) v[i] = h[i], v[i + 8] = kIv[i];
  v[12] ^= t;
  if (last) v[14] = ~v[14];
  for (int r = 0; r < 12; ++r) {
    const uint8_t *s = kSigma[r];
    auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
      v[a] = v[a] + v[b] + x, v[d] = rotr(v[d] ^ v[a], 32);
      v[c] = v[c] + v[d], v[b] = rotr(v[b] ^ v[c], 24);
      v[a] = v[a] + v[b] + y, v[d] = rotr(v[d] ^ v[a], 16);
      v[c] = v[c] + v[d], v[b] = rotr(v[b] ^ v[c], 63);
    };
    G(0, 4, 8, 12, m[s[0]], m[s[1]]);
    G(1, 5, 9, 13, m[s[2]], m[s[3]]);
    G(2, 6, 10, 14, m[s[4]], m[s[5]]);
    G(3, 7, 11, 15, m[s[6]], m[s[7]]);
    G(0, 5, 10, 15, m[s[8]], m[s[9]]);
    G(1, 6, 11, 12, m[s[10]], m[s[11]]);
    G(2, 7, 8, 13, m[s[12]], m[s[13]]);
    G(3, 4, 9, 14, m[s[14]], m[s[15]]);
  }
  for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

// ------------------------------------------------------------------------------------------------ zstd (dlopen: no headers in the image)
struct ZBufIn {
  const void *src;
  size_t size, pos;
};
struct ZBufOut {
  void *dst;
  size_t size, pos;
};
struct Zstd {
  size_t (*compressBound)(size_t) = nullptr;
  size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
  unsigned (*isError)(size_t) = nullptr;
  void *(*createDCtx)() = nullptr;
  size_t (*freeDCtx)(void *) = nullptr;
  size_t (*decompressStream)(void *, ZBufOut *, ZBufIn *) = nullptr;
  bool ok = false;
};

const Zstd &zstd() {
  static Zstd z;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    z.compressBound = reinterpret_cast<decltype(z.compressBound)>(dlsym(h, "ZSTD_compressBound"));
    z.compress = reinterpret_cast<decltype(z.compress)>(dlsym(h, "ZSTD_compress"));
    z.isError = reinterpret_cast<decltype(z.isError)>(dlsym(h, "ZSTD_isError"));
    z.createDCtx = reinterpret_cast<decltype(z.createDCtx)>(dlsym(h, "ZSTD_createDCtx"));
    z.freeDCtx = reinterpret_cast<decltype(z.freeDCtx)>(dlsym(h, "ZSTD_freeDCtx"));
    z.decompressStream = reinterpret_cast<decltype(z.decompressStream)>(dlsym(h, "ZSTD_decompressStream"));
    z.ok = z.compressBound && z.compress && z.isError && z.createDCtx && z.freeDCtx && z.decompressStream;
  });
  return z;
}

// ------------------------------------------------------------------------------------------------ SEALHeader
constexpr size_t kHdr = 16;
constexpr size_t kCtMeta = 32 + 1 + 8 + 8 + 8 + 8 + 8;  // parms_id .. correction_factor

struct Header {
  int compr;
  uint64_t size;
};

Header read_header(const uint8_t *in, size_t len) {
  if (!in || len < kHdr) throw std::invalid_argument("buffer too small for a SEALHeader");
  uint16_t magic;
  std::memcpy(&magic, in, 2);
  if (magic != 0xA15E || in[2] != kHdr) throw std::logic_error("loaded SEALHeader is invalid");
  if (in[3] != 4 && in[3] != 3) throw std::logic_error("incompatible version");  // SEAL 4.0 loads 3.4+ headers of the same layout
  Header h;
  h.compr = in[5];
  std::memcpy(&h.size, in + 8, 8);
  if (h.compr > kComprZstd) throw std::logic_error("unsupported compression mode");
  if (h.size < kHdr || h.size > len) throw std::logic_error("loaded SEALHeader is invalid");
  return h;
}

void write_header(uint8_t *out, int compr, uint64_t size) {
  const uint16_t magic = 0xA15E;
  std::memcpy(out, &magic, 2);
  out[2] = kHdr, out[3] = 4, out[4] = 0, out[5] = static_cast<uint8_t>(compr), out[6] = out[7] = 0;
  std::memcpy(out + 8, &size, 8);
}

// Inflate the body of the object at `in` (header already parsed). Raw bodies are returned as a view (no copy).
struct Body {
  const uint8_t *p;
  size_t n;
  std::vector<uint8_t> own;
};

// `limit`: the largest body the encryption parameters allow for this kind of object. These streams come from other parties
// (gRPC payloads), so the output buffer never grows beyond it: a crafted stream is rejected as invalid data instead of
// forcing multi-GB host allocations before the member checks run.
void inflate_body(const uint8_t *in, const Header &h, size_t hint, size_t limit, Body &b) {
  const uint8_t *src = in + kHdr;
  const size_t n = h.size - kHdr;
  if (h.compr == kComprNone) {
    b.p = src, b.n = n;
    return;
  }
  limit += 4096;  // slack: a decoder may need one more call (with room) to report the end of a body of exactly `limit` bytes
  b.own.resize(std::min(limit, hint ? hint : 4 * n + 4096));
  auto grow = [&] {
    if (b.own.size() >= limit) throw std::logic_error("stream decompression failed: object larger than the encryption parameters allow");
    b.own.resize(std::min(limit, b.own.size() * 2));
  };
  size_t produced = 0;
  if (h.compr == kComprZlib) {
    z_stream zs{};
    if (inflateInit(&zs) != Z_OK) throw std::logic_error("stream decompression failed");
    zs.next_in = const_cast<Bytef *>(src);
    size_t fed = 0;
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {
      if (zs.avail_in == 0 && fed < n) {
        const size_t chunk = std::min<size_t>(n - fed, 1u << 30);
        zs.next_in = const_cast<Bytef *>(src + fed), zs.avail_in = static_cast<uInt>(chunk), fed += chunk;
      }
      if (produced == b.own.size()) grow();
      const size_t room = std::min<size_t>(b.own.size() - produced, 1u << 30);
      zs.next_out = b.own.data() + produced, zs.avail_out = static_cast<uInt>(room);
      rc = inflate(&zs, Z_NO_FLUSH);
      produced += room - zs.avail_out;
      if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) {
        inflateEnd(&zs);
        throw std::logic_error("stream decompression failed");
      }
      if (rc == Z_BUF_ERROR && zs.avail_in == 0 && fed == n) {
        inflateEnd(&zs);
        throw std::logic_error("stream decompression failed");
      }
    }
    inflateEnd(&zs);
  } else {
    const Zstd &z = zstd();
    if (!z.ok) throw std::runtime_error("libzstd.so.1 is not available: cannot read zstd-compressed SEAL objects");
    void *d = z.createDCtx();
    if (!d) throw std::logic_error("stream decompression failed");
    ZBufIn zi{src, n, 0};
    size_t rc = 1;
    while (zi.pos < zi.size || rc != 0) {
      if (produced == b.own.size()) grow();
      ZBufOut zo{b.own.data(), b.own.size(), produced};
      rc = z.decompressStream(d, &zo, &zi);
      const bool stalled = zo.pos == produced && zi.pos == zi.size && rc != 0;
      produced = zo.pos;
      if (z.isError(rc) || (stalled && produced < b.own.size())) {
        z.freeDCtx(d);
        throw std::logic_error("stream decompression failed");
      }
    }
    z.freeDCtx(d);
  }
  b.own.resize(produced);
  b.p = b.own.data(), b.n = produced;
}

// Deflate `body` into out (after a header); returns total object size.
size_t deflate_object(const uint8_t *body, size_t n, int compr, uint8_t *out, size_t cap) {
  if (compr == kComprNone) {
    if (cap < kHdr + n) throw std::invalid_argument("output buffer too small");
    std::memcpy(out + kHdr, body, n);
    write_header(out, compr, kHdr + n);
    return kHdr + n;
  }
  size_t stored = 0;
  if (compr == kComprZlib) {
    uLongf dl = cap > kHdr ? static_cast<uLongf>(cap - kHdr) : 0;
    const int rc = compress2(out + kHdr, &dl, body, static_cast<uLong>(n), Z_DEFAULT_COMPRESSION);
    if (rc == Z_BUF_ERROR) throw std::invalid_argument("output buffer too small");
    if (rc != Z_OK) throw std::logic_error("stream compression failed");
    stored = dl;
  } else if (compr == kComprZstd) {
    const Zstd &z = zstd();
    if (!z.ok) throw std::runtime_error("libzstd.so.1 is not available: cannot write zstd-compressed SEAL objects");
    if (cap < kHdr) throw std::invalid_argument("output buffer too small");
    const size_t rc = z.compress(out + kHdr, cap - kHdr, body, n, 3);
    if (z.isError(rc)) throw std::invalid_argument("output buffer too small or compression failed");
    stored = rc;
  } else {
    throw std::invalid_argument("unsupported compression mode");
  }
  write_header(out, compr, kHdr + stored);
  return kHdr + stored;
}

struct Rd {
  const uint8_t *p;
  size_t n, o = 0;
  void need(size_t k) const {
    if (o + k > n || o + k < o) throw std::logic_error("unexpected end of serialized data");
  }
  uint64_t u64() {
    need(8);
    uint64_t v;
    std::memcpy(&v, p + o, 8);
    o += 8;
    return v;
  }
  uint8_t u8() {
    need(1);
    return p[o++];
  }
  void skip(size_t k) {
    need(k);
    o += k;
  }
};

// Parses Ciphertext::save_members from `rd`; returns a pointer to the residues. level: 0 data, 1 key.
const uint8_t *parse_ct_members(const Ring &r, Rd &rd, int level, bool want_ntt, int *size_out) {
  uint64_t want[4], got[4];
  parms_id(r, level, want);
  for (auto &g : got) g = rd.u64();
  if (std::memcmp(want, got, 32)) throw std::logic_error("ciphertext data is invalid: parms_id does not match the encryption parameters");
  const bool ntt = rd.u8() != 0;
  const uint64_t size = rd.u64(), n = rd.u64(), limbs = rd.u64();
  rd.skip(16);  // scale (1.0 for BFV), correction_factor (1)
  const uint64_t lv = level ? r.K() : r.L();
  if (n != r.N || limbs != lv || size < 2 || size > 6) throw std::logic_error("ciphertext data is invalid");
  if (ntt != want_ntt) throw std::logic_error(want_ntt ? "key data is invalid: not in NTT form" : "ciphertext data is invalid: BFV ciphertexts on this path are not in NTT form");
  // DynArray object (always stored raw inside the already-inflated parent)
  rd.need(kHdr);
  const Header h = read_header(rd.p + rd.o, rd.n - rd.o);
  if (h.compr != kComprNone) throw std::logic_error("unexpected compression of the ciphertext payload");
  rd.skip(kHdr);
  const uint64_t count = rd.u64();
  if (count != size * limbs * n) throw std::logic_error("ciphertext data is invalid: seeded (Serializable) ciphertexts are not supported");
  if (h.size != kHdr + 8 + count * 8) throw std::logic_error("ciphertext data is invalid");
  const uint8_t *data = rd.p + rd.o;
  rd.skip(count * 8);
  *size_out = static_cast<int>(size);
  return data;
}

// is_data_valid_for: every residue below its modulus
void check_ranges(const Ring &r, const uint64_t *ct, int size, int limbs) {
  for (int c = 0; c < size; ++c)
    for (int i = 0; i < limbs; ++i) {
      const uint64_t qi = r.q[static_cast<size_t>(i)];
      const uint64_t *p = ct + (static_cast<size_t>(c) * limbs + i) * r.N;
      uint64_t bad = 0;
      for (uint64_t j = 0; j < r.N; ++j) bad |= static_cast<uint64_t>(p[j] >= qi);
      if (bad) throw std::logic_error("ciphertext data is invalid");
    }
}

}  // namespace

void blake2b_256(const void *in, size_t len, uint8_t out[32]) {
  uint64_t h[8];
  for (int i = 0; i < 8; ++i) h[i] = kIv[i];
  h[0] ^= 0x01010000ULL ^ 32;  // digest length 32, no key, fanout = depth = 1
  const uint8_t *p = static_cast<const uint8_t *>(in);
  uint64_t t = 0;
  while (len > 128) {
    t += 128;
    compress(h, p, t, false);
    p += 128, len -= 128;
  }
  uint8_t last[128] = {0};
  if (len) std::memcpy(last, p, len);
  t += len;
  compress(h, last, t, true);
  std::memcpy(out, h, 32);
}

void parms_id(const Ring &r, int level, uint64_t out[4]) {
  std::vector<uint64_t> w;
  w.push_back(1);  // scheme_type::bfv
  w.push_back(r.N);
  const int n = level ? r.K() : r.L();
  for (int i = 0; i < n; ++i) w.push_back(r.q[static_cast<size_t>(i)]);
  w.push_back(r.t);
  uint8_t d[32];
  blake2b_256(w.data(), w.size() * 8, d);
  std::memcpy(out, d, 32);
}

size_t ct_save_bound(const Ring &r, int size) {
  const size_t raw = kHdr + kCtMeta + kHdr + 8 + static_cast<size_t>(size) * r.L() * r.N * 8;
  return raw + raw / 128 + 1024;  // covers ZSTD_compressBound / compressBound of the body
}

size_t ct_save(const Ring &r, const uint64_t *ct, int size, int compr, uint8_t *out, size_t cap) {
  if (!ct || !out) throw std::invalid_argument("null buffer");
  if (size < 2 || size > 6) throw std::invalid_argument("ciphertext size must be in [2, 6]");
  if (compr < 0 || compr > kComprZstd) throw std::invalid_argument("unsupported compression mode");
  const uint64_t count = static_cast<uint64_t>(size) * r.L() * r.N;
  const size_t body_n = kCtMeta + kHdr + 8 + count * 8;
  // the raw object is laid out in place when no compression is asked for; otherwise in a temporary
  std::vector<uint8_t> tmp;
  uint8_t *body;
  if (compr == kComprNone) {
    if (cap < kHdr + body_n) throw std::invalid_argument("output buffer too small");
    body = out + kHdr;
  } else {
    tmp.resize(body_n);
    body = tmp.data();
  }
  uint64_t id[4];
  parms_id(r, 0, id);
  size_t o = 0;
  auto put = [&](const void *p, size_t n) {
    std::memcpy(body + o, p, n);
    o += n;
  };
  put(id, 32);
  const uint8_t ntt = 0;
  put(&ntt, 1);
  const uint64_t s64 = static_cast<uint64_t>(size), n64 = r.N, l64 = static_cast<uint64_t>(r.L()), cf = 1;
  const double scale = 1.0;
  put(&s64, 8), put(&n64, 8), put(&l64, 8), put(&scale, 8), put(&cf, 8);
  write_header(body + o, kComprNone, kHdr + 8 + count * 8);
  o += kHdr;
  put(&count, 8);
  put(ct, count * 8);
  if (compr == kComprNone) {
    write_header(out, kComprNone, kHdr + body_n);
    return kHdr + body_n;
  }
  return deflate_object(body, body_n, compr, out, cap);
}

size_t ct_load(const Ring &r, const uint8_t *in, size_t len, uint64_t *ct, size_t cap_words, int *size) {
  if (!ct || !size) throw std::invalid_argument("null buffer");
  const Header h = read_header(in, len);
  Body b;
  const size_t ct_max = kCtMeta + kHdr + 8 + 3 * static_cast<size_t>(r.L()) * r.N * 8;  // a size-3 ciphertext, the largest on the path
  inflate_body(in, h, ct_max, ct_max, b);
  Rd rd{b.p, b.n};
  int sz = 0;
  const uint8_t *data = parse_ct_members(r, rd, 0, false, &sz);
  const size_t words = static_cast<size_t>(sz) * r.L() * r.N;
  if (words > cap_words) throw std::invalid_argument("ciphertext buffer too small for the loaded size");
  std::memcpy(ct, data, words * 8);
  check_ranges(r, ct, sz, r.L());
  *size = sz;
  return h.size;
}

size_t keys_walk(const Ring &r, const uint8_t *in, size_t len, const std::function<void(uint64_t, const uint64_t *)> &on_key) {
  const Header h = read_header(in, len);
  Body b;
  // one key = L PublicKey objects (size-2 ciphertexts at key level, each with its own header); at most N key slots
  const size_t pk_max = kCtMeta + kHdr + 8 + static_cast<size_t>(2) * r.K() * r.N * 8;
  const size_t set_max = 32 + 8 + r.N * (8 + static_cast<size_t>(r.L()) * (kHdr + pk_max));
  inflate_body(in, h, 0, set_max, b);
  Rd rd{b.p, b.n};
  uint64_t want[4], got[4];
  parms_id(r, 1, want);
  for (auto &g : got) g = rd.u64();
  if (std::memcmp(want, got, 32)) throw std::logic_error("key data is invalid: parms_id does not match the encryption parameters");
  const uint64_t dim1 = rd.u64();
  if (dim1 > r.N) throw std::logic_error("key data is invalid");
  const int L = r.L(), K = r.K();
  const size_t per = static_cast<size_t>(2) * K * r.N;
  std::vector<uint64_t> ksk(static_cast<size_t>(L) * per);
  for (uint64_t idx = 0; idx < dim1; ++idx) {
    const uint64_t dim2 = rd.u64();
    if (!dim2) continue;
    if (dim2 != static_cast<uint64_t>(L)) throw std::logic_error("key data is invalid: unexpected decomposition count");
    for (int j = 0; j < L; ++j) {
      rd.need(kHdr);
      const Header kh = read_header(rd.p + rd.o, rd.n - rd.o);
      Body kb;
      inflate_body(rd.p + rd.o, kh, pk_max, pk_max, kb);
      Rd krd{kb.p, kb.n};
      int sz = 0;
      const uint8_t *data = parse_ct_members(r, krd, 1, true, &sz);
      if (sz != 2) throw std::logic_error("key data is invalid");
      std::memcpy(ksk.data() + static_cast<size_t>(j) * per, data, per * 8);
      check_ranges(r, ksk.data() + static_cast<size_t>(j) * per, 2, K);
      rd.skip(kh.size);
    }
    on_key(idx, ksk.data());
  }
  return h.size;
}

}  // namespace sealio
}  // namespace hhe
