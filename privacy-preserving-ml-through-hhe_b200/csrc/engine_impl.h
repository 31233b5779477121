// Private helpers shared by the engine's translation units (engine*.cu). The engine is split over several files only so
// that they compile in parallel; each instantiates the kernels it launches.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

#include "engine.h"

namespace hhe {

namespace {  // NOLINT: one copy per translation unit on purpose

constexpr int kEwThreads = 256;

inline size_t ew_grid(size_t total) { return (total + kEwThreads - 1) / kEwThreads; }

inline int ntt_threads(int logS) {
  int groups = 1 << (logS - kRadixLog);
  return std::max(32, std::min(HHE_MAX_THREADS, groups));
}

inline bool getenv_flag(const char *name) {
  const char *v = std::getenv(name);
  return v && *v && *v != '0';
}

inline TabMap map_mod(int limbs, int period, int base) {
  TabMap m{};
  for (int l = 0; l < limbs && l < kMaxMapLimbs; ++l) m.id[l] = static_cast<unsigned char>(base + (l % period));
  return m;
}

// mod-2N inverse of an odd Galois element
inline u32 inv_mod_2n(u32 elt, u64 two_n) {
  u64 inv = 1;
  for (int i = 0; i < 6; ++i) inv = (inv * (2 - static_cast<u64>(elt) * inv)) & (two_n - 1);
  return static_cast<u32>(inv);
}

}  // namespace

#define HHE_DISPATCH_LOG(value, ...)                                                       \
  switch (value) {                                                                         \
    case 8: { constexpr int LOGV = 8; __VA_ARGS__; } break;                                \
    case 9: { constexpr int LOGV = 9; __VA_ARGS__; } break;                                \
    case 10: { constexpr int LOGV = 10; __VA_ARGS__; } break;                              \
    case 11: { constexpr int LOGV = 11; __VA_ARGS__; } break;                              \
    case 12: { constexpr int LOGV = 12; __VA_ARGS__; } break;                              \
    case 13: { constexpr int LOGV = 13; __VA_ARGS__; } break;                              \
    case 14: { constexpr int LOGV = 14; __VA_ARGS__; } break;                              \
    default: throw std::invalid_argument("poly_modulus_degree not supported by the shared-memory NTT (512..16384)"); \
  }

}  // namespace hhe
