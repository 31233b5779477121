// Host/device glue. Kernel bodies are written once as functors whose operator() is structured in
// barrier-separated phases: FOR_THREADS(tid) { ... } SYNC(); No per-thread state crosses a SYNC() -- everything a
// later phase needs lives in shared or global memory. Under nvcc a body runs as one CUDA thread of a CTA; when the
// same header is compiled by g++ with -DHHE_EMULATE (tests/emul only, never shipped, never loaded by the package)
// FOR_THREADS becomes a sequential loop over the CTA's threads so the index arithmetic of every kernel can be unit
// tested on a machine without a GPU. The emulation build is a test harness for host logic, not a CPU fallback.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__) && !defined(HHE_EMULATE)
#define HHE_CUDA 1
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#else
#define HD inline
#define DEV inline
#endif

#if defined(__CUDA_ARCH__)
#define FOR_THREADS(tid, nt) for (int tid = threadIdx.x, hhe_once_ = 1; hhe_once_; hhe_once_ = 0)
#define SYNC() __syncthreads()
#define SYNCWARP() __syncwarp()
#elif defined(HHE_EMUL_ORDER) && HHE_EMUL_ORDER == 1
// race check of the emulation harness (tools/order_emul.sh): the threads of a phase run in DESCENDING order ...
#define FOR_THREADS(tid, nt) for (int hhe_i_ = (nt)-1, tid = hhe_i_; hhe_i_ >= 0; tid = --hhe_i_)
#define SYNC() ((void)0)
#define SYNCWARP() ((void)0)
#elif defined(HHE_EMUL_ORDER) && HHE_EMUL_ORDER == 2
// ... or in a scrambled order (odd multiplier modulo a power-of-two CTA size; ascending otherwise): a phase whose result depends on
// the order of its threads (a missing barrier between a write and another thread's read) no longer matches the oracle
#define FOR_THREADS(tid, nt)                                                                                                       \
  for (int hhe_i_ = 0, hhe_p_ = (((nt) & ((nt)-1)) == 0), tid = hhe_p_ ? (11 & ((nt)-1)) : 0; hhe_i_ < (nt);                       \
       ++hhe_i_, tid = hhe_p_ ? ((hhe_i_ * 37 + 11) & ((nt)-1)) : hhe_i_)
#define SYNC() ((void)0)
#define SYNCWARP() ((void)0)
#else
#define FOR_THREADS(tid, nt) for (int tid = 0; tid < (nt); ++tid)
#define SYNC() ((void)0)
#define SYNCWARP() ((void)0)
#endif

#ifndef HHE_MAX_THREADS
#define HHE_MAX_THREADS 1024
#endif

namespace hhe {
using u32 = uint32_t;
using u64 = uint64_t;

// One thread of the CTA asks the L2 to fetch `bytes` (multiple of 16) at the 16-byte aligned global address p
// (cp.async.bulk.prefetch.L2, SASS UBLKPF.L2): no register, no LSU slot, nothing to wait for. Kernels whose load phase is
// a chain of dependent DRAM round trips use it to pull the operands of the CTA that will run two waves later into the L2.
HD void cta_prefetch_l2(const void *p, std::size_t bytes) {
#if defined(__CUDA_ARCH__)
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(static_cast<unsigned>(bytes)) : "memory");
#else
  (void)p;
  (void)bytes;
#endif
}

HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return static_cast<u64>((static_cast<unsigned __int128>(a) * b) >> 64);
#endif
}
}  // namespace hhe
