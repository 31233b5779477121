// Device runtime shim: memory, copies, stream sync and kernel launch.
// CUDA build (nvcc, sm_100a): the real thing. -DHHE_EMULATE (g++, tests/emul only): the grid is walked on the host
// so host-side orchestration and kernel index arithmetic can be unit-tested without a GPU. The emulation object is
// never part of libhhe_b200.so and the Python package never loads it.
#pragma once
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "hd.h"

#ifdef HHE_CUDA
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#endif

#ifndef HHE_MAX_THREADS
#define HHE_MAX_THREADS 1024
#endif

namespace hhe {

#ifdef HHE_CUDA
inline void cuda_check(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// launch bounds: bodies may declare kMaxThreads / kMinBlocks (e.g. 512 x 2 CTAs per SM); default HHE_MAX_THREADS x 1
template <class B, class = void>
struct BodyBounds {
  static constexpr int kT = HHE_MAX_THREADS, kM = 1;
};
template <class B>
struct BodyBounds<B, decltype(void(B::kMinBlocks))> {
  static constexpr int kT = B::kMaxThreads, kM = B::kMinBlocks;
};

// Programmatic dependent launch (PDL): every kernel is launched with the programmatic-stream-serialization attribute and starts with
// griddepcontrol.wait (all memory of the preceding kernel is visible after it; a no-op for an ordinary launch) followed by
// griddepcontrol.launch_dependents, so the NEXT kernel's launch and CTA set-up overlap this kernel's last wave instead of
// starting after it has drained. The engine's chains are hundreds of short dependent kernels (518 rotations x 3 per PASTA block).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <class Body>
__global__ void __launch_bounds__(BodyBounds<Body>::kT, BodyBounds<Body>::kM) kernel_entry(const Body body) {
  extern __shared__ __align__(16) unsigned char hhe_smem[];
  pdl_prologue();
  body(static_cast<int>(blockIdx.x), static_cast<int>(blockDim.x), hhe_smem);
}

// Two-CTA thread-block cluster: the bodies of a pair run phase1 on their own shared memory, meet at a cluster barrier,
// then run phase2 with a pointer to the partner's shared memory (distributed shared memory, read-only use).
// Barrier semantics are chosen per body (Body::kPeerSmem, default true):
//   kPeerSmem  phase2 reads the partner's shared memory: the first barrier is a release/acquire pair (cluster.sync()); the
//              closing barrier only keeps this CTA's shared memory alive until the partner is done with it, which needs no
//              memory ordering: relaxed arrive (no MEMBAR.ALL.GPU + ERRBAR behind the CTA's global stores).
//   !kPeerSmem the barrier only separates the pair's global READS (consumed into registers before the arrive) from the
//              first global WRITE of an in-place transform: relaxed arrive, and no closing barrier at all.
template <class B, class = void>
struct BodyPeerSmem {
  static constexpr bool value = true;
};
template <class B>
struct BodyPeerSmem<B, decltype(void(B::kPeerSmem))> {
  static constexpr bool value = B::kPeerSmem;
};

__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}

template <class Body>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BodyBounds<Body>::kT, BodyBounds<Body>::kM) kernel_entry_c2(const Body body, const int strict) {
  extern __shared__ __align__(16) unsigned char hhe_smem[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  pdl_prologue();
  body.phase1(static_cast<int>(blockIdx.x), static_cast<int>(blockDim.x), hhe_smem);
  if (BodyPeerSmem<Body>::value || strict)
    cluster.sync();
  else
    cluster_sync_relaxed();
  const unsigned char *peer = cluster.map_shared_rank(hhe_smem, cluster.block_rank() ^ 1u);
  body.phase2(static_cast<int>(blockIdx.x), static_cast<int>(blockDim.x), hhe_smem, peer);
  if (strict)
    cluster.sync();
  else if (BodyPeerSmem<Body>::value)
    cluster_sync_relaxed();  // the partner may still be reading this CTA's shared memory
}

// Eight-CTA cluster (portable maximum): phase1 on the CTA's own shared memory, a release/acquire cluster barrier, phase2 with the
// shared-memory base of every rank (distributed shared memory, read-only), closing barrier so no CTA exits while it is being read.
template <class Body>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(BodyBounds<Body>::kT, BodyBounds<Body>::kM) kernel_entry_c8(const Body body) {
  extern __shared__ __align__(16) unsigned char hhe_smem[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  pdl_prologue();
  body.phase1(static_cast<int>(blockIdx.x), static_cast<int>(blockDim.x), hhe_smem);
  cluster.sync();
  unsigned char *peers[8];
#pragma unroll
  for (unsigned r = 0; r < 8; ++r) peers[r] = cluster.map_shared_rank(hhe_smem, r);
  body.phase2(static_cast<int>(blockIdx.x), static_cast<int>(blockDim.x), peers);
  cluster_sync_relaxed();
}
#endif

// Per-kernel device-time accounting (hhe_profile_*): when enabled every launch is bracketed by CUDA events on the
// launching stream; durations are resolved lazily after a stream sync. Used by bench.py for the live roofline figure.
struct KernelStat {
  const char *name;
  uint64_t launches;
  double ms;
};

struct Device {
#ifdef HHE_CUDA
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  struct Pending {
    int kid;
    cudaEvent_t a, b;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;
#endif
  static constexpr int kMaxDevices = 64;
  int ordinal = 0;  // CUDA device this context lives on
  int sm_count = 1;
  bool strict_cluster = false;  // HHE_STRICT_CLUSTER=1: release/acquire cluster barriers everywhere (A/B switch)
  uint64_t launches = 0;
  bool profiling = false;
  std::vector<KernelStat> stats;

  int kernel_id(const char *name) {
    for (size_t i = 0; i < stats.size(); ++i)
      if (!std::strcmp(stats[i].name, name)) return static_cast<int>(i);
    stats.push_back(KernelStat{name, 0, 0.0});
    return static_cast<int>(stats.size() - 1);
  }
  void profile_reset() {
    profile_resolve();
    for (auto &s : stats) s.launches = 0, s.ms = 0.0;
  }
  void profile_resolve() {
#ifdef HHE_CUDA
    if (pending.empty()) return;
    cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    for (auto &p : pending) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, p.a, p.b);
      stats[p.kid].ms += ms;
      event_pool.push_back(p.a);
      event_pool.push_back(p.b);
    }
    pending.clear();
#endif
  }
#ifdef HHE_CUDA
  cudaEvent_t get_event() {
    if (!event_pool.empty()) {
      cudaEvent_t e = event_pool.back();
      event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cuda_check(cudaEventCreate(&e), "cudaEventCreate");
    return e;
  }
#endif

  void *dmalloc(size_t bytes) {
#ifdef HHE_CUDA
    void *p = nullptr;
    cuda_check(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc");
    return p;
#else
    void *p = std::malloc(bytes ? bytes : 1);
    if (!p) throw std::runtime_error("malloc failed");
    return p;
#endif
  }
  void dfree(void *p) {
#ifdef HHE_CUDA
    if (p) cudaFree(p);
#else
    std::free(p);
#endif
  }
  void h2d(void *dst, const void *src, size_t bytes) {
#ifdef HHE_CUDA
    cuda_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D");
#else
    std::memcpy(dst, src, bytes);
#endif
  }
  void d2h(void *dst, const void *src, size_t bytes) {
#ifdef HHE_CUDA
    cuda_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H");
#else
    std::memcpy(dst, src, bytes);
#endif
  }
  void d2d(void *dst, const void *src, size_t bytes) {
#ifdef HHE_CUDA
    cuda_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream), "cudaMemcpyAsync D2D");
#else
    std::memmove(dst, src, bytes);
#endif
  }
  void zero(void *dst, size_t bytes) {
#ifdef HHE_CUDA
    cuda_check(cudaMemsetAsync(dst, 0, bytes, stream), "cudaMemsetAsync");
#else
    std::memset(dst, 0, bytes);
#endif
  }
  void sync() {
#ifdef HHE_CUDA
    cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
#endif
  }

  // ---- overlapped delivery of results to host memory (pinned-memory streaming at the boundary) ----
  // d2h_begin(slot, ...) enqueues the copy of a finished device buffer on a SECOND stream, ordered behind everything the
  // compute stream holds at that moment, and returns; the compute stream goes on with the next chunk. Pinned destinations
  // (cudaHostAlloc / cudaHostRegister memory, e.g. a serving loop's ring) are written directly; pageable destinations
  // (a std::vector, seal::Ciphertext::data()) go through one of two pinned staging buffers and are filled by d2h_end.
  // d2h_end(slot) blocks until the data of that slot has reached its destination (and must precede reuse of the device buffer).
#ifdef HHE_CUDA
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t produced = nullptr;
  struct Stage {
    void *pinned = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    void *dst = nullptr;
    size_t bytes = 0;
    bool pending = false, staged = false;
  } stage[2];
#endif
  void d2h_begin(int slot, void *host_dst, const void *dev_src, size_t bytes) {
#ifdef HHE_CUDA
    d2h_end(slot);
    Stage &st = stage[slot];
    if (!copy_stream) {
      cuda_check(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking), "cudaStreamCreate(copy)");
      cuda_check(cudaEventCreateWithFlags(&produced, cudaEventDisableTiming), "cudaEventCreate");
    }
    if (!st.done) cuda_check(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming), "cudaEventCreate");
    cudaPointerAttributes attr{};
    const bool pinned_dst = cudaPointerGetAttributes(&attr, host_dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();  // a pageable pointer is not an error here
    void *target = host_dst;
    st.staged = !pinned_dst;
    if (st.staged) {
      if (st.cap < bytes) {
        if (st.pinned) cudaFreeHost(st.pinned);
        st.pinned = nullptr, st.cap = 0;
        cuda_check(cudaHostAlloc(&st.pinned, bytes, cudaHostAllocDefault), "cudaHostAlloc(staging)");
        st.cap = bytes;
      }
      target = st.pinned;
    }
    cuda_check(cudaEventRecord(produced, stream), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(copy_stream, produced, 0), "cudaStreamWaitEvent");
    cuda_check(cudaMemcpyAsync(target, dev_src, bytes, cudaMemcpyDeviceToHost, copy_stream), "cudaMemcpyAsync D2H (copy stream)");
    cuda_check(cudaEventRecord(st.done, copy_stream), "cudaEventRecord");
    st.dst = host_dst, st.bytes = bytes, st.pending = true;
#else
    (void)slot;
    std::memcpy(host_dst, dev_src, bytes);
#endif
  }
  void d2h_end(int slot) {
#ifdef HHE_CUDA
    Stage &st = stage[slot];
    if (!st.pending) return;
    cuda_check(cudaEventSynchronize(st.done), "cudaEventSynchronize(copy)");
    if (st.staged) std::memcpy(st.dst, st.pinned, st.bytes);
    st.pending = false;
#else
    (void)slot;
#endif
  }
  void d2h_release() {
#ifdef HHE_CUDA
    for (auto &st : stage) {
      if (st.pending) cudaEventSynchronize(st.done);
      st.pending = false;
      if (st.pinned) cudaFreeHost(st.pinned);
      if (st.done) cudaEventDestroy(st.done);
      st = Stage{};
    }
    if (produced) cudaEventDestroy(produced), produced = nullptr;
    if (copy_stream) cudaStreamDestroy(copy_stream), copy_stream = nullptr;
#endif
  }

#ifdef HHE_CUDA
  bool pdl = true;  // programmatic dependent launch (HHE_NO_PDL=1 switches it off)
  template <class Kernel, class... Args>
  void launch_ex(Kernel kernel, size_t grid, int nt, size_t smem_bytes, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(nt));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cuda_check(cudaLaunchKernelEx(&cfg, kernel, args...), "cudaLaunchKernelEx");
  }
#endif

  template <class Body>
  void launch(const Body &body, size_t grid, int nt, size_t smem_bytes) {
    if (grid == 0) return;
    ++launches;
    int kid = -1;
    if (profiling) {
      kid = kernel_id(Body::kName);
      ++stats[kid].launches;
    }
#ifdef HHE_CUDA
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    if (profiling) {
      ev_a = get_event();
      ev_b = get_event();
      cuda_check(cudaEventRecord(ev_a, stream), "cudaEventRecord");
    }
    if (smem_bytes > 48 * 1024) {
      static size_t configured[kMaxDevices] = {};  // per Body instantiation and device (the attribute is per device)
      size_t &conf = configured[ordinal & (kMaxDevices - 1)];
      if (smem_bytes > conf) {
        cuda_check(cudaFuncSetAttribute(kernel_entry<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem_bytes)),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        conf = smem_bytes;
      }
    }
    launch_ex(kernel_entry<Body>, grid, nt, smem_bytes, body);
    cuda_check(cudaGetLastError(), "kernel launch");
    if (profiling) {
      cuda_check(cudaEventRecord(ev_b, stream), "cudaEventRecord");
      pending.push_back(Pending{kid, ev_a, ev_b});
      if (pending.size() >= 8192) profile_resolve();
    }
#else
    std::vector<unsigned char> smem(smem_bytes + 16);
    for (size_t b = 0; b < grid; ++b) body(static_cast<int>(b), nt, smem.data());
#endif
  }

  // grid must be a multiple of 8: CTAs 8p .. 8p+7 form a cluster (see kernel_entry_c8)
  template <class Body>
  void launch_cluster8(const Body &body, size_t grid, int nt, size_t smem_bytes) {
    if (grid == 0) return;
    if (grid & 7) throw std::invalid_argument("cluster launch needs a grid that is a multiple of 8");
    ++launches;
    int kid = -1;
    if (profiling) {
      kid = kernel_id(Body::kName);
      ++stats[kid].launches;
    }
#ifdef HHE_CUDA
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    if (profiling) {
      ev_a = get_event();
      ev_b = get_event();
      cuda_check(cudaEventRecord(ev_a, stream), "cudaEventRecord");
    }
    if (smem_bytes > 48 * 1024) {
      static size_t configured[kMaxDevices] = {};  // per Body instantiation and device
      size_t &conf = configured[ordinal & (kMaxDevices - 1)];
      if (smem_bytes > conf) {
        cuda_check(cudaFuncSetAttribute(kernel_entry_c8<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem_bytes)),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        conf = smem_bytes;
      }
    }
    launch_ex(kernel_entry_c8<Body>, grid, nt, smem_bytes, body);
    cuda_check(cudaGetLastError(), "cluster-8 kernel launch");
    if (profiling) {
      cuda_check(cudaEventRecord(ev_b, stream), "cudaEventRecord");
      pending.push_back(Pending{kid, ev_a, ev_b});
      if (pending.size() >= 8192) profile_resolve();
    }
#else
    std::vector<std::vector<unsigned char>> sm(8, std::vector<unsigned char>(smem_bytes + 16));
    unsigned char *peers[8];
    for (int r = 0; r < 8; ++r) peers[r] = sm[r].data();
    for (size_t b = 0; b < grid; b += 8) {
      for (int r = 0; r < 8; ++r) body.phase1(static_cast<int>(b + r), nt, peers[r]);
      for (int r = 0; r < 8; ++r) body.phase2(static_cast<int>(b + r), nt, peers);
    }
#endif
  }

  // grid must be even: CTAs 2p and 2p+1 form a cluster (see kernel_entry_c2)
  template <class Body>
  void launch_cluster2(const Body &body, size_t grid, int nt, size_t smem_bytes) {
    if (grid == 0) return;
    if (grid & 1) throw std::invalid_argument("cluster launch needs an even grid");
    ++launches;
    int kid = -1;
    if (profiling) {
      kid = kernel_id(Body::kName);
      ++stats[kid].launches;
    }
#ifdef HHE_CUDA
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    if (profiling) {
      ev_a = get_event();
      ev_b = get_event();
      cuda_check(cudaEventRecord(ev_a, stream), "cudaEventRecord");
    }
    if (smem_bytes > 48 * 1024) {
      static size_t configured[kMaxDevices] = {};  // per Body instantiation and device
      size_t &conf = configured[ordinal & (kMaxDevices - 1)];
      if (smem_bytes > conf) {
        cuda_check(cudaFuncSetAttribute(kernel_entry_c2<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem_bytes)),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        conf = smem_bytes;
      }
    }
    launch_ex(kernel_entry_c2<Body>, grid, nt, smem_bytes, body, strict_cluster ? 1 : 0);
    cuda_check(cudaGetLastError(), "cluster kernel launch");
    if (profiling) {
      cuda_check(cudaEventRecord(ev_b, stream), "cudaEventRecord");
      pending.push_back(Pending{kid, ev_a, ev_b});
      if (pending.size() >= 8192) profile_resolve();
    }
#else
    std::vector<unsigned char> s0(smem_bytes + 16), s1(smem_bytes + 16);
    for (size_t b = 0; b < grid; b += 2) {
      body.phase1(static_cast<int>(b), nt, s0.data());
      body.phase1(static_cast<int>(b + 1), nt, s1.data());
      body.phase2(static_cast<int>(b), nt, s0.data(), s1.data());
      body.phase2(static_cast<int>(b + 1), nt, s1.data(), s0.data());
    }
#endif
  }
};

}  // namespace hhe
