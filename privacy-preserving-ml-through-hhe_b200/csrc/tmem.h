// Tensor memory (TMEM) used as a per-thread accumulator file.
//
// Blackwell's 256 KB of tensor memory per SM is normally the tcgen05.mma accumulator; nothing on this path uses the
// tensor cores, so the key-switch kernel parks its 128 KB of inner-product accumulators there (tcgen05.st / tcgen05.ld,
// SASS STTM / LDTM). That frees the shared memory they occupied, so two CTAs fit per SM and their load / transform /
// multiply-accumulate phases overlap instead of running in lock step.
//
// Layout: TMEM is 128 lanes x 512 columns x 32 bit; a warp may only touch lanes [32*(warp%4), +32). With the 32x32b
// shape thread i of a warp owns TMEM lane 32*(warp%4)+i, and N consecutive columns of that lane are N registers of that
// thread. Warps w, w+4, w+8, ... share a lane quarter and use disjoint column blocks. A slot = 16 columns = 8 doubles.
//
// Under -DHHE_EMULATE (tests/emul) the same interface is backed by plain memory.
#pragma once
#include "hd.h"
#include "modarith_f64.h"

namespace hhe {

struct TmemAcc {
#if defined(__CUDA_ARCH__)
  u32 addr0;
  static DEV TmemAcc make(u32 base, int tid, int /*nt*/, int slots_per_thread, double * /*emu*/) {
    const u32 warp = static_cast<u32>(tid) >> 5;
    return TmemAcc{base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * static_cast<u32>(slots_per_thread) * 16u};
  }
  DEV void ld8(int slot, double *v) const {
    u32 r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr0 + static_cast<u32>(slot) * 16u));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __hiloint2double(static_cast<int>(r[2 * e + 1]), static_cast<int>(r[2 * e]));
  }
  DEV void st8(int slot, const double *v) const {
    u32 r[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      r[2 * e] = static_cast<u32>(__double2loint(v[e]));
      r[2 * e + 1] = static_cast<u32>(__double2hiint(v[e]));
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        :
        : "r"(addr0 + static_cast<u32>(slot) * 16u), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
          "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    // the wait is needed right here: deferring it to the next access of the slot gave wrong sums on B200 (the source
    // registers are evidently read asynchronously), see DESIGN.md
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
#else
  double *emu;
  int tid, nt;
  static inline TmemAcc make(u32, int tid, int nt, int, double *emu) { return TmemAcc{emu, tid, nt}; }
  inline void ld8(int slot, double *v) const {
    for (int e = 0; e < 8; ++e) v[e] = emu[static_cast<size_t>(slot * 8 + e) * nt + tid];
  }
  inline void st8(int slot, const double *v) const {
    for (int e = 0; e < 8; ++e) emu[static_cast<size_t>(slot * 8 + e) * nt + tid] = v[e];
  }
#endif
};

// columns a CTA of `nt` threads needs for `slots_per_thread` slots per thread (power of two >= 32, <= 512)
HD int tmem_columns(int nt, int slots_per_thread) {
  const int blocks = ((nt >> 5) + 3) >> 2;
  int need = blocks * slots_per_thread * 16, c = 32;
  while (c < need) c <<= 1;
  return c;
}

#if defined(__CUDA_ARCH__)
// executed by every thread of the CTA; returns the TMEM base address of the allocation
DEV u32 tmem_alloc_cta(u32 *slot_in_smem, int ncols) {
  if (threadIdx.x < 32) {
    const u32 dst = static_cast<u32>(__cvta_generic_to_shared(slot_in_smem));
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  return *slot_in_smem;
}
DEV void tmem_free_cta(u32 base, int ncols) {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(ncols) : "memory");
}
#endif

}  // namespace hhe
