// Pipe-throughput microbenchmarks for sm_100a (B200): which units can the exact modular arithmetic of the BFV engine use
// concurrently?  Each kernel runs ITER iterations of an unrolled body of independent chains; results are reported as
// warp-instructions per clock per SM sub-partition (a unit that accepts one warp instruction every c cycles shows 1/c).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int CH = 8;  // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(512) bench(double *out, double seed, unsigned long long iseed) {
  double x[CH], y = seed * 1.0000001, z = seed * 0.5;
  unsigned long long u[CH];
  unsigned long long m = iseed | 1;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    x[c] = seed + c + threadIdx.x;
    u[c] = iseed + c * 77 + threadIdx.x;
  }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (MODE == 0) {  // DFMA only: 8 per chain
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(y), "d"(z));
      } else if (MODE == 1) {  // FRND only (cvt.rni.f64.f64): 2 per chain
#pragma unroll
        for (int k = 0; k < 2; ++k) asm volatile("cvt.rni.f64.f64 %0, %0;" : "+d"(x[c]));
      } else if (MODE == 2) {  // 7 DFMA + 1 FRND
#pragma unroll
        for (int k = 0; k < 7; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(y), "d"(z));
        asm volatile("cvt.rni.f64.f64 %0, %0;" : "+d"(x[c]));
      } else if (MODE == 3) {  // 64-bit mul.hi only: 2 per chain
#pragma unroll
        for (int k = 0; k < 2; ++k) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(u[c]) : "l"(m));
      } else if (MODE == 4) {  // 8 DFMA + 1 mul.hi.u64 + 1 mul.lo.u64
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(y), "d"(z));
        asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(u[c]) : "l"(m));
        asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(u[c]) : "l"(m));
      } else if (MODE == 5) {  // 1 mul.hi.u64 + 1 mul.lo.u64 (integer half of mode 4)
        asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(u[c]) : "l"(m));
        asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(u[c]) : "l"(m));
      } else if (MODE == 6) {  // DADD only: 8 per chain
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(y));
      } else if (MODE == 7) {  // cvt.rni.s64.f64 + cvt.rn.f64.s64 round trip
        long long t;
        asm volatile("cvt.rni.s64.f64 %0, %1;" : "=l"(t) : "d"(x[c]));
        asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(x[c]) : "l"(t));
      } else if (MODE == 8) {  // 8 DFMA + 4 x 32-bit mad.wide.u32 (IMAD.WIDE)
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(y), "d"(z));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          unsigned lo = (unsigned)u[c];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(u[c]) : "r"(lo), "r"((unsigned)m));
        }
      } else if (MODE == 9) {  // 4 x mad.wide.u32 only
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          unsigned lo = (unsigned)u[c];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(u[c]) : "r"(lo), "r"((unsigned)m));
        }
      } else if (MODE == 10) {  // 8 DFMA + 8 shared-memory-free ALU ops (LOP3/IADD3) : issue-slot sharing check
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(y), "d"(z));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          unsigned lo = (unsigned)u[c];
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(lo) : "r"((unsigned)m + k));
          u[c] = lo;
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c] + (double)u[c];
  if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char *name, double inst_per_chain_iter) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double *out;
  cudaMalloc(&out, 8);
  const int blocks = sms * 4, threads = 512;  // 64 warps per SM
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  bench<MODE><<<blocks, threads>>>(out, 1.5, 12345);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 1.5, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double warps_per_smsp = blocks * (threads / 32.0) / sms / 4.0;
  const double winst = warps_per_smsp * ITER * CH * inst_per_chain_iter;  // warp instructions per sub-partition
  const double clocks = best * 1e-3 * khz * 1e3;                          // at the nominal max clock
  printf("%-44s %8.3f ms  %7.4f warp-inst/clk/SMSP (at %d MHz nominal)  cycles per chain-iter per warp: %.2f\n", name, best,
         winst / clocks, khz / 1000, clocks / (warps_per_smsp * ITER * CH));
  cudaFree(out);
}

int main() {
  run<0>("DFMA x8", 8);
  run<6>("DADD x8", 8);
  run<1>("FRND.F64 x2", 2);
  run<2>("DFMA x7 + FRND x1", 8);
  run<3>("mul.hi.u64 x2", 2);
  run<5>("mul.hi.u64 + mul.lo.u64", 2);
  run<4>("DFMA x8 + mul.hi.u64 + mul.lo.u64", 10);
  run<9>("mad.wide.u32 x4", 4);
  run<8>("DFMA x8 + mad.wide.u32 x4", 12);
  run<7>("cvt f64->s64->f64", 2);
  run<10>("DFMA x8 + XOR x8", 16);
  return 0;
}
