// probe_pipes.cu -- pipe micro-benchmarks on the GPU box (not part of the product):
// DFMA rate, IMAD.WIDE rate, IADD3 rate, and whether DFMA / IADD3 issue concurrently with IMAD.WIDE.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_pipes tools/probe_pipes.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// MODE bit 0: IMAD.WIDE chains, bit 1: DFMA chains, bit 2: IADD3 (add.cc chains)
template <int MODE>
__global__ void __launch_bounds__(256) k_probe(uint32_t iters, uint32_t* sink) {
  uint64_t y0 = threadIdx.x, y1 = blockIdx.x, y2 = 3, y3 = 5;
  uint32_t m = threadIdx.x * 2654435761u + 1;
  double d0 = threadIdx.x * 1.0, d1 = 1.5, d2 = 2.5, d3 = 3.5, dm = 1.0000001, dc = 0.5;
  uint32_t a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE & 1) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y0) : "r"((uint32_t)y1), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y1) : "r"((uint32_t)y2), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y2) : "r"((uint32_t)y3), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y3) : "r"((uint32_t)y0), "r"(m));
      }
      if (MODE & 2) {
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d0) : "d"(dm), "d"(dc));
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d1) : "d"(dm), "d"(dc));
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d2) : "d"(dm), "d"(dc));
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d3) : "d"(dm), "d"(dc));
      }
      if (MODE & 4) {
        asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;"
                     : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3) : "r"(a4), "r"(a5), "r"(a6), "r"(a7));
        asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;"
                     : "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(a0), "r"(a1), "r"(a2), "r"(a3));
      }
    }
  }
  uint32_t s = (uint32_t)(y0 ^ y1 ^ y2 ^ y3) ^ (uint32_t)__double_as_longlong(d0 + d1 + d2 + d3) ^ a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (s == 0x13572468u) sink[0] = s;
}

template <int MODE>
static void run(const char* name, uint32_t* sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t iters = 4000;
  const unsigned grid = 148 * 8, block = 256;
  k_probe<MODE><<<grid, block>>>(iters, sink);
  cudaEventRecord(e0);
  k_probe<MODE><<<grid, block>>>(iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double threads = (double)grid * block, per = 8.0 * iters;
  printf("{\"probe\": \"%s\", \"ms\": %.3f", name, ms);
  if (MODE & 1) printf(", \"imad_wide_T_per_s\": %.3f", threads * per * 4 / (ms * 1e-3) / 1e12);
  if (MODE & 2) printf(", \"dfma_T_per_s\": %.3f", threads * per * 4 / (ms * 1e-3) / 1e12);
  if (MODE & 4) printf(", \"iadd_T_per_s\": %.3f", threads * per * 8 / (ms * 1e-3) / 1e12);
  printf("}\n");
}

int main() {
  uint32_t* sink;
  cudaMalloc(&sink, 256);
  run<1>("imad_wide", sink);
  run<2>("dfma", sink);
  run<4>("iadd3", sink);
  run<3>("imad_wide+dfma", sink);
  run<5>("imad_wide+iadd3", sink);
  run<6>("dfma+iadd3", sink);
  run<7>("all", sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
