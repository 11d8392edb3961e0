// probe_dfma_modmul.cu -- feasibility probe (not part of the product): a 254-bit Montgomery multiplication on
// the FP64 pipe (5 x 52-bit limbs held in doubles, products split into high / low halves with two round-toward-zero
// FMAs each, after Emmart / Zheng / Weems 2018) next to the integer (IMAD.WIDE) multiplier of csrc/ff.cuh:
//   1. correctness of the DFMA multiplier against the integer one on random inputs;
//   2. throughput of each alone and of a kernel whose warps alternate between the two (the pipes are separate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I zk-dcap-verifier_b200/csrc -o tools/probe_dfma_modmul tools/probe_dfma_modmul.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ff.cuh"

using namespace zk;

// Fq in 52-bit limbs
__device__ __constant__ double kP52[5] = {(double)0x8c16d87cfd47ULL, (double)0x916871ca8d3c2ULL, (double)0x181585d97816aULL,
                                          (double)0xa029b85045b68ULL, (double)0x30644e72e131ULL};
#define NP52 ((double)0x20782e4866389ULL)      // -p^-1 mod 2^52

struct F52 {
  double l[5];
};

__device__ __forceinline__ double two52() { return 4503599627370496.0; }                       // 2^52
__device__ __forceinline__ double two104() { return 20282409603651670423947251286016.0; }      // 2^104
__device__ __forceinline__ double two104p52() { return 20282409603651674927546878656512.0; }   // 2^104 + 2^52

constexpr uint64_t kHiOff = 0x467ULL << 52, kLoOff = 0x433ULL << 52, kMask52 = (1ULL << 52) - 1;

// (hi, lo) bit patterns of a*b: hi = floor(ab / 2^52) (+ kHiOff), lo = ab mod 2^52 (+ kLoOff)
__device__ __forceinline__ void mul52(double a, double b, uint64_t& hi, uint64_t& lo) {
  double ph = __fma_rz(a, b, two104());
  double sub = two104p52() - ph;
  double pl = __fma_rz(a, b, sub);
  hi = (uint64_t)__double_as_longlong(ph);
  lo = (uint64_t)__double_as_longlong(pl);
}
__device__ __forceinline__ double u52_to_double(uint64_t v) {     // v < 2^52
  return __longlong_as_double((long long)(v | kLoOff)) - two52();
}

// Montgomery product a * b * 2^-260 mod p, limbs normalised to [0, 2^52), value < 2p for inputs < 2p
__device__ __forceinline__ F52 mont52(const F52& a, const F52& b) {
  uint64_t col[10];
  // every column starts at minus the offsets that the PRODUCT phase adds to it
#pragma unroll
  for (int k = 0; k < 10; k++) {
    int nlo = (k <= 4) ? k + 1 : 9 - k;                  // pairs (i, j) with i + j = k
    int nhi = (k >= 1) ? ((k - 1 <= 4) ? k : 10 - k) : 0; // pairs with i + j = k - 1
    if (k > 9) nlo = 0;
    col[k] = 0 - ((uint64_t)nlo * kLoOff + (uint64_t)nhi * kHiOff);
  }
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) {
      uint64_t hi, lo;
      mul52(a.l[i], b.l[j], hi, lo);
      col[i + j] += lo;
      col[i + j + 1] += hi;
    }
  // word-serial reduction
#pragma unroll
  for (int i = 0; i < 5; i++) {
    double t = u52_to_double(col[i] & kMask52);
    uint64_t qh, ql;
    mul52(t, NP52, qh, ql);
    double q = __longlong_as_double((long long)ql) - two52();      // (t * np) mod 2^52
#pragma unroll
    for (int j = 0; j < 5; j++) {
      uint64_t hi, lo;
      mul52(q, kP52[j], hi, lo);
      col[i + j] += lo - kLoOff;
      col[i + j + 1] += hi - kHiOff;
    }
    col[i + 1] += col[i] >> 52;        // low 52 bits are zero now
  }
  F52 r;
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    uint64_t v = col[5 + k] + carry;
    carry = v >> 52;
    r.l[k] = u52_to_double(v & kMask52);
  }
  return r;
}

// 8 x u32 (value < 2^256) <-> 5 x 52-bit limbs
__device__ __forceinline__ F52 to52(const Fq& a) {
  uint64_t w[4];
#pragma unroll
  for (int i = 0; i < 4; i++) w[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
  F52 r;
  r.l[0] = u52_to_double(w[0] & kMask52);
  r.l[1] = u52_to_double(((w[0] >> 52) | (w[1] << 12)) & kMask52);
  r.l[2] = u52_to_double(((w[1] >> 40) | (w[2] << 24)) & kMask52);
  r.l[3] = u52_to_double(((w[2] >> 28) | (w[3] << 36)) & kMask52);
  r.l[4] = u52_to_double(w[3] >> 16);
  return r;
}
__device__ __forceinline__ Fq from52(const F52& a) {
  uint64_t v[5];
#pragma unroll
  for (int i = 0; i < 5; i++) v[i] = (uint64_t)__double_as_longlong(a.l[i] + two52()) & kMask52;
  uint64_t w[4];
  w[0] = v[0] | (v[1] << 52);
  w[1] = (v[1] >> 12) | (v[2] << 40);
  w[2] = (v[2] >> 24) | (v[3] << 28);
  w[3] = (v[3] >> 36) | (v[4] << 16);
  Fq r;
#pragma unroll
  for (int i = 0; i < 4; i++) { r.l[2 * i] = (uint32_t)w[i]; r.l[2 * i + 1] = (uint32_t)(w[i] >> 32); }
  return r;
}

// ---- correctness: out[i] = mont52(a, b) converted back, next to the integer product scaled to the same Montgomery radix
__global__ void k_check(const uint4* a, const uint4* b, uint4* out52, uint4* outint, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = Fq::load(a + 2 * i), y = Fq::load(b + 2 * i);
  from52(mont52(to52(x), to52(y))).reduce().store_raw(out52 + 2 * i);
  // integer: x*y*2^-256; the DFMA result is x*y*2^-260 = (x*y*2^-256) * 2^-4: multiply the integer result by 2^252 * 2^-256
  Fq s = Fq::zero(); s.l[7] = 1u << 28;        // 2^252
  ((x * y) * s).reduce().store_raw(outint + 2 * i);
}

template <int MODE>   // 0 integer, 1 dfma, 2 warps alternate
__global__ void __launch_bounds__(256) k_bench(uint32_t iters, uint32_t* sink) {
  bool use52 = MODE == 1 || (MODE == 2 && ((threadIdx.x >> 5) & 1));
  Fq a = Fq::one(), b = Fq::r2(), c = Fq::one() + Fq::r2(), d = Fq::r2().dbl();
  a.l[0] ^= threadIdx.x; b.l[1] ^= blockIdx.x;
  a = a.reduce(); b = b.reduce(); c = c.reduce(); d = d.reduce();
  uint32_t acc;
  if (use52) {
    F52 A = to52(a), B = to52(b), Cc = to52(c), D = to52(d);
    for (uint32_t i = 0; i < iters; i++) { A = mont52(A, B); B = mont52(B, Cc); Cc = mont52(Cc, D); D = mont52(D, A); }
    Fq s = from52(A) + from52(B) + from52(Cc) + from52(D);
    acc = s.l[0] ^ s.l[7];
  } else {
    for (uint32_t i = 0; i < iters; i++) { a = a * b; b = b * c; c = c * d; d = d * a; }
    Fq s = a + b + c + d;
    acc = s.l[0] ^ s.l[7];
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE>
static double run(uint32_t* sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t iters = 2000;
  const unsigned grid = 148 * 8, block = 256;
  k_bench<MODE><<<grid, block>>>(iters, sink);
  cudaEventRecord(e0);
  k_bench<MODE><<<grid, block>>>(iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double)grid * block * 4.0 * iters / (ms * 1e-3) / 1e9;
}

int main() {
  const size_t n = 1 << 16;
  uint32_t* h = (uint32_t*)malloc(n * 32 * 2);
  uint64_t s = 0x9E3779B97F4A7C15ULL;
  for (size_t i = 0; i < n * 16; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s >> 16); }
  for (size_t i = 0; i < 2 * n; i++) h[8 * i + 7] &= 0x1FFFFFFFu;     // < 2^253 < p
  uint4 *da, *db, *o1, *o2; uint32_t* sink;
  cudaMalloc(&da, n * 32); cudaMalloc(&db, n * 32); cudaMalloc(&o1, n * 32); cudaMalloc(&o2, n * 32); cudaMalloc(&sink, 256);
  cudaMemcpy(da, h, n * 32, cudaMemcpyHostToDevice); cudaMemcpy(db, h + 8 * n, n * 32, cudaMemcpyHostToDevice);
  k_check<<<(unsigned)((n + 127) / 128), 128>>>(da, db, o1, o2, n);
  uint32_t* r1 = (uint32_t*)malloc(n * 32); uint32_t* r2 = (uint32_t*)malloc(n * 32);
  cudaMemcpy(r1, o1, n * 32, cudaMemcpyDeviceToHost); cudaMemcpy(r2, o2, n * 32, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (size_t i = 0; i < n; i++) if (memcmp(r1 + 8 * i, r2 + 8 * i, 32)) bad++;
  printf("{\"probe\": \"dfma_modmul_check\", \"n\": %zu, \"mismatches\": %zu}\n", n, bad);
  if (bad) {
    for (size_t i = 0; i < n && bad; i++) if (memcmp(r1 + 8 * i, r2 + 8 * i, 32)) {
      printf("first mismatch at %zu: ", i);
      for (int k = 7; k >= 0; k--) printf("%08x", r1[8 * i + k]);
      printf(" vs ");
      for (int k = 7; k >= 0; k--) printf("%08x", r2[8 * i + k]);
      printf("\n"); break;
    }
  }
  double gi = run<0>(sink), gd = run<1>(sink), gm = run<2>(sink);
  printf("{\"probe\": \"modmul_G_per_s\", \"imad\": %.2f, \"dfma\": %.2f, \"warps_alternate\": %.2f}\n", gi, gd, gm);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
