// ff.cuh -- BN254 prime-field arithmetic for sm_100a, 8 x 32-bit limbs in registers.
//
// Replaces halo2curves 0.3.1 bn256::{Fr,Fq} (src/bn256/{fr,fq}.rs, src/derive/field.rs;
// pinned at Cargo.lock:1329-1331) on the device.  The memory format is unchanged:
// 4 x u64 little-endian limbs in Montgomery form with R = 2^256, which is bit-identical
// to 8 x u32 little-endian limbs, so host buffers are used as they are.
//
// Internal representation is LAZY: every value is kept in [0, 2p).  Because p < 2^254,
// 4p < R and a Montgomery product of two values < 2p is again < 2p without a final
// subtraction.  `reduce()` brings a value to the canonical [0, p) range; every store to
// memory that leaves the library goes through it, so outputs match halo2curves bit for bit.
//
// The multiplier is an interleaved (CIOS) Montgomery multiplication on two staggered
// accumulators: products a[j]*b[i] with even j are added into X (aligned to limb 0) and
// those with odd j into Y (aligned to limb 1) so that every row is ONE uninterrupted
// carry chain of mad.lo.cc / madc.hi.cc pairs (ptxas pairs them into IMAD.WIDE with
// carry).  After each round the value is divided by 2^32, which makes X and Y swap roles.
// Cost: 128 wide multiply-adds + 8 low multiplies + ~30 adds per multiplication.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_HD_NOINL __host__ __device__
#else
#define ZK_HD inline
#define ZK_HD_NOINL inline
#endif

namespace zk {

// ---- carry-chain primitives --------------------------------------------------------------
// Each primitive is ONE inline-asm block on the device (the carry flag never has to
// survive between asm statements) and a plain 64-bit emulation on the host.  The host
// versions exist so that the limb-level algorithms can be unit-tested without a GPU
// (tests/test_ff_host.py); they are not used by any product path.

// r = a + b (8 limbs), carry out dropped (callers guarantee a + b < 2^256)
ZK_HD void add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
#endif
}
// r = a - b (8 limbs); returns an all-ones mask when a < b, else 0
ZK_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t bw;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(bw)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t t = (uint64_t)a[i] - b[i] - c;
    r[i] = (uint32_t)t; c = (t >> 32) & 1;
  }
  bw = c ? 0xffffffffu : 0u;
#endif
  return bw;
}
// out[0..7] = {a0,a2,a4,a6} * b as four 64-bit products
ZK_HD void mul_row(uint32_t* o, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;"
      : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
  const uint32_t a[4] = {a0, a2, a4, a6};
  for (int j = 0; j < 4; j++) {
    uint64_t p = (uint64_t)a[j] * b;
    o[2 * j] = (uint32_t)p; o[2 * j + 1] = (uint32_t)(p >> 32);
  }
#endif
}
// acc[0..7] += {a0,a2,a4,a6} * b as one carry chain; the carry out is added into `top`
ZK_HD void mad_row(uint32_t* acc, uint32_t& top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
  const uint32_t a[4] = {a0, a2, a4, a6};
  uint64_t c = 0;
  for (int j = 0; j < 4; j++) {
    uint64_t p = (uint64_t)a[j] * b;
    c += (uint64_t)acc[2 * j] + (uint32_t)p; acc[2 * j] = (uint32_t)c; c >>= 32;
    c += (uint64_t)acc[2 * j + 1] + (uint32_t)(p >> 32); acc[2 * j + 1] = (uint32_t)c; c >>= 32;
  }
  top += (uint32_t)c;
#endif
}
// same without a carry sink (the caller knows the chain cannot overflow)
ZK_HD void mad_row_nc(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
  uint32_t sink = 0;
  mad_row(acc, sink, a0, a2, a4, a6, b);
#endif
}
// The /2^32 step of the interleaved multiplier.  xn0 += x[1] (x[1] is the limb that lands
// on position 0), and, carrying on from that addition, yn[0..7] = {x[2..7],0,0} +
// {a1,a3,a5,a7} * b.
ZK_HD void shift_mad_row(uint32_t& xn0, uint32_t* yn, const uint32_t* x, uint32_t a1, uint32_t a3,
                         uint32_t a5, uint32_t a7, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "madc.lo.cc.u32 %1, %16, %20, %10;\n\t"
      "madc.hi.cc.u32 %2, %16, %20, %11;\n\t"
      "madc.lo.cc.u32 %3, %17, %20, %12;\n\t"
      "madc.hi.cc.u32 %4, %17, %20, %13;\n\t"
      "madc.lo.cc.u32 %5, %18, %20, %14;\n\t"
      "madc.hi.cc.u32 %6, %18, %20, %15;\n\t"
      "madc.lo.cc.u32 %7, %19, %20, 0;\n\t"
      "madc.hi.u32 %8, %19, %20, 0;"
      : "+r"(xn0), "=r"(yn[0]), "=r"(yn[1]), "=r"(yn[2]), "=r"(yn[3]), "=r"(yn[4]), "=r"(yn[5]), "=r"(yn[6]), "=r"(yn[7])
      : "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
#else
  uint64_t c = (uint64_t)xn0 + x[1];
  xn0 = (uint32_t)c; c >>= 32;
  const uint32_t a[4] = {a1, a3, a5, a7};
  for (int j = 0; j < 4; j++) {
    uint64_t p = (uint64_t)a[j] * b;
    uint32_t s0 = (2 * j + 2 < 8) ? x[2 * j + 2] : 0, s1 = (2 * j + 3 < 8) ? x[2 * j + 3] : 0;
    c += (uint64_t)s0 + (uint32_t)p; yn[2 * j] = (uint32_t)c; c >>= 32;
    c += (uint64_t)s1 + (uint32_t)(p >> 32); yn[2 * j + 1] = (uint32_t)c; c >>= 32;
  }
#endif
}
// r[0..7] = y[0..7] + {x[1..7], 0}
ZK_HD void merge_shift(uint32_t* r, const uint32_t* y, const uint32_t* x) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]),
        "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)y[i] + (i < 7 ? x[i + 1] : 0); r[i] = (uint32_t)c; c >>= 32; }
#endif
}

// ---- primitives of the dedicated squaring ---------------------------------------------------------------------------------
// N products {a[0..N-1]} * b added into N consecutive 64-bit slots acc[0..2N-1] as one carry chain; the carry out is ADDED
// into acc[2N] (the caller guarantees that limb holds at most a few earlier carries, so it cannot overflow)
template <int N>
ZK_HD void mad_slots(uint32_t* acc, const uint32_t* a, uint32_t b) {
#ifdef __CUDA_ARCH__
  if (N == 1) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t addc.u32 %2, %2, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(a[0]), "r"(b));
  } else if (N == 2) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t addc.u32 %4, %4, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]) : "r"(a[0]), "r"(a[2]), "r"(b));
  } else if (N == 3) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t madc.hi.cc.u32 %5, %9, %10, %5;\n\t addc.u32 %6, %6, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(b));
  } else {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
  }
#else
  uint64_t c = 0;
  for (int j = 0; j < N; j++) {
    uint64_t p = (uint64_t)a[2 * j] * b;
    c += (uint64_t)acc[2 * j] + (uint32_t)p; acc[2 * j] = (uint32_t)c; c >>= 32;
    c += (uint64_t)acc[2 * j + 1] + (uint32_t)(p >> 32); acc[2 * j + 1] = (uint32_t)c; c >>= 32;
  }
  acc[2 * N] += (uint32_t)c;
#endif
}
// t[0..15] = 2 * t[0..15] + sum_i a[i]^2 * 2^(64 i)   (the doubling as funnel shifts, the diagonal as one carry chain)
ZK_HD void dbl_add_diag(uint32_t* t, const uint32_t* a) {
#pragma unroll
  for (int i = 15; i > 0; i--) t[i] = (t[i] << 1) | (t[i - 1] >> 31);
  t[0] <<= 1;
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t madc.hi.u32 %15, %23, %23, %15;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]),
        "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t p = (uint64_t)a[i] * a[i];
    c += (uint64_t)t[2 * i] + (uint32_t)p; t[2 * i] = (uint32_t)c; c >>= 32;
    c += (uint64_t)t[2 * i + 1] + (uint32_t)(p >> 32); t[2 * i + 1] = (uint32_t)c; c >>= 32;
  }
#endif
}
// r[1..15] = e[1..15] + o[0..14] (one chain; r[0] = e[0]); the sum is known to fit 512 bits
ZK_HD void merge_even_odd(uint32_t* e, const uint32_t* o) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %0, %15;\n\t addc.cc.u32 %1, %1, %16;\n\t addc.cc.u32 %2, %2, %17;\n\t addc.cc.u32 %3, %3, %18;\n\t"
      "addc.cc.u32 %4, %4, %19;\n\t addc.cc.u32 %5, %5, %20;\n\t addc.cc.u32 %6, %6, %21;\n\t addc.cc.u32 %7, %7, %22;\n\t"
      "addc.cc.u32 %8, %8, %23;\n\t addc.cc.u32 %9, %9, %24;\n\t addc.cc.u32 %10, %10, %25;\n\t addc.cc.u32 %11, %11, %26;\n\t"
      "addc.cc.u32 %12, %12, %27;\n\t addc.cc.u32 %13, %13, %28;\n\t addc.u32 %14, %14, %29;"
      : "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]),
        "+r"(e[9]), "+r"(e[10]), "+r"(e[11]), "+r"(e[12]), "+r"(e[13]), "+r"(e[14]), "+r"(e[15])
      : "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
        "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]));
#else
  uint64_t c = 0;
  for (int i = 1; i < 16; i++) { c += (uint64_t)e[i] + o[i - 1]; e[i] = (uint32_t)c; c >>= 32; }
#endif
}

struct FrParams {
  static constexpr uint32_t M0 = 0xf0000001u, M1 = 0x43e1f593u, M2 = 0x79b97091u, M3 = 0x2833e848u,
                            M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xefffffffu;   // -r^-1 mod 2^32
  // R mod r (Montgomery one)
  static constexpr uint32_t R0 = 0x4ffffffbu, R1 = 0xac96341cu, R2_ = 0x9f60cd29u, R3_ = 0x36fc7695u,
                            R4 = 0x7879462eu, R5 = 0x666ea36fu, R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
  // R^2 mod r
  static constexpr uint32_t S0 = 0xae216da7u, S1 = 0x1bb8e645u, S2 = 0xe35c59e3u, S3 = 0x53fe3ab1u,
                            S4 = 0x53bb8085u, S5 = 0x8c49833du, S6 = 0x7f4e44a5u, S7 = 0x0216d0b1u;
};
struct FqParams {
  static constexpr uint32_t M0 = 0xd87cfd47u, M1 = 0x3c208c16u, M2 = 0x6871ca8du, M3 = 0x97816a91u,
                            M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xe4866389u;
  static constexpr uint32_t R0 = 0xc58f0d9du, R1 = 0xd35d438du, R2_ = 0xf5c70b3du, R3_ = 0x0a78eb28u,
                            R4 = 0x7879462cu, R5 = 0x666ea36fu, R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
  static constexpr uint32_t S0 = 0x538afa89u, S1 = 0xf32cfc5bu, S2 = 0xd44501fbu, S3 = 0xb5e71911u,
                            S4 = 0x0a417ff6u, S5 = 0x47ab1effu, S6 = 0xcab8351fu, S7 = 0x06d89f71u;
};

template <class P>
struct Fp {
  uint32_t l[8];

  // ---- constants -------------------------------------------------------------------
  ZK_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
  }
  ZK_HD static Fp one() {
    Fp r;
    r.l[0] = P::R0; r.l[1] = P::R1; r.l[2] = P::R2_; r.l[3] = P::R3_;
    r.l[4] = P::R4; r.l[5] = P::R5; r.l[6] = P::R6; r.l[7] = P::R7;
    return r;
  }
  ZK_HD static Fp r2() {
    Fp r;
    r.l[0] = P::S0; r.l[1] = P::S1; r.l[2] = P::S2; r.l[3] = P::S3;
    r.l[4] = P::S4; r.l[5] = P::S5; r.l[6] = P::S6; r.l[7] = P::S7;
    return r;
  }
  // k*p for k = 1, 2 as a limb vector
  template <int K>
  ZK_HD static Fp kp() {
    Fp r;
    if (K == 1) {
      r.l[0] = P::M0; r.l[1] = P::M1; r.l[2] = P::M2; r.l[3] = P::M3;
      r.l[4] = P::M4; r.l[5] = P::M5; r.l[6] = P::M6; r.l[7] = P::M7;
    } else {
      r.l[0] = P::M0 << 1;
      r.l[1] = (P::M1 << 1) | (P::M0 >> 31); r.l[2] = (P::M2 << 1) | (P::M1 >> 31);
      r.l[3] = (P::M3 << 1) | (P::M2 >> 31); r.l[4] = (P::M4 << 1) | (P::M3 >> 31);
      r.l[5] = (P::M5 << 1) | (P::M4 >> 31); r.l[6] = (P::M6 << 1) | (P::M5 >> 31);
      r.l[7] = (P::M7 << 1) | (P::M6 >> 31);
    }
    return r;
  }

  // ---- memory ------------------------------------------------------------------------
  // 32-byte element = two 16-byte vector accesses
  ZK_HD static Fp load(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fp r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
  }
  // store as is (caller guarantees the range it wants)
  ZK_HD void store_raw(void* p) const {
    uint4* q = reinterpret_cast<uint4*>(p);
    uint4 a, b;
    a.x = l[0]; a.y = l[1]; a.z = l[2]; a.w = l[3];
    b.x = l[4]; b.y = l[5]; b.z = l[6]; b.w = l[7];
    q[0] = a; q[1] = b;
  }
  // canonical store
  ZK_HD void store(void* p) const { reduce().store_raw(p); }

  // ---- range handling ------------------------------------------------------------------
  // [0,2p) -> [0,p)
  ZK_HD Fp reduce() const {
    Fp t, m = kp<1>();
    uint32_t bw = sub8(t.l, l, m.l);
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = bw ? l[i] : t.l[i];
    return r;
  }
  ZK_HD bool is_zero() const {
    Fp r = reduce();
    return (r.l[0] | r.l[1] | r.l[2] | r.l[3] | r.l[4] | r.l[5] | r.l[6] | r.l[7]) == 0;
  }
  // raw bit test, valid for values known to be canonical (e.g. straight from memory)
  ZK_HD bool is_zero_raw() const { return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7]) == 0; }

  // ---- add / sub ---------------------------------------------------------------------
  ZK_HD friend Fp operator+(const Fp& a, const Fp& b) {
    Fp s, t, m = kp<2>();
    add8(s.l, a.l, b.l);              // < 4p < 2^256
    uint32_t bw = sub8(t.l, s.l, m.l);
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = bw ? s.l[i] : t.l[i];
    return r;
  }
  ZK_HD friend Fp operator-(const Fp& a, const Fp& b) {
    Fp d, m = kp<2>();
    uint32_t bw = sub8(d.l, a.l, b.l);
#pragma unroll
    for (int i = 0; i < 8; i++) m.l[i] &= bw;   // add 2p back only when a < b
    Fp r;
    add8(r.l, d.l, m.l);
    return r;
  }
  ZK_HD Fp neg() const { return zero() - *this; }
  ZK_HD Fp dbl() const { return *this + *this; }

  // ---- Montgomery multiplication -------------------------------------------------------
  // one reduction round on (X aligned at limb 0, Y aligned at limb 1): adds m*p with
  // m = X[0] * INV, which clears X[0]
  ZK_HD static void redc_round(uint32_t* X, uint32_t* Y) {
    uint32_t m = X[0] * P::INV;
    mad_row(X, Y[7], P::M0, P::M2, P::M4, P::M6, m);
    mad_row_nc(Y, P::M1, P::M3, P::M5, P::M7, m);
  }
  ZK_HD friend Fp operator*(const Fp& a, const Fp& b) {
    uint32_t X[8], Y[8];
    mul_row(X, a.l[0], a.l[2], a.l[4], a.l[6], b.l[0]);
    mul_row(Y, a.l[1], a.l[3], a.l[5], a.l[7], b.l[0]);
    redc_round(X, Y);
#pragma unroll
    for (int i = 1; i < 8; i++) {
      // divide by 2^32: Y becomes the limb-0 aligned accumulator; X[1] is folded into it and
      // X[2..7] seed the new limb-1 aligned accumulator, which takes the odd products.
      uint32_t Xn[8], Yn[8];
#pragma unroll
      for (int j = 0; j < 8; j++) Xn[j] = Y[j];
      shift_mad_row(Xn[0], Yn, X, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
      mad_row(Xn, Yn[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
      redc_round(Xn, Yn);
#pragma unroll
      for (int j = 0; j < 8; j++) { X[j] = Xn[j]; Y[j] = Yn[j]; }
    }
    Fp r;
    merge_shift(r.l, Y, X);   // X[0] == 0; result = (X >> 32) + Y  < 2p
    return r;
  }
  // Dedicated squaring: the 28 products a[i] * a[j], i < j, are formed once (as carry chains over the even- and the
  // odd-position product slots), doubled, and the 8 squares a[i]^2 added as one more chain: 36 wide multiplications instead
  // of 64 for the operand products; the Montgomery reduction (64 more) then runs on the low half of the 512-bit square with
  // the same staggered accumulators as operator*, and the high half is added at the end.  Result < 2p as for operator*
  // ((a^2 + m p) / R with a < 2p, m < R).
  ZK_HD Fp sqr() const {
    const uint32_t* a = l;
    uint32_t E[17], O[17];
#pragma unroll
    for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
    // E[p]: products whose low limb sits at the even position p; O[p]: at the odd position p + 1.  Row i multiplies a[i]
    // with a[j], j > i; the j of one parity give consecutive slots.  Carries land on limbs no later row's carry can overflow.
    mad_slots<3>(E + 2, a + 2, a[0]);      // a0 * {a2, a4, a6}      -> positions 2..7,   carry -> 8
    mad_slots<4>(O + 0, a + 1, a[0]);      // a0 * {a1, a3, a5, a7}  -> positions 1..8,   carry -> 9
    mad_slots<3>(E + 4, a + 3, a[1]);      // a1 * {a3, a5, a7}      -> positions 4..9,   carry -> 10
    mad_slots<3>(O + 2, a + 2, a[1]);      // a1 * {a2, a4, a6}      -> positions 3..8,   carry -> 9
    mad_slots<2>(E + 6, a + 4, a[2]);      // a2 * {a4, a6}          -> positions 6..9,   carry -> 10
    mad_slots<3>(O + 4, a + 3, a[2]);      // a2 * {a3, a5, a7}      -> positions 5..10,  carry -> 11
    mad_slots<2>(E + 8, a + 5, a[3]);      // a3 * {a5, a7}          -> positions 8..11,  carry -> 12
    mad_slots<2>(O + 6, a + 4, a[3]);      // a3 * {a4, a6}          -> positions 7..10,  carry -> 11
    mad_slots<1>(E + 10, a + 6, a[4]);     // a4 * a6                -> positions 10..11, carry -> 12
    mad_slots<2>(O + 8, a + 5, a[4]);      // a4 * {a5, a7}          -> positions 9..12,  carry -> 13
    mad_slots<1>(E + 12, a + 7, a[5]);     // a5 * a7                -> positions 12..13, carry -> 14
    mad_slots<1>(O + 10, a + 6, a[5]);     // a5 * a6                -> positions 11..12, carry -> 13
    mad_slots<1>(O + 12, a + 7, a[6]);     // a6 * a7                -> positions 13..14, carry -> 15
    merge_even_odd(E, O);                  // E = sum of the off-diagonal products
    dbl_add_diag(E, a);                    // E = a^2 (512 bits)
    // Montgomery reduction of the low 256 bits; X aligned at limb 0, Y at limb 1 as in operator*
    uint32_t X[8], Y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) X[j] = E[j];
    {
      uint32_t m = X[0] * P::INV, top = 0;
      mad_row(X, top, P::M0, P::M2, P::M4, P::M6, m);
      mul_row(Y, P::M1, P::M3, P::M5, P::M7, m);
      // the carry out of X's row has weight 2^256 = limb 7 of Y (Y is aligned at limb 1)
      Y[7] += top;
    }
#pragma unroll
    for (int i = 1; i < 8; i++) {
      uint32_t Xn[8], Yn[8];
#pragma unroll
      for (int j = 0; j < 8; j++) Xn[j] = Y[j];
      const uint32_t m = (Y[0] + X[1]) * P::INV;
      shift_mad_row(Xn[0], Yn, X, P::M1, P::M3, P::M5, P::M7, m);
      mad_row(Xn, Yn[7], P::M0, P::M2, P::M4, P::M6, m);
#pragma unroll
      for (int j = 0; j < 8; j++) { X[j] = Xn[j]; Y[j] = Yn[j]; }
    }
    Fp r, hi;
    merge_shift(r.l, Y, X);
#pragma unroll
    for (int j = 0; j < 8; j++) hi.l[j] = E[8 + j];
    Fp out;
    add8(out.l, r.l, hi.l);
    return out;
  }

  // Montgomery form -> canonical integer (multiply by 1 = one REDC pass), fully reduced
  ZK_HD Fp from_mont() const {
    Fp o = zero();
    o.l[0] = 1;
    return ((*this) * o).reduce();
  }
  ZK_HD Fp to_mont() const { return (*this) * r2(); }

  // a^e, 256-bit exponent as 8 limbs
  ZK_HD_NOINL static Fp pow(const Fp& a, const uint32_t* e) {
    Fp acc = one();
    for (int i = 255; i >= 0; i--) {
      acc = acc.sqr();
      if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * a;
    }
    return acc;
  }
  ZK_HD_NOINL static Fp pow_u64(const Fp& a, uint64_t e) {
    if (!e) return one();
    int top = 63;
    while (!((e >> top) & 1)) top--;          // start at the leading one: no squarings of 1 (most exponents here are < 2^21)
    Fp acc = a;
    for (int i = top - 1; i >= 0; i--) {
      acc = acc.sqr();
      if ((e >> i) & 1) acc = acc * a;
    }
    return acc;
  }
  // Fermat inversion a^(p-2); inv(0) = 0
  ZK_HD_NOINL static Fp inv(const Fp& a) {
    const uint32_t e[8] = {P::M0 - 2, P::M1, P::M2, P::M3, P::M4, P::M5, P::M6, P::M7};
    return pow(a, e);
  }
  ZK_HD friend bool operator==(const Fp& a, const Fp& b) { return (a - b).is_zero(); }
  ZK_HD friend bool operator!=(const Fp& a, const Fp& b) { return !(a == b); }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace zk
