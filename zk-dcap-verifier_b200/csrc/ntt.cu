// ntt.cu -- Fr radix-2 NTT for sm_100a and the EvaluationDomain transforms built on it.
//
// Replaces halo2_proofs arithmetic::best_fft and poly::EvaluationDomain::{lagrange_to_coeff,
// coeff_to_lagrange, coeff_to_extended, extended_to_coeff, divide_by_vanishing_poly}
// ([UPSTREAM] src/arithmetic.rs, src/poly/domain.rs @ c254c75f, Cargo.lock:1314-1316), reached
// from keygen_pk (circuits/src/sgx_dcap_verifier.rs:807) and create_proof (:814).
//
// Algorithm.  A 2^L-point transform is the L stages of a decimation-in-frequency radix-2
// network, executed in ceil(L/8) passes.  A pass takes the next t stages: the 2^t points that
// those stages connect are 2^(L-t) apart, so a CTA stages a tile of C adjacent columns x 2^t
// rows in shared memory (rows are C*32 B contiguous in HBM -> coalesced), runs the t stages
// there, and writes the tile back with the pass's digit bit-reversed and moved below the
// still-unprocessed bits.  After the last pass the array is in natural order: no separate
// bit-reversal pass ever touches HBM.  Each pass therefore moves exactly 64 B per point.
// All butterflies use the GLOBAL radix-2 twiddle omega^e read from one table of N/2 powers
// (L2 resident for the sizes of interest), so a transform costs (N/2)*L multiplications
// minus the trivial ones -- there are no extra inter-pass twiddle multiplications.
// The inverse direction reuses the same table through omega^-e = -omega^(N/2-e).
// Coset scaling (zeta^(i mod 3)), zero padding and the 1/N factor are fused into the first
// pass's load and the last pass's store.
#include <stdlib.h>

#include "common.cuh"
#include "ff.cuh"

using namespace zk;

namespace {

constexpr int kTileLogDefault = 10;   // 2^10 points = 32 KiB of shared memory per CTA (6 CTAs / SM)
constexpr int kMaxTDefault = 8;       // stages per pass
// tuning overrides (ZK_NTT_TILE_LOG in [6,12], ZK_NTT_MAX_T in [2,10]); read once
int env_int(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (!v) return dflt;
  int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}
const int kTileLog = env_int("ZK_NTT_TILE_LOG", kTileLogDefault, 6, 12);
const int kMaxT = env_int("ZK_NTT_MAX_T", kMaxTDefault, 2, 10) > kTileLog ? kTileLog
                                                                           : env_int("ZK_NTT_MAX_T", kMaxTDefault, 2, 10);
constexpr int kThreads = 256;
#ifndef ZK_NTT_UNROLL
#define ZK_NTT_UNROLL 1
#endif
constexpr int kNttUnroll = ZK_NTT_UNROLL;   // butterflies of one stage a thread works on together

struct PassConsts {
  Fr pre[3];    // multiplier for input index i, selected by i mod 3 (pre_mode 1)
  Fr post[3];   // multiplier for output index i, selected by i mod 3 (post_mode 2) or post[0] (post_mode 1)
};

struct PassArgs {
  const uint4* in;
  uint4* out;
  const uint4* tw;         // tw[e << tw_shift] = omega^e, e < N/2
  size_t in_stride;        // elements between consecutive polynomials of a batch (input)
  size_t out_stride;
  size_t n_in;             // input elements that exist; indices >= n_in read as zero
  uint32_t L, s0, t, logC, tw_shift;
  int inverse, pre_mode, post_mode;
};

__device__ __forceinline__ Fr lds_fr(const uint4* p0, const uint4* p1, uint32_t e) {
  uint4 a = p0[e], b = p1[e];
  Fr r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
  r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  return r;
}
__device__ __forceinline__ void sts_fr(uint4* p0, uint4* p1, uint32_t e, const Fr& v) {
  p0[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  p1[e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

#ifndef ZK_NTT_MIN_CTAS
#define ZK_NTT_MIN_CTAS 4
#endif
// one butterfly output of the DIF network: (x - y) * omega^ex, resp. (y - x) * omega^(N/2 - ex) = (x - y) * omega^-ex
__device__ __forceinline__ Fr bf_diff(const Fr& x, const Fr& y, size_t ex, const PassArgs& A, uint32_t halfN_sh) {
  if (ex == 0) return x - y;
  if (!A.inverse) return (x - y) * Fr::load(A.tw + 2 * (ex << A.tw_shift));
  return (y - x) * Fr::load(A.tw + 2 * ((((size_t)1 << halfN_sh) - ex) << A.tw_shift));
}

// R4: two stages per trip through shared memory (radix-4 butterflies in registers): half the LDS/STS traffic and
// half the barriers of the radix-2 loop; needs ~2x the registers, so 2 CTAs per SM instead of 4.
template <int R4>
#ifndef ZK_NTT_R4_CTAS
#define ZK_NTT_R4_CTAS 3
#endif
__global__ void __launch_bounds__(kThreads, R4 ? ZK_NTT_R4_CTAS : ZK_NTT_MIN_CTAS) k_ntt_pass_t(PassArgs A, PassConsts K) {
  extern __shared__ uint4 smem[];
  const uint32_t t = A.t, logC = A.logC, C = 1u << logC, L = A.L, P = A.s0;
  const uint32_t rows = 1u << t, tile = rows << logC;
  uint4* p0 = smem;                 // low 16 bytes of every element
  uint4* p1 = smem + tile + 4;      // high 16 bytes (+64 B so the two planes fall on different banks)
  const size_t rest0 = (size_t)blockIdx.x << logC;
  const size_t row_stride = (size_t)1 << (L - t);
  const uint4* in = A.in + 2 * (size_t)blockIdx.y * A.in_stride;
  uint4* out = A.out + 2 * (size_t)blockIdx.y * A.out_stride;

  // ---- load tile (coalesced rows of C*32 bytes) -------------------------------------------
  if (A.pre_mode == 0) {
    for (uint32_t q = threadIdx.x; q < 2 * tile; q += kThreads) {
      uint32_t m = q >> (logC + 1), off = q & (2 * C - 1);
      size_t elem = (size_t)m * row_stride + rest0 + (off >> 1);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (elem < A.n_in) v = in[2 * elem + (off & 1)];
      uint32_t e = (m << logC) + (off >> 1);
      if (off & 1) p1[e] = v; else p0[e] = v;
    }
  } else {
    for (uint32_t e = threadIdx.x; e < tile; e += kThreads) {
      uint32_t m = e >> logC, c = e & (C - 1);
      size_t elem = (size_t)m * row_stride + rest0 + c;
      Fr v = Fr::zero();
      if (elem < A.n_in) {
        v = Fr::load(in + 2 * elem);
        uint32_t r3 = (uint32_t)(elem % 3);
        if (r3 == 1) v = v * K.pre[1];
        else if (r3 == 2) v = v * K.pre[2];
      }
      sts_fr(p0, p1, e, v);
    }
  }
  __syncthreads();

  // ---- t butterfly stages in shared memory ---------------------------------------------------
  const uint32_t halfN_sh = L - 1;    // log2(N/2)
  uint32_t u = 1;
  if (R4) {
    for (; u + 1 <= t; u += 2) {
      const uint32_t s = P + u, hb = t - u, hb2 = hb - 1;   // partner distances 2^hb (stage u) and 2^hb2 (stage u+1) rows
      for (uint32_t b = threadIdx.x; b < (tile >> 2); b += kThreads) {
        uint32_t c = b & (C - 1), mm = b >> logC;
        uint32_t m_lo = mm & ((1u << hb2) - 1), m_hi = mm >> hb2;
        uint32_t m = (m_hi << (hb + 1)) | m_lo;               // row with bits hb and hb2 cleared
        uint32_t e00 = (m << logC) + c, e01 = e00 + (1u << (hb2 + logC)), e10 = e00 + (1u << (hb + logC)), e11 = e10 + (1u << (hb2 + logC));
        Fr x0 = lds_fr(p0, p1, e00), x1 = lds_fr(p0, p1, e01), y0 = lds_fr(p0, p1, e10), y1 = lds_fr(p0, p1, e11);
        size_t lo = (rest0 + c) >> P;
        size_t ja = ((size_t)m_lo << (L - P - t)) + lo, jb = ((size_t)(m_lo | (1u << hb2)) << (L - P - t)) + lo;
        Fr s0 = x0 + y0, d0 = bf_diff(x0, y0, ja << (s - 1), A, halfN_sh);
        Fr s1 = x1 + y1, d1 = bf_diff(x1, y1, jb << (s - 1), A, halfN_sh);
        size_t exc = ja << s;                                  // stage u+1: the row bits below hb2 are m_lo for all four
        sts_fr(p0, p1, e00, s0 + s1);
        sts_fr(p0, p1, e01, bf_diff(s0, s1, exc, A, halfN_sh));
        sts_fr(p0, p1, e10, d0 + d1);
        sts_fr(p0, p1, e11, bf_diff(d0, d1, exc, A, halfN_sh));
      }
      __syncthreads();
    }
  }
  for (; u <= t; u++) {
    const uint32_t s = P + u, hb = t - u;             // partner distance 2^hb rows
#pragma unroll kNttUnroll
    for (uint32_t b = threadIdx.x; b < (tile >> 1); b += kThreads) {
      uint32_t c = b & (C - 1), mm = b >> logC;
      uint32_t m_lo = mm & ((1u << hb) - 1), m_hi = mm >> hb;
      uint32_t m = (m_hi << (hb + 1)) | m_lo;
      uint32_t e0 = (m << logC) + c, e1 = e0 + (1u << (hb + logC));
      Fr x = lds_fr(p0, p1, e0), y = lds_fr(p0, p1, e1);
      // logical index bits below this pass's digit = the still-unprocessed bits of the column
      size_t lo = (rest0 + c) >> P;
      size_t j = ((size_t)m_lo << (L - P - t)) + lo;
      Fr sum = x + y, d = bf_diff(x, y, j << (s - 1), A, halfN_sh);
      sts_fr(p0, p1, e0, sum);
      sts_fr(p0, p1, e1, d);
    }
    __syncthreads();
  }

  // ---- store: out[U * 2^(t+P) + rev_t(m) * 2^P + Pb] -------------------------------------------
  const uint32_t revsh = 32 - t;
  for (uint32_t o = threadIdx.x; o < tile; o += kThreads) {
    uint32_t r, c;
    if (P == 0) { c = o >> t; r = o & (rows - 1); }      // each column's 2^t outputs are contiguous
    else { r = o >> logC; c = o & (C - 1); }             // each output row holds C contiguous columns
    uint32_t m = t ? (__brev(r) >> revsh) : 0;
    size_t rest = rest0 + c;
    size_t U = rest >> P, Pb = rest & (((size_t)1 << P) - 1);
    size_t dst = (U << (t + P)) + ((size_t)r << P) + Pb;
    Fr v = lds_fr(p0, p1, (m << logC) + c);
    if (A.post_mode == 1) v = v * K.post[0];
    else if (A.post_mode == 2) v = v * K.post[dst % 3];
    v.store(out + 2 * dst);
  }
}

// ---- warp-resident pass -----------------------------------------------------------------------------------------------
// The same pass (t <= 7 stages of the DIF network, same twiddles, same digit-reversed write-back) with NO shared memory
// and NO barrier: a warp owns 128 points = 2^t rows x 2^(7-t) adjacent columns, four per lane.  Point q (7 bits: row in the
// top t bits, column below) starts in register q >> 5 of lane q & 31.  Stages 1 and 2 pair registers of one lane; every
// later stage first swaps the lane bit it needs with the upper register bit (each lane keeps the half it owns and receives
// the other half from its partner, 16 SHFL), so that again both operands of a butterfly sit in one lane and every lane
// performs two full butterflies per stage (no lane computes a product that is thrown away).  Warps never wait for each
// other, so the multiplier pipe is fed by whichever warps are in their arithmetic phase (the shared-memory version
// synchronises the 8 warps of a CTA seven times per pass: ncu showed the pipe idle 26 % of the time,
// profiles/r1_fullproof_ntt_pass_v5.txt).  The last stage of the last pass has only trivial twiddles and is done with
// additions alone.
__device__ __forceinline__ uint32_t ntt_w_logical(uint32_t lane, uint32_t a, uint32_t u_done) {
  // logical point index of register `a` of `lane` after the exchanges of stages 3..u_done
  uint32_t q = (a & 1u) << 5;
  const uint32_t r1 = (a >> 1) & 1u;
  if (u_done < 3) return q | (r1 << 6) | lane;
  const uint32_t low = 7 - u_done;                       // lane bits below `low` still hold their own index bit
  q |= r1 << low;
  q |= ((lane >> 4) & 1u) << 6;
  q |= ((lane & 0xFu) >> low) << (low + 1);              // lane bit b in [low, 3] holds index bit b + 1
  q |= lane & ((1u << low) - 1u);
  return q;
}
__device__ __forceinline__ Fr ntt_w_twiddle(const PassArgs& A, uint32_t qx, uint32_t u, uint32_t cw_log, size_t rest0) {
  const uint32_t hb = A.t - u, halfN_sh = A.L - 1;
  const size_t m_lo = (qx >> cw_log) & ((1u << hb) - 1u);
  const size_t c = qx & ((1u << cw_log) - 1u);
  const size_t j = (m_lo << (A.L - A.s0 - A.t)) + ((rest0 + c) >> A.s0);
  const size_t ex = j << (A.s0 + u - 1);
  // omega^-ex = -omega^(N/2 - ex); ex = 0 maps to table[0] = 1 (the caller negates through (y - x) only when ex != 0)
  const size_t idx = (A.inverse && ex) ? (((size_t)1 << halfN_sh) - ex) : ex;
  return Fr::load(A.tw + 2 * (idx << A.tw_shift));
}
__device__ __forceinline__ void ntt_w_butterfly(Fr& x, Fr& y, const Fr& w, bool negate) {
  Fr s = x + y, d = negate ? (y - x) : (x - y);
  x = s;
  y = d * w;
}
constexpr int kWarpPassMaxT = 7;
// coset / 1/N scaling on the first load and the last store: called, not inlined (8 more copies of the multiplier would
// triple the kernel's code size for a path only the first and last pass take)
__device__ __noinline__ Fr ntt_w_scale(Fr a, Fr b) { return a * b; }
constexpr int kWarpPassThreads = 128;     // 4 independent warps per CTA
// MINB = resident CTAs per SM the register allocation is held to (5: <= 102 registers, 4: <= 128, 3: <= 168): more
// registers let ptxas interleave the two independent multiplications of a lane (instruction-level parallelism on the
// IMAD carry chains) at the price of fewer resident warps; ZK_NTT_W_CTAS selects, measured default below
template <int MINB>
__global__ void __launch_bounds__(kWarpPassThreads, MINB) k_ntt_pass_w(PassArgs A, PassConsts K) {
  const uint32_t lane = threadIdx.x & 31u, T = A.t, L = A.L, P = A.s0;
  const uint32_t cw_log = 7 - T, cw_mask = (1u << cw_log) - 1u;
  const size_t warp = (size_t)blockIdx.x * (kWarpPassThreads / 32) + (threadIdx.x >> 5);
  const size_t rest0 = warp << cw_log;
  const uint32_t row_sh = L - T;
  const uint4* in = A.in + 2 * (size_t)blockIdx.y * A.in_stride;
  uint4* out = A.out + 2 * (size_t)blockIdx.y * A.out_stride;
  Fr v[4];
#pragma unroll
  for (uint32_t a = 0; a < 4; a++) {
    const uint32_t q = (a << 5) | lane;
    const size_t elem = ((size_t)(q >> cw_log) << row_sh) + rest0 + (q & cw_mask);
    v[a] = Fr::zero();
    if (elem < A.n_in) {
      v[a] = Fr::load(in + 2 * elem);
      if (A.pre_mode) {
        const uint32_t r3 = (uint32_t)(elem % 3);
        if (r3 == 1) v[a] = ntt_w_scale(v[a], K.pre[1]);
        else if (r3 == 2) v[a] = ntt_w_scale(v[a], K.pre[2]);
      }
    }
  }
  const bool inv = A.inverse != 0;
  // the negation of the inverse direction applies only where the exponent is non-zero: (y - x) * omega^(N/2 - ex)
  auto neg_for = [&](uint32_t qx, uint32_t u) -> bool {
    if (!inv) return false;
    const uint32_t hb = T - u;
    const size_t m_lo = (qx >> cw_log) & ((1u << hb) - 1u);
    return (m_lo << (L - P - T)) + ((rest0 + (qx & cw_mask)) >> P) != 0;
  };
  // stage 1: register bit 1
  if (T >= 1) {
    if (P + 1 == L) {
      Fr s0 = v[0] + v[2], d0 = v[0] - v[2], s1 = v[1] + v[3], d1 = v[1] - v[3];
      v[0] = s0; v[2] = d0; v[1] = s1; v[3] = d1;
    } else {
      const uint32_t q0 = ntt_w_logical(lane, 0, 1), q1 = ntt_w_logical(lane, 1, 1);
      Fr w0 = ntt_w_twiddle(A, q0, 1, cw_log, rest0), w1 = ntt_w_twiddle(A, q1, 1, cw_log, rest0);
      ntt_w_butterfly(v[0], v[2], w0, neg_for(q0, 1));
      ntt_w_butterfly(v[1], v[3], w1, neg_for(q1, 1));
    }
  }
  // stage 2: register bit 0
  if (T >= 2) {
    if (P + 2 == L) {
      Fr s0 = v[0] + v[1], d0 = v[0] - v[1], s1 = v[2] + v[3], d1 = v[2] - v[3];
      v[0] = s0; v[1] = d0; v[2] = s1; v[3] = d1;
    } else {
      const uint32_t q0 = ntt_w_logical(lane, 0, 2), q1 = ntt_w_logical(lane, 2, 2);
      Fr w0 = ntt_w_twiddle(A, q0, 2, cw_log, rest0), w1 = ntt_w_twiddle(A, q1, 2, cw_log, rest0);
      ntt_w_butterfly(v[0], v[1], w0, neg_for(q0, 2));
      ntt_w_butterfly(v[2], v[3], w1, neg_for(q1, 2));
    }
  }
  // stages 3..T: swap lane bit 7-u with register bit 1, then pair (v[k], v[k+2])
#pragma unroll 1
  for (uint32_t u = 3; u <= T; u++) {
    const uint32_t d = 7 - u, bit = (lane >> d) & 1u;
    const bool trivial = (P + u == L);
    const uint32_t q0 = ntt_w_logical(lane, 0, u), q1 = ntt_w_logical(lane, 1, u);
    Fr w0, w1;
    if (!trivial) { w0 = ntt_w_twiddle(A, q0, u, cw_log, rest0); w1 = ntt_w_twiddle(A, q1, u, cw_log, rest0); }
#pragma unroll
    for (uint32_t k = 0; k < 2; k++) {
#pragma unroll
      for (uint32_t i = 0; i < 8; i++) {
        const uint32_t send = bit ? v[k].l[i] : v[k + 2].l[i];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1u << d);
        if (bit) v[k].l[i] = recv; else v[k + 2].l[i] = recv;
      }
    }
    if (trivial) {
      Fr s0 = v[0] + v[2], d0 = v[0] - v[2], s1 = v[1] + v[3], d1 = v[1] - v[3];
      v[0] = s0; v[2] = d0; v[1] = s1; v[3] = d1;
    } else {
      ntt_w_butterfly(v[0], v[2], w0, neg_for(q0, u));
      ntt_w_butterfly(v[1], v[3], w1, neg_for(q1, u));
    }
  }
  // store with the pass's digit bit-reversed
  const uint32_t u_done = T >= 3 ? T : 2;
#pragma unroll
  for (uint32_t a = 0; a < 4; a++) {
    const uint32_t q = ntt_w_logical(lane, a, u_done);
    const uint32_t m = q >> cw_log;
    const size_t rest = rest0 + (q & cw_mask);
    const uint32_t r = T ? (__brev(m) >> (32 - T)) : 0;
    const size_t U = rest >> P, Pb = rest & (((size_t)1 << P) - 1);
    const size_t dst = (U << (T + P)) + ((size_t)r << P) + Pb;
    Fr o = v[a];
    if (A.post_mode == 1) o = ntt_w_scale(o, K.post[0]);
    else if (A.post_mode == 2) o = ntt_w_scale(o, K.post[dst % 3]);
    o.store(out + 2 * dst);
  }
}

// tw[e] = omega^e for e < half_n; each thread produces kTwChunk consecutive powers
constexpr int kTwChunk = 32;
__global__ void k_gen_twiddles(uint4* tw, Fr omega, size_t half_n) {
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t start = g * kTwChunk;
  if (start >= half_n) return;
  Fr w = Fr::pow_u64(omega, (uint64_t)start);
  for (int i = 0; i < kTwChunk && start + i < half_n; i++) {
    w.store(tw + 2 * (start + i));
    w = w * omega;
  }
}

// a[i] *= t[i & mask]   (divide_by_vanishing_poly)
__global__ void k_mul_periodic(uint4* a, const uint4* t, size_t n, size_t mask) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  (Fr::load(a + 2 * i) * Fr::load(t + 2 * (i & mask))).store(a + 2 * i);
}

Fr fr_from_host(const zk_fr* v) { Fr r; memcpy(r.l, v, 32); return r; }
void fr_to_host(const Fr& v, zk_fr* out) { Fr c = v.reduce(); memcpy(out, c.l, 32); }
bool fr_same(const zk_fr& a, const zk_fr& b) { return memcmp(&a, &b, 32) == 0; }

}  // namespace

// Device table of omega^e, e < 2^(log_n-1), cached per (omega, log_n) on the context.
int zk_twiddles_get(zk_ctx* ctx, const zk_fr* omega, uint32_t log_n, const void** table) {
  for (auto& t : ctx->twiddles)
    if (t.log_n == log_n && fr_same(t.omega, *omega)) { *table = t.table; return ZK_OK; }
  size_t half = log_n ? ((size_t)1 << (log_n - 1)) : 1;
  void* d = nullptr;
  ZK_CUDA(ctx, cudaMalloc(&d, half * 32));
  size_t threads = (half + kTwChunk - 1) / kTwChunk;
  ZK_LAUNCH(ctx, "ntt_twiddles", k_gen_twiddles, zk_div_up(threads, 128), 128, 0, (uint4*)d, fr_from_host(omega), half);
  zk_ctx::tw_entry e; e.log_n = log_n; e.omega = *omega; e.table = d;
  ctx->twiddles.push_back(e);
  *table = d;
  return ZK_OK;
}

namespace {

struct NttJob {
  const void* in;       // device
  void* out;            // device, may alias `in` only when the transform is a single pass... never aliased here
  void* tmp;            // device scratch of count * N elements (ping-pong partner of out)
  size_t count, in_stride, n_in;
  size_t out_stride;    // elements between consecutive polynomials in `out` (0 = 2^L: densely packed)
  uint32_t L;
  const void* tw; uint32_t tw_shift;
  int inverse, pre_mode, post_mode;
  PassConsts K;
};

// Runs all passes.  The data ping-pongs between job.tmp and job.out so that the LAST pass
// writes job.out; the first pass reads job.in (which is never written).
int run_ntt(zk_ctx* ctx, const NttJob& J) {
  const uint32_t L = J.L;
  const size_t N = (size_t)1 << L;
  size_t smem_max = (((size_t)1 << kTileLog) * 32) + 64;
  // radix-4 stages (ZK_NTT_RADIX4=1): measured on B200 at k = 19 -- 35.7 ms of NTT per proof with 3 CTAs / SM (39.2 ms with 2)
  // against 35.4 ms for the radix-2 loop: shared-memory traffic and barriers are not what bounds the pass, so it stays off.
  static const bool radix4 = getenv("ZK_NTT_RADIX4") && atoi(getenv("ZK_NTT_RADIX4")) == 1;
  ZK_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass_t<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  ZK_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  // ZK_NTT_WARP=1: warp-resident passes (register butterflies + shuffles, no shared memory, no barriers) from 2^10 points on.
  // Measured on B200 inside the complete k = 19 proof: 35.9 ms of NTT per proof against 35.4 ms for the shared-memory passes
  // (ncu, one 2^21 pass: fmaheavy 77.5 % vs 73.8 % active, the same 168 us -- the extra activity is the index arithmetic of
  // the exchanges), so the shared-memory passes stay the default (profiles/r2_ncu_ntt_pass_{w,t}.txt).
  static const bool want_warp = getenv("ZK_NTT_WARP") && atoi(getenv("ZK_NTT_WARP")) == 1;
  const bool warp_pass = want_warp && L >= 10;
  const uint32_t max_t = warp_pass ? (uint32_t)kWarpPassMaxT : (uint32_t)kMaxT;
  uint32_t npass = L == 0 ? 1 : (L + max_t - 1) / max_t;
  uint32_t base_t = L / npass, extra = L % npass;   // first `extra` passes take one more stage
  uint32_t min_t = base_t;
  uint32_t s0 = 0;
  const void* src = J.in;
  size_t src_stride = J.in_stride;
  if (npass > 1 && !J.tmp) ZK_FAIL(ctx, ZK_ERR_STATE, "ntt: %u passes need a ping-pong buffer", npass);
  for (uint32_t p = 0; p < npass; p++) {
    uint32_t t = base_t + (p < extra ? 1 : 0);
    uint32_t logC = kTileLog - t;
    if (logC > L - t) logC = L - t;
    if (npass > 1 && logC > min_t) logC = min_t;      // keep written runs contiguous (P >= logC)
    bool last = (p + 1 == npass);
    // ping-pong so that the last pass lands in J.out
    void* dst = ((npass - 1 - p) % 2 == 0) ? J.out : J.tmp;
    PassArgs A;
    A.in = (const uint4*)src; A.out = (uint4*)dst; A.tw = (const uint4*)J.tw;
    A.in_stride = src_stride; A.out_stride = (dst == J.out && J.out_stride) ? J.out_stride : N;
    A.n_in = (p == 0) ? J.n_in : N;
    A.L = L; A.s0 = s0; A.t = t; A.logC = logC; A.tw_shift = J.tw_shift;
    A.inverse = J.inverse;
    A.pre_mode = (p == 0) ? J.pre_mode : 0;
    A.post_mode = last ? J.post_mode : 0;
    if (warp_pass) {
      dim3 grid((unsigned)(N >> 9), (unsigned)J.count);           // 4 warps x 128 points per CTA
      static const int minb = getenv("ZK_NTT_W_CTAS") ? atoi(getenv("ZK_NTT_W_CTAS")) : 4;
      if (minb >= 5) { ZK_LAUNCH(ctx, "ntt_pass", k_ntt_pass_w<5>, grid, kWarpPassThreads, 0, A, J.K); }
      else if (minb == 4) { ZK_LAUNCH(ctx, "ntt_pass", k_ntt_pass_w<4>, grid, kWarpPassThreads, 0, A, J.K); }
      else { ZK_LAUNCH(ctx, "ntt_pass", k_ntt_pass_w<3>, grid, kWarpPassThreads, 0, A, J.K); }
    } else {
      size_t tile = (size_t)1 << (t + logC);
      dim3 grid((unsigned)(N >> (t + logC)), (unsigned)J.count);
      size_t smem = tile * 32 + 64;
      if (radix4) { ZK_LAUNCH(ctx, "ntt_pass", k_ntt_pass_t<1>, grid, kThreads, smem, A, J.K); }
      else { ZK_LAUNCH(ctx, "ntt_pass", k_ntt_pass_t<0>, grid, kThreads, smem, A, J.K); }
    }
    src = dst; src_stride = A.out_stride;
    s0 += t;
  }
  return ZK_OK;
}

}  // namespace

// ---- best_fft -------------------------------------------------------------------------------

extern "C" int zk_ntt_dev(zk_ctx* ctx, zk_fr* d_a, const zk_fr* omega, uint32_t log_n) {
  if (!ctx || !d_a || !omega || log_n > 28) ZK_FAIL(ctx ? ctx : nullptr, ZK_ERR_ARG, "zk_ntt: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (log_n == 0) return ZK_OK;
  size_t N = (size_t)1 << log_n;
  const void* tw = nullptr;
  ZK_TRY(zk_twiddles_get(ctx, omega, log_n, &tw));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->ntt_buf, N * 32));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->ntt_buf2, N * 32));
  // out-of-place into ntt_buf (ping-pong with ntt_buf2), then copy back: best_fft is in place
  NttJob J; memset(&J, 0, sizeof J);
  J.in = d_a; J.out = ctx->ntt_buf.p; J.tmp = ctx->ntt_buf2.p; J.count = 1; J.in_stride = N; J.n_in = N;
  J.L = log_n; J.tw = tw; J.tw_shift = 0;
  ZK_TRY(run_ntt(ctx, J));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_a, ctx->ntt_buf.p, N * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  return ZK_OK;
}

extern "C" int zk_ntt(zk_ctx* ctx, zk_fr* a, const zk_fr* omega, uint32_t log_n) {
  if (!ctx) return ZK_ERR_ARG;
  if (!a || !omega || log_n > 28) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_ntt: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t N = (size_t)1 << log_n;
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, N * 32));
  ZK_CUDA(ctx, cudaMemcpyAsync(ctx->misc.p, a, N * 32, cudaMemcpyHostToDevice, ctx->stream));
  ZK_TRY(zk_ntt_dev(ctx, (zk_fr*)ctx->misc.p, omega, log_n));
  ZK_CUDA(ctx, cudaMemcpyAsync(a, ctx->misc.p, N * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}

// ---- EvaluationDomain -------------------------------------------------------------------------

struct zk_domain {
  zk_ctx* ctx;
  uint32_t j, k, extended_k;
  size_t n, ext_n;
  Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor;
  const void* tw_ext = nullptr;      // omega_ext^e, e < ext_n/2 (serves all four directions)
  void* t_evaluations = nullptr;     // 2^(extended_k-k) values
  zk_scratch work_a, work_b;
};

namespace {
// ROOT_OF_UNITY = 7^((r-1)/2^28) (halo2curves bn256/fr.rs), canonical value, little-endian u32 limbs
Fr root_of_unity() {
  Fr v;
  const uint32_t c[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
  memcpy(v.l, c, 32);
  return v.to_mont();
}
Fr zeta() {
  Fr v;
  const uint32_t c[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
  memcpy(v.l, c, 32);
  return v.to_mont();
}
Fr fr_u64(uint64_t x) { Fr v = Fr::zero(); v.l[0] = (uint32_t)x; v.l[1] = (uint32_t)(x >> 32); return v.to_mont(); }
}  // namespace

extern "C" int zk_domain_create(zk_ctx* ctx, uint32_t j, uint32_t k, zk_domain** out) {
  if (!ctx || !out) return ZK_ERR_ARG;
  if (j < 2 || k > 27) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_domain_create: need j >= 2 and k <= 27 (got j=%u k=%u)", j, k);
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  zk_domain* d = new zk_domain();
  d->ctx = ctx; d->j = j; d->k = k; d->n = (size_t)1 << k;
  uint32_t ek = k;
  while (((size_t)1 << ek) < d->n * (j - 1)) ek++;
  if (ek > 28) { delete d; ZK_FAIL(ctx, ZK_ERR_ARG, "zk_domain_create: extended_k %u exceeds the 2-adicity 28", ek); }
  d->extended_k = ek; d->ext_n = (size_t)1 << ek;
  Fr eo = root_of_unity();
  for (uint32_t i = ek; i < 28; i++) eo = eo.sqr();
  d->extended_omega = eo;
  d->extended_omega_inv = Fr::inv(eo);
  Fr o = eo, oi = d->extended_omega_inv;
  for (uint32_t i = k; i < ek; i++) { o = o.sqr(); oi = oi.sqr(); }
  d->omega = o; d->omega_inv = oi;
  d->g_coset = zeta(); d->g_coset_inv = d->g_coset.sqr();
  d->ifft_divisor = Fr::inv(fr_u64((uint64_t)1 << k));
  d->extended_ifft_divisor = Fr::inv(fr_u64((uint64_t)1 << ek));
  // t_evaluations[i] = 1 / ((zeta * omega_ext^i)^n - 1), i < 2^(ek-k)
  size_t nt = (size_t)1 << (ek - k);
  std::vector<zk_fr> te(nt);
  Fr cur = Fr::pow_u64(d->g_coset, (uint64_t)d->n), step = Fr::pow_u64(eo, (uint64_t)d->n);
  for (size_t i = 0; i < nt; i++) {
    fr_to_host(Fr::inv(cur - Fr::one()), &te[i]);
    cur = cur * step;
  }
  if (cudaMalloc(&d->t_evaluations, nt * 32) != cudaSuccess) { delete d; ZK_FAIL(ctx, ZK_ERR_OOM, "t_evaluations alloc"); }
  ZK_CUDA(ctx, cudaMemcpyAsync(d->t_evaluations, te.data(), nt * 32, cudaMemcpyHostToDevice, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  zk_fr eo_h; fr_to_host(eo, &eo_h);
  int r = zk_twiddles_get(ctx, &eo_h, ek, &d->tw_ext);
  if (r != ZK_OK) { cudaFree(d->t_evaluations); delete d; return r; }
  *out = d;
  return ZK_OK;
}
extern "C" void zk_domain_destroy(zk_domain* d) {
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  if (d->t_evaluations) cudaFree(d->t_evaluations);
  if (d->work_a.p) cudaFree(d->work_a.p);
  if (d->work_b.p) cudaFree(d->work_b.p);
  delete d;
}
// internal accessors used by evalh.cu
const void* zk_domain_twiddles(const zk_domain* d) { return d->tw_ext; }
size_t zk_domain_ext_n(const zk_domain* d) { return d->ext_n; }
extern "C" uint32_t zk_domain_k(const zk_domain* d) { return d ? d->k : 0; }
extern "C" uint32_t zk_domain_extended_k(const zk_domain* d) { return d ? d->extended_k : 0; }
extern "C" int zk_domain_constants(const zk_domain* d, zk_fr out[8]) {
  if (!d || !out) return ZK_ERR_ARG;
  const Fr* src[8] = {&d->omega, &d->omega_inv, &d->extended_omega, &d->extended_omega_inv,
                      &d->g_coset, &d->g_coset_inv, &d->ifft_divisor, &d->extended_ifft_divisor};
  for (int i = 0; i < 8; i++) fr_to_host(*src[i], &out[i]);
  return ZK_OK;
}

namespace {
// transform `count` polynomials: src (stride in_stride, n_in valid) -> dst (stride 2^L)
int domain_ntt(zk_domain* d, const void* src, void* dst, size_t count, size_t in_stride, size_t n_in,
               uint32_t L, int inverse, int pre_mode, int post_mode, const PassConsts& K, size_t out_stride = 0) {
  zk_ctx* ctx = d->ctx;
  size_t N = (size_t)1 << L;
  uint32_t npass = L == 0 ? 1 : (L + kMaxT - 1) / kMaxT;
  if (L > (uint32_t)kWarpPassMaxT) npass = 2;        // warp-resident passes take at most 7 stages each
  NttJob J; memset(&J, 0, sizeof J);
  J.in = src; J.out = dst; J.count = count; J.in_stride = in_stride; J.n_in = n_in; J.L = L;
  J.tw = d->tw_ext; J.tw_shift = d->extended_k - L;
  J.inverse = inverse; J.pre_mode = pre_mode; J.post_mode = post_mode; J.K = K;
  J.out_stride = out_stride;
  if (npass > 1) {
    ZK_TRY(zk_scratch_reserve(ctx, d->work_b, count * N * 32));
    J.tmp = d->work_b.p;
  }
  if (L == 0) {   // a single point: the pass kernel handles t = 0 (pure copy with scaling)
    J.tmp = nullptr;
  }
  return run_ntt(ctx, J);
}
}  // namespace

extern "C" int zk_lagrange_to_coeff_dev(zk_domain* d, zk_fr* d_a, size_t count) {
  if (!d || !d_a) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!count) return ZK_OK;
  ZK_TRY(zk_scratch_reserve(ctx, d->work_a, count * d->n * 32));
  PassConsts K; memset(&K, 0, sizeof K); K.post[0] = d->ifft_divisor;
  ZK_TRY(domain_ntt(d, d_a, d->work_a.p, count, d->n, d->n, d->k, 1, 0, 1, K));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_a, d->work_a.p, count * d->n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  return ZK_OK;
}
// out-of-place lagrange_to_coeff (internal: the prover keeps the Lagrange values for the commitment running on
// another stream)
int zk_lagrange_to_coeff_oop(zk_domain* d, const zk_fr* d_src, zk_fr* d_dst, size_t count) {
  if (!d || !d_src || !d_dst) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!count) return ZK_OK;
  PassConsts K; memset(&K, 0, sizeof K); K.post[0] = d->ifft_divisor;
  return domain_ntt(d, d_src, d_dst, count, d->n, d->n, d->k, 1, 0, 1, K);
}
extern "C" int zk_coeff_to_lagrange_dev(zk_domain* d, zk_fr* d_a, size_t count) {
  if (!d || !d_a) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!count) return ZK_OK;
  ZK_TRY(zk_scratch_reserve(ctx, d->work_a, count * d->n * 32));
  PassConsts K; memset(&K, 0, sizeof K);
  ZK_TRY(domain_ntt(d, d_a, d->work_a.p, count, d->n, d->n, d->k, 0, 0, 0, K));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_a, d->work_a.p, count * d->n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  return ZK_OK;
}
extern "C" int zk_coeff_to_extended_dev(zk_domain* d, const zk_fr* d_a, zk_fr* d_out_ext, size_t count) {
  if (!d || !d_a || !d_out_ext) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!count) return ZK_OK;
  PassConsts K; memset(&K, 0, sizeof K);
  K.pre[0] = Fr::one(); K.pre[1] = d->g_coset; K.pre[2] = d->g_coset_inv;    // zeta^(i mod 3)
  return domain_ntt(d, d_a, d_out_ext, count, d->n, d->n, d->extended_k, 0, 1, 0, K);
}
// coeff_to_extended of `count` polynomials whose cosets go to slots `out_stride` elements apart (a multiple of the
// extended size; e.g. the A' columns of all lookup arguments into every third evaluate_h slot): ONE batch of launches
int zk_coeff_to_extended_strided(zk_domain* d, const zk_fr* d_a, zk_fr* d_out_ext, size_t count, size_t out_stride) {
  if (!d || !d_a || !d_out_ext) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!count) return ZK_OK;
  if (out_stride < d->ext_n) ZK_FAIL(ctx, ZK_ERR_ARG, "coeff_to_extended: output stride below the extended size");
  PassConsts K; memset(&K, 0, sizeof K);
  K.pre[0] = Fr::one(); K.pre[1] = d->g_coset; K.pre[2] = d->g_coset_inv;
  return domain_ntt(d, d_a, d_out_ext, count, d->n, d->n, d->extended_k, 0, 1, 0, K, out_stride);
}
extern "C" int zk_extended_to_coeff_dev(zk_domain* d, zk_fr* d_a_ext, zk_fr* d_out) {
  if (!d || !d_a_ext || !d_out) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  // ifft * 1/ext_n, then coefficient i times zeta^-(i mod 3); keep the first n*(j-1)
  ZK_TRY(zk_scratch_reserve(ctx, d->work_a, d->ext_n * 32));
  PassConsts K; memset(&K, 0, sizeof K);
  K.post[0] = d->extended_ifft_divisor;
  K.post[1] = d->extended_ifft_divisor * d->g_coset_inv;
  K.post[2] = d->extended_ifft_divisor * d->g_coset;
  ZK_TRY(domain_ntt(d, d_a_ext, d->work_a.p, 1, d->ext_n, d->ext_n, d->extended_k, 1, 0, 2, K));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_out, d->work_a.p, d->n * (d->j - 1) * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  return ZK_OK;
}
extern "C" int zk_divide_by_vanishing_poly_dev(zk_domain* d, zk_fr* d_a_ext) {
  if (!d || !d_a_ext) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t mask = ((size_t)1 << (d->extended_k - d->k)) - 1;
  ZK_LAUNCH(ctx, "vanishing_div", k_mul_periodic, zk_div_up(d->ext_n, 256), 256, 0, (uint4*)d_a_ext,
            (const uint4*)d->t_evaluations, d->ext_n, mask);
  return ZK_OK;
}

// ---- host-buffer variants -------------------------------------------------------------------
namespace {
int stage_in(zk_ctx* ctx, zk_scratch& s, const void* host, size_t bytes) {
  ZK_TRY(zk_scratch_reserve(ctx, s, bytes));
  ZK_CUDA(ctx, cudaMemcpyAsync(s.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return ZK_OK;
}
int stage_out(zk_ctx* ctx, void* host, const void* dev, size_t bytes) {
  ZK_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}
}  // namespace

extern "C" int zk_lagrange_to_coeff(zk_domain* d, zk_fr* a) {
  if (!d || !a) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(stage_in(ctx, ctx->misc, a, d->n * 32));
  ZK_TRY(zk_lagrange_to_coeff_dev(d, (zk_fr*)ctx->misc.p, 1));
  return stage_out(ctx, a, ctx->misc.p, d->n * 32);
}
extern "C" int zk_coeff_to_lagrange(zk_domain* d, zk_fr* a) {
  if (!d || !a) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(stage_in(ctx, ctx->misc, a, d->n * 32));
  ZK_TRY(zk_coeff_to_lagrange_dev(d, (zk_fr*)ctx->misc.p, 1));
  return stage_out(ctx, a, ctx->misc.p, d->n * 32);
}
extern "C" int zk_coeff_to_extended(zk_domain* d, const zk_fr* a, zk_fr* out_ext) {
  if (!d || !a || !out_ext) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(stage_in(ctx, ctx->misc, a, d->n * 32));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->ntt_buf, d->ext_n * 32));
  ZK_TRY(zk_coeff_to_extended_dev(d, (const zk_fr*)ctx->misc.p, (zk_fr*)ctx->ntt_buf.p, 1));
  return stage_out(ctx, out_ext, ctx->ntt_buf.p, d->ext_n * 32);
}
extern "C" int zk_extended_to_coeff(zk_domain* d, const zk_fr* a_ext, zk_fr* out) {
  if (!d || !a_ext || !out) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(stage_in(ctx, ctx->misc, a_ext, d->ext_n * 32));
  size_t nout = d->n * (d->j - 1);
  ZK_TRY(zk_scratch_reserve(ctx, ctx->ntt_buf, nout * 32));
  ZK_TRY(zk_extended_to_coeff_dev(d, (zk_fr*)ctx->misc.p, (zk_fr*)ctx->ntt_buf.p));
  return stage_out(ctx, out, ctx->ntt_buf.p, nout * 32);
}
extern "C" int zk_divide_by_vanishing_poly(zk_domain* d, zk_fr* a_ext) {
  if (!d || !a_ext) return ZK_ERR_ARG;
  zk_ctx* ctx = d->ctx;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(stage_in(ctx, ctx->misc, a_ext, d->ext_n * 32));
  ZK_TRY(zk_divide_by_vanishing_poly_dev(d, (zk_fr*)ctx->misc.p));
  return stage_out(ctx, a_ext, ctx->misc.p, d->ext_n * 32);
}
