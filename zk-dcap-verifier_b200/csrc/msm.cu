// msm.cu -- BN254 G1 multi-scalar multiplication (Pippenger) for sm_100a.
//
// Replaces halo2_proofs arithmetic::best_multiexp / multiexp_serial ([UPSTREAM]
// src/arithmetic.rs @ c254c75f, Cargo.lock:1314-1316) behind ParamsKZG::commit and
// commit_lagrange (src/poly/kzg/commitment.rs), i.e. every KZG commitment made by
// create_proof (circuits/src/sgx_dcap_verifier.rs:814), keygen_vk (:803) and the
// snark-verifier-sdk provers (crates/p256-ecdsa/src/base.rs:193-212).
//
// The result of an MSM is a group element, so any correct evaluation order gives the same
// canonical point; the device algorithm is therefore free to differ from the CPU one:
//
//   1. digits    scalar -> canonical integer (one Montgomery reduction) -> W signed digits of
//                c bits, d in (-2^(c-1), 2^(c-1)], so only 2^(c-1) buckets per window exist.
//   2. sort      counting sort of the (digit, point) pairs by bucket: histogram (atomics in
//                L2), exclusive scan, scatter.  An entry is a 32-bit point index + sign bit.
//   3. buckets   every bucket (split over S threads when there are few buckets) adds up its
//                points with mixed XYZZ additions (8M + 2S, no inversion); points are fetched
//                as 64-byte AoS records because the access is a gather.
//   4. reduce    sum_b (b+1) * B_b  =  sum_k 2^k * (sum of the buckets whose weight has bit k),
//                i.e. c independent plain sums done as parallel segmented sums, then a c-step
//                double-and-add -- no long serial running-sum chain.
//   5. windows   with resident bases the table 2^(c*w) * G_i is precomputed once per SRS, so all
//                windows share ONE bucket set and step 5 vanishes; otherwise the W window sums
//                are combined by Horner with c doublings each.
#include <stdlib.h>

#include "common.cuh"
#include "ec.cuh"

using namespace zk;

namespace {

struct MsmPlan {
  uint32_t c, W, G;        // window bits, windows, bucket groups (1 when the window table is used)
  uint32_t B;              // buckets per group = 2^(c-1)
  uint32_t S;              // threads per bucket
  size_t n, count;         // points per MSM, MSMs in the batch
  size_t stride;           // window table: table[w * stride + i]
  size_t nb;               // G * B buckets per MSM
  size_t ent_per_msm;      // n * W
  size_t sc_stride;        // scalars of MSM m start at m * sc_stride + base_off (a base RANGE of a longer column)
  size_t base_off;         // first base of the range: table index = w * stride + base_off + i
};

// ---- 1+2: digits, histogram, scatter --------------------------------------------------------------

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const uint4* scalars, MsmPlan P, uint32_t* counts,
                                                   const uint32_t* starts, uint32_t* entries, size_t i_begin,
                                                   size_t i_end) {
  size_t i = i_begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t msm = blockIdx.y;
  if (i >= i_end) return;
  Fr s = Fr::load(scalars + 2 * (msm * P.sc_stride + P.base_off + i)).from_mont();
  uint32_t* cnt = counts + msm * P.nb;
  int carry = 0;
  for (uint32_t w = 0; w < P.W; w++) {
    int d = next_signed_digit(s.l, (int)P.c, carry);
    if (d == 0) continue;
    uint32_t neg = d < 0 ? 1u : 0u;
    uint32_t mag = (uint32_t)(d < 0 ? -d : d);
    size_t bucket = (P.G == 1 ? 0 : (size_t)w * P.B) + (mag - 1);
    if (!SCATTER) {
      atomicAdd(&cnt[bucket], 1u);
    } else {
      uint32_t pos = atomicAdd(&cnt[bucket], 1u);   // cnt is the zeroed cursor array here
      uint32_t idx = (P.G == 1) ? (uint32_t)(w * P.stride + P.base_off + i) : (uint32_t)i;
      entries[msm * P.ent_per_msm + starts[msm * (P.nb + 1) + bucket] + pos] = idx | (neg << 31);
    }
  }
}

// The same two passes with the bucket counters of the MSM in SHARED memory: a CTA takes a contiguous chunk of one MSM's
// scalars, counts its digits on chip, and global memory sees ONE atomic per (CTA, non-empty bucket) instead of one per
// digit.  That one atomic also reserves the CTA's share of the bucket; the scatter pass (same grid, same chunks) starts its
// shared counters at those positions and hands them out with shared-memory atomics.  The order of
// the entries inside a bucket is arbitrary either way (the bucket sum does not depend on it).  Needs nb * 4 bytes of shared
// memory (128 KB for the resident window table at c = 16); ncu on the global-atomic version: 73 of 78 stall cycles per issue
// on the atomics' round trip at full occupancy (profiles/r1_fullproof_msm_digits_v7.txt).
constexpr uint32_t kDigitsSmemThreads = 1024;
template <class F>
__device__ __forceinline__ void for_each_digit(const uint4* scalars, const MsmPlan& P, size_t msm, size_t i, F&& f) {
  Fr s = Fr::load(scalars + 2 * (msm * P.sc_stride + P.base_off + i)).from_mont();
  int carry = 0;
  for (uint32_t w = 0; w < P.W; w++) {
    int d = next_signed_digit(s.l, (int)P.c, carry);
    if (d == 0) continue;
    uint32_t neg = d < 0 ? 1u : 0u;
    uint32_t mag = (uint32_t)(d < 0 ? -d : d);
    uint32_t bucket = (P.G == 1 ? 0u : w * P.B) + (mag - 1);
    uint32_t idx = (P.G == 1) ? (uint32_t)(w * P.stride + P.base_off + i) : (uint32_t)i;
    f(bucket, idx | (neg << 31));
  }
}
template <bool SCATTER>
__global__ void __launch_bounds__(kDigitsSmemThreads, 1) k_msm_digits_smem(const uint4* scalars, MsmPlan P, uint32_t* counts,
                                                                          const uint32_t* starts, uint32_t* entries, size_t i_begin,
                                                                          size_t i_end, size_t chunk, uint32_t* cta_base) {
  extern __shared__ uint32_t sh_cnt[];
  const size_t msm = blockIdx.y;
  const size_t lo = i_begin + blockIdx.x * chunk;
  const size_t hi = lo + chunk < i_end ? lo + chunk : i_end;
  const uint32_t nb = (uint32_t)P.nb;
  // this CTA's share of every bucket starts at base[b] inside the bucket: reserved by the histogram pass (the value its one
  // global atomic returned), read back by the scatter pass, which therefore walks its chunk only once
  uint32_t* base = cta_base + ((i_begin / chunk + blockIdx.x) * (size_t)gridDim.y + blockIdx.y) * P.nb;   // launches start on chunk boundaries
  if (!SCATTER) {
    uint32_t* cnt = counts + msm * P.nb;
    for (uint32_t b = threadIdx.x; b < nb; b += kDigitsSmemThreads) sh_cnt[b] = 0;
    __syncthreads();
    for (size_t i = lo + threadIdx.x; i < hi; i += kDigitsSmemThreads)
      for_each_digit(scalars, P, msm, i, [&](uint32_t bucket, uint32_t) { atomicAdd(&sh_cnt[bucket], 1u); });
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += kDigitsSmemThreads) {
      uint32_t v = sh_cnt[b];
      base[b] = v ? atomicAdd(&cnt[b], v) : 0u;
    }
    return;
  }
  const uint32_t* st = starts + msm * (P.nb + 1);
  for (uint32_t b = threadIdx.x; b < nb; b += kDigitsSmemThreads) sh_cnt[b] = st[b] + base[b];
  __syncthreads();
  uint32_t* ent = entries + msm * P.ent_per_msm;
  for (size_t i = lo + threadIdx.x; i < hi; i += kDigitsSmemThreads)
    for_each_digit(scalars, P, msm, i, [&](uint32_t bucket, uint32_t e) { ent[atomicAdd(&sh_cnt[bucket], 1u)] = e; });
}

// exclusive scan of nb counts per MSM (one CTA per MSM), writes nb+1 offsets and clears the counts
__global__ void __launch_bounds__(1024) k_msm_scan(uint32_t* counts, uint32_t* starts, size_t nb, unsigned long long* counters) {
  __shared__ uint32_t part[1024];
  size_t msm = blockIdx.x;
  uint32_t* cnt = counts + msm * nb;
  uint32_t* st = starts + msm * (nb + 1);
  size_t per = (nb + 1023) / 1024;
  size_t lo = threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
  uint32_t sum = 0;
  for (size_t b = lo; b < hi; b++) sum += cnt[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (size_t b = lo; b < hi; b++) { uint32_t c = cnt[b]; st[b] = run; run += c; cnt[b] = 0; }
  if (threadIdx.x == 1023) {
    st[nb] = part[1023];
    if (counters) { atomicAdd(counters, (unsigned long long)part[1023]); atomicAdd(counters + 1, 1ull); }
  }
}

// ---- 3: bucket accumulation -------------------------------------------------------------------------

// Load-balanced segmented accumulation.  The sorted entry list is cut into chunks of T entries, one
// chunk per thread, REGARDLESS of bucket boundaries, so a bucket holding half of all entries (bit /
// byte valued witness columns, a narrow top window) costs the same as a uniform distribution.
// A run of equal-bucket entries that lies completely inside a chunk is that bucket's final sum and is
// written to `buckets`; the (at most two) runs that continue into a neighbouring chunk go to the
// thread's two partial slots, tagged with the bucket and with which side is still open.  The slot
// array is again sorted by bucket, so the same reduction is applied to it (k_msm_acc_partials) until
// one thread sees everything.  Empty buckets keep the zeroed (= identity) value.
constexpr uint32_t kSlotEmpty = 0xFFFFFFFFu;
constexpr uint32_t kOpenLeft = 0x80000000u, kOpenRight = 0x40000000u, kSlotIdMask = 0x3FFFFFFFu;

// DIRECT = false: entry q names a table point (index | sign << 31).  DIRECT = true: the q-th point of this
// MSM's own point list (the output of the batched-affine pairing round) is used as it is.
template <bool DIRECT>
__global__ void __launch_bounds__(128, 4) k_msm_accumulate(const uint4* points, size_t pts_stride, const uint32_t* entries,
                                                        const uint32_t* starts, MsmPlan P, uint32_t T,
                                                        uint32_t nthreads, uint4* buckets, uint32_t* slot_ids,
                                                        uint4* slot_pts) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t msm = blockIdx.y;
  if (t >= nthreads) return;
  const uint32_t* st = starts + msm * (P.nb + 1);
  const uint32_t* ent = entries + msm * P.ent_per_msm;
  uint32_t* sid = slot_ids + (msm * nthreads + t) * 2;
  uint4* spt = slot_pts + 8 * ((msm * nthreads + t) * 2);
  uint4* bk = buckets + 8 * (msm * P.nb);
  const uint32_t total = st[P.nb];
  uint32_t id0 = kSlotEmpty, id1 = kSlotEmpty;
  const uint64_t q0 = (uint64_t)t * T;
  if (q0 < total) {
    const uint32_t q1 = (uint32_t)(q0 + T < total ? q0 + T : total);
    // bucket holding entry q0: the last b with st[b] <= q0 (and st[b+1] > q0)
    uint32_t lo = 0, hi = (uint32_t)P.nb;      // invariant: st[lo] <= q0 < st[hi]
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (st[mid] <= q0) lo = mid; else hi = mid;
    }
    // ONE flat loop of exactly q1 - q0 additions per thread: every lane of a warp stays in lockstep on the
    // expensive madd; a bucket boundary only triggers a short (divergent) flush.
    uint32_t b = lo, bend = st[b + 1];
    bool open_l = st[b] < q0;
    XYZZ acc = XYZZ::identity();
    for (uint32_t q = (uint32_t)q0; q < q1; q++) {
      if (q == bend) {                       // bucket b ended inside the chunk
        if (!open_l) acc.store(bk + 8 * (size_t)b);
        else { id0 = b | kOpenLeft; acc.store(spt); }
        acc = XYZZ::identity();
        open_l = false;
        b++;
        while (st[b + 1] == q) b++;          // skip empty buckets
        bend = st[b + 1];
      }
      Affine pt;
      if (DIRECT) {
        pt = Affine::load(points + 4 * (msm * pts_stride + q));
      } else {
        uint32_t e = ent[q];
        pt = Affine::load(points + 4 * (size_t)(e & 0x7fffffffu)).cneg((e >> 31) != 0);
      }
      acc = acc.madd(pt);
    }
    bool open_r = bend > q1;
    if (!open_l && !open_r) {
      acc.store(bk + 8 * (size_t)b);
    } else if (open_l) {
      id0 = b | kOpenLeft | (open_r ? kOpenRight : 0u);
      acc.store(spt);
    } else {
      id1 = b | kOpenRight;
      acc.store(spt + 8);
    }
  }
  sid[0] = id0; sid[1] = id1;
}

// ---- 3a: one batched-affine pairing round -----------------------------------------------------------------
// Inside every bucket the entries are added in pairs (2j, 2j+1) as AFFINE points: lambda = dy / dx needs one
// inversion per addition, and all additions of a round are independent, so the inversions are shared with
// Montgomery's trick: every thread multiplies up the denominators of its M pairs, the 256 thread products of
// a CTA are combined by a prefix / suffix product scan in shared memory, ONE field inversion (thread 0) serves
// the whole CTA (256 * M additions), and the per-pair inverses are unwound backwards.  Cost per addition:
// 1 (running product) + 2 (unwind) + 3 (lambda, lambda^2, y3) = 6 multiplications instead of 10 for the
// inversion-free XYZZ mixed addition.  The round halves the entry list; its output (affine points, bucket
// offsets ceil(m/2)) feeds the chunked XYZZ accumulation.  (0,0) encodes the identity, as in halo2curves.
__global__ void k_pair_counts(const uint32_t* starts, size_t nb, uint32_t* out_counts) {
  size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t msm = blockIdx.y;
  if (b >= nb) return;
  const uint32_t* st = starts + msm * (nb + 1);
  out_counts[msm * nb + b] = (st[b + 1] - st[b] + 1) >> 1;
}

struct PairCase { Affine p, q; Fq d; int kind; };   // kind: 0 add, 1 double, 2 result = p, 3 result = q, 4 identity
// Table points and round outputs are stored canonically, so equality / identity are raw limb compares.
__device__ __forceinline__ bool raw_eq(const Fq& a, const Fq& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i];
  return o == 0;
}
__device__ __forceinline__ Fq load_y_signed(const uint4* points, uint32_t e) {
  Fq y = Fq::load(points + 4 * (size_t)(e & 0x7fffffffu) + 2);
  if (e >> 31) y = y.neg().reduce();
  return y;
}
// Phase 1 needs only the denominator: read the two x coordinates (32 B each); y is touched only when the x's
// coincide or one of them is zero (doubling / cancellation / identity), which is rare.
__device__ __forceinline__ Fq pair_denominator(const uint4* points, const uint32_t* ent, uint32_t q0, bool has_q) {
  if (!has_q) return Fq::one();
  uint32_t e0 = ent[q0], e1 = ent[q0 + 1];
  Fq xp = Fq::load(points + 4 * (size_t)(e0 & 0x7fffffffu)), xq = Fq::load(points + 4 * (size_t)(e1 & 0x7fffffffu));
  bool same = raw_eq(xp, xq);
  if (!same && !xp.is_zero_raw() && !xq.is_zero_raw()) return xq - xp;
  Fq yp = load_y_signed(points, e0), yq = load_y_signed(points, e1);
  if ((xp.is_zero_raw() && yp.is_zero_raw()) || (xq.is_zero_raw() && yq.is_zero_raw())) return Fq::one();   // identity operand
  if (!same) return xq - xp;
  if (raw_eq(yp, yq)) return yp.dbl();          // doubling
  return Fq::one();                             // P = -Q
}
__device__ __forceinline__ PairCase pair_load(const uint4* points, const uint32_t* ent, uint32_t q0, bool has_q) {
  PairCase c;
  uint32_t e0 = ent[q0];
  c.p = Affine::load(points + 4 * (size_t)(e0 & 0x7fffffffu));
  if (e0 >> 31) c.p.y = c.p.y.neg().reduce();
  c.d = Fq::one();
  if (!has_q) { c.kind = 2; c.q = c.p; return c; }
  uint32_t e1 = ent[q0 + 1];
  c.q = Affine::load(points + 4 * (size_t)(e1 & 0x7fffffffu));
  if (e1 >> 31) c.q.y = c.q.y.neg().reduce();
  if (c.q.x.is_zero_raw() && c.q.y.is_zero_raw()) { c.kind = 2; return c; }
  if (c.p.x.is_zero_raw() && c.p.y.is_zero_raw()) { c.kind = 3; return c; }
  if (raw_eq(c.p.x, c.q.x)) {
    if (raw_eq(c.p.y, c.q.y)) { c.kind = 1; c.d = c.p.y.dbl(); } else { c.kind = 4; }
    return c;
  }
  c.kind = 0; c.d = c.q.x - c.p.x;
  return c;
}

constexpr uint32_t kPairThreads = 256;
__global__ void __launch_bounds__(kPairThreads, 3) k_pair_add(const uint4* points, const uint32_t* entries,
                                                           const uint32_t* starts, const uint32_t* out_starts,
                                                           MsmPlan P, uint32_t M, size_t out_cap, uint4* out_pts,
                                                           uint4* scratch) {
  __shared__ uint4 sm_pre[2 * kPairThreads];    // prefix products (two 16-byte halves per element)
  __shared__ uint4 sm_suf[2 * kPairThreads];
  __shared__ uint4 sm_inv[2];
  const size_t msm = blockIdx.y;
  const size_t gthreads = (size_t)gridDim.x * kPairThreads;
  const size_t g = blockIdx.x * (size_t)kPairThreads + threadIdx.x;
  const uint32_t* st = starts + msm * (P.nb + 1);
  const uint32_t* ost = out_starts + msm * (P.nb + 1);
  const uint32_t* ent = entries + msm * P.ent_per_msm;
  uint4* out = out_pts + 4 * (msm * out_cap);
  uint4* scr = scratch + 2 * (msm * (size_t)M * gthreads);
  const uint32_t out_total = ost[P.nb];
  const uint64_t o0 = (uint64_t)g * M;
  const uint32_t o1 = (uint32_t)(o0 + M < out_total ? o0 + M : out_total);
  // bucket of output slot o0
  uint32_t b0 = 0;
  if (o0 < out_total) {
    uint32_t lo = 0, hi = (uint32_t)P.nb;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (ost[mid] <= o0) lo = mid; else hi = mid; }
    b0 = lo;
  }
  // ---- phase 1: running product of the denominators ----------------------------------------------------
  Fq acc = Fq::one();
  {
    uint32_t b = b0;
    for (uint32_t o = (uint32_t)o0, i = 0; o < o1 && o0 < out_total; o++, i++) {
      while (o >= ost[b + 1]) b++;
      uint32_t j = o - ost[b], q0 = st[b] + 2 * j;
      Fq d = pair_denominator(points, ent, q0, q0 + 1 < st[b + 1]);
      acc.store_raw(scr + 2 * ((size_t)i * gthreads + g));
      acc = acc * d;
    }
  }
  // ---- phase 2: CTA-wide product scans and one inversion -------------------------------------------------
  auto put = [](uint4* a, uint32_t i, const Fq& v) {
    a[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]); a[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
  };
  auto get = [](const uint4* a, uint32_t i) {
    uint4 x = a[2 * i], y = a[2 * i + 1]; Fq v;
    v.l[0] = x.x; v.l[1] = x.y; v.l[2] = x.z; v.l[3] = x.w; v.l[4] = y.x; v.l[5] = y.y; v.l[6] = y.z; v.l[7] = y.w;
    return v;
  };
  const uint32_t t = threadIdx.x;
  put(sm_pre, t, acc); put(sm_suf, t, acc);
  __syncthreads();
  for (uint32_t off = 1; off < kPairThreads; off <<= 1) {
    Fq a = get(sm_pre, t), b = get(sm_suf, t);
    bool hp = t >= off, hs = t + off < kPairThreads;
    Fq pa = hp ? get(sm_pre, t - off) : Fq::one();
    Fq sb = hs ? get(sm_suf, t + off) : Fq::one();
    __syncthreads();
    if (hp) put(sm_pre, t, a * pa);
    if (hs) put(sm_suf, t, b * sb);
    __syncthreads();
  }
  if (t == 0) { Fq gi = Fq::inv(get(sm_pre, kPairThreads - 1)); put(sm_inv, 0, gi); }
  __syncthreads();
  Fq inv_run = get(sm_inv, 0);
  if (t > 0) inv_run = inv_run * get(sm_pre, t - 1);
  if (t + 1 < kPairThreads) inv_run = inv_run * get(sm_suf, t + 1);
  // ---- phase 3: unwind backwards, finish the additions ---------------------------------------------------
  if (o0 < out_total) {
    // bucket of the LAST slot, then walk down
    uint32_t b = b0;
    while (o1 - 1 >= ost[b + 1]) b++;
    for (uint32_t o = o1; o-- > (uint32_t)o0;) {
      while (o < ost[b]) b--;
      uint32_t i = o - (uint32_t)o0;
      uint32_t j = o - ost[b], q0 = st[b] + 2 * j;
      PairCase c = pair_load(points, ent, q0, q0 + 1 < st[b + 1]);
      Fq pv = Fq::load(scr + 2 * ((size_t)i * gthreads + g));
      Fq inv_d = inv_run * pv;
      inv_run = inv_run * c.d;
      Affine r;
      if (c.kind <= 1) {
        Fq num, xq;
        if (c.kind == 0) { num = c.q.y - c.p.y; xq = c.q.x; }
        else { Fq xx = c.p.x.sqr(); num = xx.dbl() + xx; xq = c.p.x; }
        Fq lam = num * inv_d;
        r.x = lam.sqr() - c.p.x - xq;
        r.y = lam * (c.p.x - r.x) - c.p.y;
      } else if (c.kind == 2) { r = c.p; }
      else if (c.kind == 3) { r = c.q; }
      else { r = Affine::identity(); }
      r.store(out + 4 * (size_t)o);
    }
  }
}

// Next levels: the same reduction over the slot array (ids sorted, kSlotEmpty interspersed).
// One chunk of T slots -> complete runs into `bk`, at most two open partials into (sid, spt).
__device__ __forceinline__ void partials_chunk(const uint32_t* iid, const uint4* ipt, uint32_t n_in, uint32_t T,
                                               size_t t, uint4* bk, uint32_t* sid, uint4* spt) {
  uint32_t id0 = kSlotEmpty, id1 = kSlotEmpty;
  uint64_t s0 = (uint64_t)t * T;
  uint32_t s1 = (uint32_t)(s0 + T < n_in ? s0 + T : n_in);
  uint32_t s = (uint32_t)s0;
  while (s < s1) {
    uint32_t raw = iid[s];
    if (raw == kSlotEmpty) { s++; continue; }
    uint32_t id = raw & kSlotIdMask;
    bool open_l = (raw & kOpenLeft) != 0, open_r = (raw & kOpenRight) != 0;
    XYZZ acc = XYZZ::load(ipt + 8 * (size_t)s);
    s++;
    // absorb the following slots of the same bucket that lie in this chunk
    while (s < s1) {
      uint32_t r2 = iid[s];
      if (r2 == kSlotEmpty) { s++; continue; }
      if ((r2 & kSlotIdMask) != id) break;
      acc = acc.add(XYZZ::load(ipt + 8 * (size_t)s));
      open_r = (r2 & kOpenRight) != 0;
      s++;
    }
    // a run whose first slot is open to the left (last slot open to the right) continues outside this chunk:
    // a matching neighbour inside the chunk would have been absorbed above
    if (!open_l && !open_r) {
      acc.store(bk + 8 * (size_t)id);
    } else if (open_l) {
      id0 = id | kOpenLeft | (open_r ? kOpenRight : 0u);
      acc.store(spt);
    } else {
      id1 = id | kOpenRight;
      acc.store(spt + 8);
    }
  }
  sid[0] = id0; sid[1] = id1;
}

__global__ void __launch_bounds__(128) k_msm_acc_partials(const uint32_t* in_ids, const uint4* in_pts, uint32_t n_in,
                                                          uint32_t T, uint32_t nthreads, size_t nb, uint4* buckets,
                                                          uint32_t* out_ids, uint4* out_pts, size_t in_stride,
                                                          size_t out_stride) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t msm = blockIdx.y;
  if (t >= nthreads) return;
  partials_chunk(in_ids + msm * in_stride, in_pts + 8 * (msm * in_stride), n_in, T, t, buckets + 8 * (msm * nb),
                 out_ids + msm * out_stride + 2 * t, out_pts + 8 * (msm * out_stride + 2 * t));
}

// Tail: once a level fits one CTA (n_in <= 2 * blockDim * T... here n_in <= kTailSlots) all remaining levels run
// inside ONE launch, ping-ponging between the two slot arrays with a block barrier per level.
constexpr uint32_t kTailThreads = 256, kTailT = 8, kTailSlots = kTailThreads * kTailT;
__global__ void __launch_bounds__(kTailThreads) k_msm_acc_partials_tail(uint32_t* ids_a, uint4* pts_a, uint32_t* ids_b,
                                                                        uint4* pts_b, uint32_t n_in, size_t nb,
                                                                        uint4* buckets, size_t stride_a, size_t stride_b) {
  size_t msm = blockIdx.x;
  uint32_t* ia = ids_a + msm * stride_a; uint4* pa = pts_a + 8 * (msm * stride_a);
  uint32_t* ib = ids_b + msm * stride_b; uint4* pb = pts_b + 8 * (msm * stride_b);
  uint4* bk = buckets + 8 * (msm * nb);
  for (;;) {
    uint32_t nth = (n_in + kTailT - 1) / kTailT;
    if (threadIdx.x < nth) partials_chunk(ia, pa, n_in, kTailT, threadIdx.x, bk, ib + 2 * threadIdx.x, pb + 8 * (2 * (size_t)threadIdx.x));
    if (nth == 1) break;
    __syncthreads();
    uint32_t* ti = ia; ia = ib; ib = ti;
    uint4* tp = pa; pa = pb; pb = tp;
    n_in = 2 * nth;
  }
}

// ---- 4: reduction -------------------------------------------------------------------------------------
// level 1: out[msm][g][k][j] = sum over virtual buckets v in segment j of group g whose weight
// (v / S + 1) has bit k set.
__global__ void __launch_bounds__(128) k_msm_reduce_bits(const uint4* buckets, MsmPlan P, uint32_t seg,
                                                         uint32_t nseg, uint4* out) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t msm = blockIdx.y;
  size_t total = (size_t)P.G * P.c * nseg;
  if (tid >= total) return;
  uint32_t j = (uint32_t)(tid % nseg);
  uint32_t k = (uint32_t)((tid / nseg) % P.c);
  uint32_t g = (uint32_t)(tid / ((size_t)nseg * P.c));
  size_t vper = (size_t)P.B * P.S;
  size_t v0 = (size_t)j * seg, v1 = v0 + seg < vper ? v0 + seg : vper;
  const uint4* base = buckets + 8 * (msm * P.nb * P.S + (size_t)g * vper);
  XYZZ acc = XYZZ::identity();
  for (size_t v = v0; v < v1; v++) {
    uint32_t weight = (uint32_t)(v / P.S) + 1;
    if ((weight >> k) & 1) acc = acc.add(XYZZ::load(base + 8 * v));
  }
  acc.store(out + 8 * (msm * total + tid));
}
// further levels: rows of `len` points -> rows of ceil(len/seg) points
__global__ void __launch_bounds__(128) k_xyzz_segsum(const uint4* in, size_t rows, uint32_t len, uint32_t seg,
                                                     uint32_t nseg, uint4* out) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (tid >= rows * nseg) return;
  size_t row = tid / nseg;
  uint32_t j = (uint32_t)(tid % nseg);
  uint32_t a = j * seg, b = a + seg < len ? a + seg : len;
  XYZZ acc = XYZZ::identity();
  for (uint32_t i = a; i < b; i++) acc = acc.add(XYZZ::load(in + 8 * (row * len + i)));
  acc.store(out + 8 * tid);
}
// ---- 4'': reduction by rows and columns (one bucket group, power-of-two bucket count) ---------------------------------
// Bucket b has weight b + 1.  With b = hi * nC + lo:   sum (b+1) B_b = nC * sum_hi hi * R_hi + sum_lo (lo+1) * C_lo,
// R_hi = sum_lo B (row sums), C_lo = sum_hi B (column sums).  The two families of plain sums cost 2 additions per bucket, as
// balanced trees (no long dependent chains); the weighted sums are then over nR + nC = 384 points instead of 32768, done
// bit-sliced.  ~2.1 additions per bucket against c/2 = 8 for the bit-sliced reduction over all buckets.
// level 1: thread (row hi, segment j) sums `seg` neighbouring buckets, thread (column lo, segment j) `seg` buckets nC apart.
__global__ void __launch_bounds__(128) k_msm_rowcol(const uint4* buckets, size_t nb, uint32_t lb, uint32_t seg, uint4* out_r, uint4* out_c) {
  const size_t msm = blockIdx.y;
  const uint32_t nC = 1u << lb, nR = (uint32_t)(nb >> lb);
  const uint32_t len_r = nC / seg, len_c = nR / seg;
  const size_t n_r = (size_t)nR * len_r;
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const uint4* base = buckets + 8 * (msm * nb);
  XYZZ acc = XYZZ::identity();
  if (tid < n_r) {
    uint32_t hi = (uint32_t)(tid / len_r), j = (uint32_t)(tid % len_r);
    const uint4* p = base + 8 * ((size_t)hi * nC + (size_t)j * seg);
    for (uint32_t i = 0; i < seg; i++) acc = acc.add(XYZZ::load(p + 8 * i));
    acc.store(out_r + 8 * (msm * n_r + tid));
  } else if (tid < n_r + (size_t)nC * len_c) {
    tid -= n_r;
    uint32_t lo = (uint32_t)(tid / len_c), j = (uint32_t)(tid % len_c);
    const uint4* p = base + 8 * ((size_t)j * seg * nC + lo);
    for (uint32_t i = 0; i < seg; i++) acc = acc.add(XYZZ::load(p + 8 * (size_t)i * nC));
    acc.store(out_c + 8 * (msm * (size_t)nC * len_c + tid));
  }
}
// item e (weight 2^e) of MSM m, segment j of kBitSegs:  sum of the C_lo whose (lo + 1) has bit e set and of the R_hi whose hi
// has bit (e - lb) set, restricted to the j-th part of each family
constexpr uint32_t kBitSegs = 16;
__global__ void __launch_bounds__(128) k_msm_rc_bits(const uint4* rows, const uint4* cols, uint32_t nR, uint32_t lb, uint32_t items, uint4* out) {
  const size_t msm = blockIdx.y;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= items * kBitSegs) return;
  const uint32_t e = tid / kBitSegs, j = tid % kBitSegs, nC = 1u << lb;
  XYZZ acc = XYZZ::identity();
  if (e <= lb) {
    const uint32_t sc = nC / kBitSegs;
    const uint4* c = cols + 8 * (msm * nC);
    for (uint32_t lo = j * sc; lo < (j + 1) * sc; lo++)
      if (((lo + 1) >> e) & 1) acc = acc.add(XYZZ::load(c + 8 * lo));
  }
  if (e >= lb) {
    const uint32_t sr = nR / kBitSegs;
    const uint4* r = rows + 8 * (msm * nR);
    for (uint32_t hi = j * sr; hi < (j + 1) * sr; hi++)
      if ((hi >> (e - lb)) & 1) acc = acc.add(XYZZ::load(r + 8 * hi));
  }
  acc.store(out + 8 * ((msm * items + e) * kBitSegs + j));
}

// ---- 4': reduction by chunked running sums ------------------------------------------------------------
// sum_v (v + w0) * X_v over a row of `len` points: the row is cut into chunks of m; thread (row, j) walks its chunk from
// the top, running += X, T += running, which leaves  T_j = sum_i (i + w0) X_(jm+i)  (i = position in the chunk) and
// S_j = sum_i X_(jm+i)  after 2 additions per point.  Then  sum = sum_j T_j + m * sum_j j * S_j : the first term is a
// plain sum, the second the same problem on the m-times shorter row S with w0 = 0.  ~2.1 additions per bucket in total
// instead of c/2 for the bit-sliced reduction above.
__global__ void __launch_bounds__(128) k_msm_chunk_sums(const uint4* in, size_t rows, uint32_t len, uint32_t m, uint32_t w0,
                                                        uint32_t nch, uint4* outT, uint4* outS) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (tid >= rows * nch) return;
  size_t row = tid / nch;
  uint32_t j = (uint32_t)(tid % nch);
  uint32_t lo = j * m, hi = lo + m < len ? lo + m : len;
  const uint4* base = in + 8 * (row * len);
  XYZZ run = XYZZ::identity(), T = XYZZ::identity();
  // positions hi-1 .. lo+1 carry weights >= 1 for either w0; position lo carries weight w0
  for (uint32_t v = hi; v-- > lo + 1;) {
    run = run.add(XYZZ::load(base + 8 * (size_t)v));
    T = T.add(run);
  }
  run = run.add(XYZZ::load(base + 8 * (size_t)lo));
  if (w0) T = T.add(run);
  T.store(outT + 8 * tid);
  run.store(outS + 8 * tid);
}
// R[row] = A_0 + m_0 * (A_1 + m_1 * (A_2 + ...)),  A_l = plain sum of level l's T row; lg[l] = log2 m_l
struct LevelPlan { const uint4* A[8]; uint32_t lg[8]; uint32_t levels; };
__global__ void __launch_bounds__(128) k_msm_combine_levels(LevelPlan L, size_t rows, uint4* out) {
  size_t row = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (row >= rows) return;
  XYZZ acc = XYZZ::load(L.A[L.levels - 1] + 8 * row);
  for (int l = (int)L.levels - 2; l >= 0; l--) {
    for (uint32_t d = 0; d < L.lg[l]; d++) acc = acc.dbl();
    acc = acc.add(XYZZ::load(L.A[l] + 8 * row));
  }
  acc.store(out + 8 * row);
}

// final: per MSM the G*c partial sums S_e (weight 2^e, e = g*c + k) are folded by a binary tree,
// V[2j] += 2^(stride) * V[2j+1], so the serial chain is (items - 1) doublings + log2(items) additions instead
// of items doublings + items additions; one CTA per MSM, the items live in shared memory.
constexpr uint32_t kFinalMaxItems = 512;
// `items` values V_e per MSM with weight 2^(shift * e)
// In a group of GPUs (`pp.world` > 1, raw output) the epilogue is also the exchange: the 128-byte XYZZ result is stored into
// every peer's window over NVLink, and the CTA that finishes last raises this rank's flag in every peer ("fuse the
// exchange into the reduction kernel's epilogue"); no collective call, no staging copy.
__global__ void __launch_bounds__(256) k_msm_final(const uint4* sums, uint32_t items, uint32_t shift, uint4* out, int raw, zk_p2p_push pp) {
  extern __shared__ uint4 fsm[];
  size_t msm = blockIdx.x;
  const uint4* s = sums + 8 * (msm * items);
  for (uint32_t q = threadIdx.x; q < items * 8; q += blockDim.x) fsm[q] = s[q];
  __syncthreads();
  for (uint32_t stride = 1; stride < items; stride <<= 1) {
    for (uint32_t j = threadIdx.x; (2 * j + 1) * stride < items; j += blockDim.x) {
      XYZZ hi = XYZZ::load(fsm + 8 * ((2 * j + 1) * stride));
      for (uint32_t d = 0; d < stride * shift; d++) hi = hi.dbl();
      XYZZ lo = XYZZ::load(fsm + 8 * (2 * j * stride));
      lo.add(hi).store(fsm + 8 * (2 * j * stride));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (raw) XYZZ::load(fsm).store(out + 8 * msm);          // XYZZ record, normalised by the caller
    else store_g1_normalised(out + 6 * msm, XYZZ::load(fsm).to_affine());
  }
  if (raw && pp.world > 1) {
    if (threadIdx.x < 8) {
      const uint4 v = fsm[threadIdx.x];
      for (uint32_t q = 0; q < pp.world; q++)
        if (q != pp.rank) ((uint4*)pp.peer_data[q])[8 * msm + threadIdx.x] = v;
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(pp.done, 1u) == gridDim.x - 1) {
      *pp.done = 0;
      __threadfence_system();
      for (uint32_t q = 0; q < pp.world; q++)
        if (q != pp.rank) *(volatile uint32_t*)pp.peer_flag[q] = pp.epoch;
    }
  }
}

// ---- window table: table[w][i] = 2^(c*w) * base[i] ------------------------------------------------------
__global__ void __launch_bounds__(128) k_msm_precompute(const uint4* bases, size_t n, uint32_t c, uint32_t W, uint4* table) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine a = Affine::load(bases + 4 * i);
  a.store(table + 4 * i);
  XYZZ p = XYZZ::from_affine(a);
  for (uint32_t w = 1; w < W; w++) {
    for (uint32_t k = 0; k < c; k++) p = p.dbl();
    Affine q = p.to_affine();
    q.store(table + 4 * ((size_t)w * n + i));
    p = XYZZ::from_affine(q);
  }
}

// out[i] = scalar[i] * base (double-and-add), affine.  ParamsKZG::setup-style fixed-base products.
__global__ void __launch_bounds__(128) k_fixed_base_mul(const uint4* scalars, size_t n, Affine base, uint4* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = Fr::load(scalars + 2 * i).from_mont();
  XYZZ acc = XYZZ::identity();
  for (int bit = 253; bit >= 0; bit--) {
    acc = acc.dbl();
    if ((s.l[bit >> 5] >> (bit & 31)) & 1) acc = acc.madd(base);
  }
  acc.to_affine().store(out + 4 * i);
}

// sum of `count` Jacobian points -> one normalised G1 (single thread; count is the GPU count)
__global__ void k_g1_sum(const uint4* pts, size_t count, uint4* out) {
  if (blockIdx.x || threadIdx.x) return;
  XYZZ acc = XYZZ::identity();
  for (size_t i = 0; i < count; i++) acc = acc.add(load_g1_as_xyzz(pts + 6 * i));
  store_g1_normalised(out, acc.to_affine());
}

// out[p] = sum over r < parts of pts[r * count + p]  (partial MSMs of the ranks of a group, one thread per output)
__global__ void k_g1_sum_strided(const uint4* pts, size_t parts, size_t count, uint4* out, int raw) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= count) return;
  XYZZ acc = XYZZ::identity();
  if (raw) {                                 // XYZZ records in and out
    for (size_t r = 0; r < parts; r++) acc = acc.add(XYZZ::load(pts + 8 * (r * count + p)));
    acc.store(out + 8 * p);
  } else {
    for (size_t r = 0; r < parts; r++) acc = acc.add(load_g1_as_xyzz(pts + 6 * (r * count + p)));
    store_g1_normalised(out + 6 * p, acc.to_affine());
  }
}

uint32_t ilog2(size_t x) { uint32_t l = 0; while (((size_t)1 << (l + 1)) <= x) l++; return l; }

MsmPlan make_plan(size_t n, size_t count, bool table, uint32_t table_c, size_t stride, int sm_count) {
  MsmPlan P; memset(&P, 0, sizeof P);
  P.n = n; P.count = count; P.stride = stride; P.sc_stride = n; P.base_off = 0;
  uint32_t lg = ilog2(n ? n : 1);
  if (table) {
    P.c = table_c;
  } else {
    int c = (int)lg - 5;            // W separate bucket sets: keep the reduction small
    if (c < 3) c = 3;
    if (c > 14) c = 14;
    P.c = (uint32_t)c;
  }
  P.W = 254 / P.c + 1;
  P.G = table ? 1 : P.W;
  P.B = 1u << (P.c - 1);
  P.nb = (size_t)P.G * P.B;
  P.ent_per_msm = n * P.W;
  (void)sm_count;
  P.S = 1;      // parallelism now comes from the chunked accumulation, not from splitting buckets
  return P;
}

// ZK_MSM_REDUCE=chunk selects the chunked running-sum reduction (fewer additions, longer dependent chains);
// the default is the bit-sliced one, which measured faster at the default window width.
bool use_chunk_reduce() {
  const char* v = getenv("ZK_MSM_REDUCE");
  return v && strcmp(v, "chunk") == 0;
}

uint32_t table_window_bits(uint32_t k) {
  if (const char* v = getenv("ZK_MSM_C")) { int c = atoi(v); if (c >= 2 && c <= 22) return (uint32_t)c; }
  // single commitments of 2^23 points and more: c = 20 (W = 13 windows instead of 16: 19 % fewer bucket additions; 2^19 buckets,
  // digits through the global-atomic path).  Measured, resident MSM, ms at c = 16 / 20 / 22 (profiles/r2_msm_window_probe.txt):
  // 2^22: 12.6 / 12.6 / 17.6, 2^23: 24.7 / 22.3 / 27.5, 2^24: 49.3 / 42.6 / 47.4, 2^25: 104.9 / 83.0 / 86.9, 2^26: 211 / 164 / 166
  if (k >= 23) return 20;
  int c = (int)k - 3;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  if (k < 14) return (uint32_t)c;
  // W = floor(254/c) + 1 moves in steps (c = 14 -> 15: 19 -> 17 windows).  Modelled work: 10 multiplications per bucket
  // addition, and per bucket ~2.1 additions in the row / column reduction plus ~3 for merging the partial sums of the
  // chunked accumulation (14 multiplications each).  The smallest c within 3 % of the minimum over [k - 3, 16] is taken.
  // Measured complete proofs (B200, ms): k = 16: c = 13 / 14 / 15: 27.6 / 26.8 / 25.7;  k = 17: 14 / 15 / 16: 42.7 / 40.2 / 41.2;
  // k = 18: 15 / 16: 71.4 / 70.8 -- the model picks 15, 15, 16 (and 16 from k = 19 on, the cap of the shared-memory digit sort).
  auto work = [&](int cc) { return (double)(254 / cc + 1) * (double)((size_t)1 << k) * 10.0 + (double)((size_t)1 << (cc - 1)) * 5.1 * 14.0; };
  double best = work(c);
  for (int cc = c + 1; cc <= 16; cc++) best = std::min(best, work(cc));
  while (c < 16 && work(c) > 1.03 * best) c++;
  return (uint32_t)c;
}

// Runs the whole pipeline for `count` MSMs over the same point array.
// `host_scalars` != nullptr: the scalars of a single MSM still live in host memory; they are uploaded into
// d_scalars in chunks on the copy stream while the histogram of the previous chunk runs.
int run_msm(zk_ctx* ctx, const void* d_scalars, const void* d_points, const MsmPlan& P, void* d_out,
            const void* host_scalars = nullptr) {
  if (P.n == 0) {
    // empty sum = identity: `count` records of the form the caller expects -- XYZZ (128 B, all zero) inside the prover /
    // group paths (ctx->msm_raw), normalised G1 (0, 1, 0) at the public ABI
    if (ctx->msm_raw) {
      ZK_CUDA(ctx, cudaMemsetAsync(d_out, 0, P.count * 128, ctx->stream));
      return ZK_OK;
    }
    std::vector<zk_g1> ids(P.count);
    memset(ids.data(), 0, ids.size() * sizeof(zk_g1));
    // Montgomery one for y
    const uint32_t one[8] = {FqParams::R0, FqParams::R1, FqParams::R2_, FqParams::R3_, FqParams::R4, FqParams::R5, FqParams::R6, FqParams::R7};
    for (auto& g : ids) memcpy(&g.y, one, 32);
    ZK_CUDA(ctx, cudaMemcpyAsync(d_out, ids.data(), ids.size() * sizeof(zk_g1), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
  }
  if (P.ent_per_msm >= ((size_t)1 << 31) || (P.G == 1 && P.W * P.stride >= ((size_t)1 << 31)))
    ZK_FAIL(ctx, ZK_ERR_ARG, "msm: n*W = %zu entries exceed the 31-bit entry index", P.ent_per_msm);
  const size_t cnt_bytes = P.count * P.nb * 4, st_bytes = P.count * (P.nb + 1) * 4;
  const size_t ent_bytes = P.count * P.ent_per_msm * 4;
  const size_t bucket_bytes = P.count * P.nb * P.S * 128;
  // reduction levels
  // short serial chains: these levels are latency bound, not throughput bound
  const uint32_t seg1 = getenv("ZK_MSM_SEG") ? (uint32_t)atoi(getenv("ZK_MSM_SEG")) : 8, seg = 4;
  size_t vper = (size_t)P.B * P.S;
  uint32_t nseg1 = (uint32_t)((vper + seg1 - 1) / seg1);
  if ((size_t)P.G * P.c > kFinalMaxItems) ZK_FAIL(ctx, ZK_ERR_ARG, "msm: %u window sums exceed the final stage", P.G * P.c);
  size_t lvl1 = P.count * P.G * P.c * nseg1;            // points after level 1
  const size_t red_bytes = lvl1 * 128;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off_cnt = 0, off_st = off_cnt + align(cnt_bytes), off_ent = off_st + align(st_bytes);
  size_t off_bk = off_ent + align(ent_bytes), off_r1 = off_bk + align(bucket_bytes), off_r2 = off_r1 + align(red_bytes);
  // chunked accumulation: T0 entries per thread at level 0, T1 slots per thread afterwards
  // Entries per thread: at most 64, and chosen so that the accumulation fills WHOLE waves of the 4 x 128-thread CTAs an SM
  // holds -- a lone 2^19 commitment is 1024 CTAs = 1.73 waves at 64 entries per thread (the second wave runs 27 % empty);
  // 56 entries make it two full, shorter waves.  Batches span tens of waves and are unaffected.
  uint32_t T0 = 64;
  if (getenv("ZK_MSM_T0")) T0 = (uint32_t)atoi(getenv("ZK_MSM_T0"));
  else {
    const size_t slots = (size_t)ctx->sm_count * 4 * 128, E = P.ent_per_msm * P.count;
    const size_t waves = (E + slots * 64 - 1) / (slots * 64);
    if (waves && waves <= 8) {
      size_t t = (E + slots * waves - 1) / (slots * waves);
      if (t < 24) t = 24;
      if (t < 64) T0 = (uint32_t)t;
    }
  }
  const uint32_t T1 = getenv("ZK_MSM_T1") ? (uint32_t)atoi(getenv("ZK_MSM_T1")) : 8;   // short serial chains: these levels are latency bound
  const uint32_t nth0 = (uint32_t)((P.ent_per_msm + T0 - 1) / T0);
  const uint32_t Tmin = T1 < kTailT ? T1 : kTailT;
  const uint32_t nth1 = (2 * nth0 + Tmin - 1) / Tmin;
  const size_t slotA = (size_t)2 * nth0, slotB = (size_t)2 * nth1 + 2;
  size_t off_sa_id = off_r2 + align(red_bytes / seg + 128 * P.count * P.G * P.c);
  size_t off_sa_pt = off_sa_id + align(P.count * slotA * 4), off_sb_id = off_sa_pt + align(P.count * slotA * 128);
  size_t off_sb_pt = off_sb_id + align(P.count * slotB * 4);
  // Batched-affine pairing round: opt-in (ZK_MSM_BA=1).  Measured on B200 (k = 19, batch of 28, profiles/
  // r1_pair_add.txt): 0.64 ms for the round + 0.63 ms XYZZ on the halved list = 1.27 ms, against 1.26 ms for XYZZ alone.
  // The round saves 3.5 of 10 multiplications per addition but moves ~320 B of mostly randomly gathered DRAM
  // traffic per addition (two passes around the shared inversion), which is what bounds it.
  const bool use_ba = getenv("ZK_MSM_BA") ? atoi(getenv("ZK_MSM_BA")) != 0 : false;
  const size_t out_cap = (P.ent_per_msm + P.nb) / 2 + 1;
  uint32_t baM = 8;
  {
    size_t want_ctas = (size_t)ctx->sm_count * 3;
    while (baM < 128 && P.count * out_cap / ((size_t)kPairThreads * baM) > want_ctas) baM <<= 1;
    if (getenv("ZK_MSM_BA_M")) baM = (uint32_t)atoi(getenv("ZK_MSM_BA_M"));
  }
  const unsigned ba_grid = zk_div_up(out_cap, (size_t)kPairThreads * baM);
  const size_t ba_threads = (size_t)ba_grid * kPairThreads;
  size_t off_ba_cnt = off_sb_pt + align(P.count * slotB * 128);
  size_t off_ba_st = off_ba_cnt + align(use_ba ? cnt_bytes : 0);
  size_t off_ba_pts = off_ba_st + align(use_ba ? st_bytes : 0);
  size_t off_ba_scr = off_ba_pts + align(use_ba ? P.count * out_cap * 64 : 0);
  // shared-memory counters when one MSM's bucket set fits (ZK_MSM_DIGITS=global keeps the per-digit global atomics)
  static const bool digits_global = getenv("ZK_MSM_DIGITS") && !strcmp(getenv("ZK_MSM_DIGITS"), "global");
  const size_t dsm = P.nb * 4;
  const bool digits_smem = !digits_global && dsm <= 200 * 1024 && P.n >= 8192;
  size_t dchunk = 0;
  if (digits_smem) {
    // per device, so not cached in a static (a process may hold contexts on several GPUs)
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_digits_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_digits_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // one CTA is resident per SM (the counters fill its shared memory).  CTAs per SM over the whole batch, measured on the
    // complete proof (scatter ms per proof): 1: 8.6, 2: 6.9, 4: 5.2, 8: 3.7, 12: 3.3, 16-32: 3.3 -- short chunks balance the
    // SMs and keep the CTAs' reserved ranges of a bucket close in time; below 4096 scalars walking the counters dominates
    static const size_t waves = getenv("ZK_MSM_DIGIT_CTAS") ? (size_t)atoi(getenv("ZK_MSM_DIGIT_CTAS")) : 12;
    size_t per_msm = ((size_t)ctx->sm_count * waves + P.count - 1) / P.count;
    dchunk = (P.n + per_msm - 1) / per_msm;
    if (dchunk < 4096) dchunk = 4096;
    dchunk = (dchunk + kDigitsSmemThreads - 1) / kDigitsSmemThreads * kDigitsSmemThreads;
  }
  size_t off_cta = off_ba_scr + align(use_ba ? P.count * baM * ba_threads * 32 : 0);
  size_t total = off_cta + align(digits_smem ? zk_div_up(P.n, dchunk) * P.count * P.nb * 4 : 0);
  ZK_TRY(zk_scratch_reserve(ctx, ctx->msm_work, total));
  char* w = (char*)ctx->msm_work.p;
  uint32_t* counts = (uint32_t*)(w + off_cnt);
  uint32_t* starts = (uint32_t*)(w + off_st);
  uint32_t* entries = (uint32_t*)(w + off_ent);
  uint32_t* cta_base = (uint32_t*)(w + off_cta);
  uint4* buckets = (uint4*)(w + off_bk);
  uint4* r1 = (uint4*)(w + off_r1);
  uint4* r2 = (uint4*)(w + off_r2);

  ZK_CUDA(ctx, cudaMemsetAsync(counts, 0, cnt_bytes, ctx->stream));
  auto launch_digits = [&](bool scatter, size_t lo, size_t hi) -> int {
    if (digits_smem) {
      dim3 g(zk_div_up(hi - lo, dchunk), (unsigned)P.count);
      if (scatter) {
        ZK_LAUNCH(ctx, "msm_digits_scatter", k_msm_digits_smem<true>, g, kDigitsSmemThreads, dsm, (const uint4*)d_scalars, P, counts,
                  (const uint32_t*)starts, entries, lo, hi, dchunk, cta_base);
      } else {
        ZK_LAUNCH(ctx, "msm_digits_hist", k_msm_digits_smem<false>, g, kDigitsSmemThreads, dsm, (const uint4*)d_scalars, P, counts,
                  (const uint32_t*)nullptr, (uint32_t*)nullptr, lo, hi, dchunk, cta_base);
      }
    } else {
      dim3 g(zk_div_up(hi - lo, 256), (unsigned)P.count);
      if (scatter) {
        ZK_LAUNCH(ctx, "msm_digits_scatter", k_msm_digits<true>, g, 256, 0, (const uint4*)d_scalars, P, counts, (const uint32_t*)starts,
                  entries, lo, hi);
      } else {
        ZK_LAUNCH(ctx, "msm_digits_hist", k_msm_digits<false>, g, 256, 0, (const uint4*)d_scalars, P, counts, (const uint32_t*)nullptr,
                  (uint32_t*)nullptr, lo, hi);
      }
    }
    return ZK_OK;
  };
  if (host_scalars && P.count == 1 && P.n >= 4096) {
    ZK_TRY(zk_copy_stream(ctx));
    const size_t chunks = 4;
    size_t per = (P.n + chunks - 1) / chunks;
    if (digits_smem) per = (per + dchunk - 1) / dchunk * dchunk;      // pieces start on chunk boundaries
    for (size_t c = 0; c < chunks; c++) {
      size_t lo = c * per, hi = lo + per < P.n ? lo + per : P.n;
      if (lo >= hi) break;
      ZK_CUDA(ctx, cudaMemcpyAsync((char*)d_scalars + lo * 32, (const char*)host_scalars + lo * 32, (hi - lo) * 32,
                                   cudaMemcpyHostToDevice, ctx->copy_stream));
      ZK_CUDA(ctx, cudaEventRecord(ctx->copy_event, ctx->copy_stream));
      ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_event, 0));
      ZK_TRY(launch_digits(false, lo, hi));
    }
  } else {
    if (host_scalars)
      ZK_CUDA(ctx, cudaMemcpyAsync((void*)d_scalars, host_scalars, P.count * P.n * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZK_TRY(launch_digits(false, 0, P.n));
  }
  if (!ctx->d_counters) {
    ZK_CUDA(ctx, cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
    ZK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
  }
  ZK_LAUNCH(ctx, "msm_scan", k_msm_scan, (unsigned)P.count, 1024, 0, counts, starts, P.nb, ctx->d_counters);
  ZK_TRY(launch_digits(true, 0, P.n));
  ZK_CUDA(ctx, cudaMemsetAsync(buckets, 0, bucket_bytes, ctx->stream));     // identity everywhere
  {
    uint32_t* ida = (uint32_t*)(w + off_sa_id); uint4* pta = (uint4*)(w + off_sa_pt);
    uint32_t* idb = (uint32_t*)(w + off_sb_id); uint4* ptb = (uint4*)(w + off_sb_pt);
    dim3 ga(zk_div_up(nth0, 128), (unsigned)P.count);
    if (use_ba) {
      uint32_t* ocnt = (uint32_t*)(w + off_ba_cnt); uint32_t* ost = (uint32_t*)(w + off_ba_st);
      uint4* opts = (uint4*)(w + off_ba_pts); uint4* scr = (uint4*)(w + off_ba_scr);
      dim3 gc(zk_div_up(P.nb, 256), (unsigned)P.count);
      ZK_LAUNCH(ctx, "msm_pair_counts", k_pair_counts, gc, 256, 0, (const uint32_t*)starts, P.nb, ocnt);
      ZK_LAUNCH(ctx, "msm_scan", k_msm_scan, (unsigned)P.count, 1024, 0, ocnt, ost, P.nb, (unsigned long long*)nullptr);
      dim3 gb(ba_grid, (unsigned)P.count);
      ZK_LAUNCH(ctx, "msm_pair_add", k_pair_add, gb, kPairThreads, 0, (const uint4*)d_points, (const uint32_t*)entries,
                (const uint32_t*)starts, (const uint32_t*)ost, P, baM, out_cap, opts, scr);
      ZK_LAUNCH(ctx, "msm_accumulate", k_msm_accumulate<true>, ga, 128, 0, (const uint4*)opts, out_cap, (const uint32_t*)nullptr,
                (const uint32_t*)ost, P, T0, nth0, buckets, ida, pta);
    } else {
      ZK_LAUNCH(ctx, "msm_accumulate", k_msm_accumulate<false>, ga, 128, 0, (const uint4*)d_points, (size_t)0,
                (const uint32_t*)entries, (const uint32_t*)starts, P, T0, nth0, buckets, ida, pta);
    }
    uint32_t n_in = 2 * nth0;
    size_t stride_in = slotA, stride_out = slotB;
    while (n_in > kTailSlots) {
      uint32_t nth = (n_in + T1 - 1) / T1;
      dim3 gp(zk_div_up(nth, 128), (unsigned)P.count);
      ZK_LAUNCH(ctx, "msm_acc_partials", k_msm_acc_partials, gp, 128, 0, (const uint32_t*)ida, (const uint4*)pta, n_in, T1,
                nth, P.nb, buckets, idb, ptb, stride_in, stride_out);
      uint32_t* ti = ida; ida = idb; idb = ti;
      uint4* tp = pta; pta = ptb; ptb = tp;
      size_t ts = stride_in; stride_in = stride_out; stride_out = ts;
      n_in = 2 * nth;
    }
    ZK_LAUNCH(ctx, "msm_acc_partials", k_msm_acc_partials_tail, (unsigned)P.count, kTailThreads, 0, ida, pta, idb, ptb, n_in,
              P.nb, buckets, stride_in, stride_out);
  }
  if (use_chunk_reduce()) {
    // chunked running sums: level 0 over the buckets of every (msm, group) row, then over the rows of chunk totals
    const size_t rws = P.count * P.G;
    const uint32_t m0 = getenv("ZK_MSM_M0") ? (uint32_t)atoi(getenv("ZK_MSM_M0")) : 8, m1 = 8, sseg = 8;
    LevelPlan L; memset(&L, 0, sizeof L);
    // scratch inside r1 / r2: each level needs T and S rows of nch points and the plain-sum ladder of T
    uint4* cur = buckets; uint32_t len = (uint32_t)(P.B * P.S);
    char* arena = (char*)r1; size_t arena_off = 0;
    auto take = [&](size_t points) { uint4* p = (uint4*)(arena + arena_off); arena_off += ((points * 128 + 255) & ~(size_t)255); return p; };
    uint32_t level = 0, w0 = 1;
    while (true) {
      uint32_t m = level == 0 ? m0 : m1;
      uint32_t nch = (len + m - 1) / m;
      uint4* T = take(rws * nch); uint4* S = take(rws * nch);
      ZK_LAUNCH(ctx, "msm_reduce_bits", k_msm_chunk_sums, zk_div_up(rws * nch, 128), 128, 0, (const uint4*)cur, rws, len, m, w0, nch, T, S);
      // plain sum of the T row
      uint32_t tl = nch; uint4* tsrc = T;
      while (tl > 1) {
        uint32_t ns = (tl + sseg - 1) / sseg;
        uint4* tdst = take(rws * ns);
        ZK_LAUNCH(ctx, "msm_segsum", k_xyzz_segsum, zk_div_up(rws * ns, 128), 128, 0, (const uint4*)tsrc, rws, tl, sseg, ns, tdst);
        tsrc = tdst; tl = ns;
      }
      L.A[level] = tsrc;
      uint32_t lg = 0; while ((1u << lg) < m) lg++;
      L.lg[level] = lg;
      level++;
      if (nch == 1 || level == 8) { if (nch != 1) ZK_FAIL(ctx, ZK_ERR_ARG, "msm: reduction needs more than 8 levels"); break; }
      cur = S; len = nch; w0 = 0;
    }
    L.levels = level;
    if (arena_off + rws * 128 + 256 > off_sa_id - off_r1) ZK_FAIL(ctx, ZK_ERR_STATE, "msm: reduction arena overflow");
    uint4* R = take(rws);
    ZK_LAUNCH(ctx, "msm_final", k_msm_combine_levels, zk_div_up(rws, 128), 128, 0, L, rws, R);
    size_t fsmem = (size_t)P.G * 128;
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFinalMaxItems * 128)));
    ZK_LAUNCH(ctx, "msm_final", k_msm_final, (unsigned)P.count, 256, fsmem, (const uint4*)R, P.G, P.c, (uint4*)d_out, ctx->msm_raw ? 1 : 0, (ctx->msm_raw && ctx->p2p) ? *ctx->p2p : zk_p2p_push());
    return ZK_OK;
  }
  // rows and columns (ZK_MSM_REDUCE=bits keeps the bit-sliced reduction over all buckets, which also serves G > 1)
  {
    const uint32_t lb = (P.c - 1 + 1) / 2;
    const size_t nC = (size_t)1 << lb, nR = P.nb >> lb;
    const char* rv = getenv("ZK_MSM_REDUCE");
    if (P.G == 1 && P.S == 1 && P.nb == ((size_t)1 << (P.c - 1)) && nR >= kBitSegs && nC >= kBitSegs && !(rv && !strcmp(rv, "bits"))) {
      const uint32_t sg = 8, len_r = (uint32_t)(nC / sg), len_c = (uint32_t)(nR / sg), items = P.c - 1;
      size_t used = 0;                                                  // bump allocation in r1 (XYZZ points)
      auto take = [&](size_t pts) { uint4* q = r1 + 8 * used; used += pts; return q; };
      uint4* part_r = take(P.count * nR * len_r);
      uint4* part_c = take(P.count * nC * len_c);
      dim3 g1(zk_div_up(nR * len_r + nC * len_c, 128), (unsigned)P.count);
      ZK_LAUNCH(ctx, "msm_reduce_rowcol", k_msm_rowcol, g1, 128, 0, (const uint4*)buckets, P.nb, lb, sg, part_r, part_c);
      auto tree = [&](uint4* src, size_t rws, uint32_t len, uint4** res) -> int {
        while (len > 1) {
          uint32_t ns = (len + seg - 1) / seg;
          uint4* dst = take(rws * ns);
          ZK_LAUNCH(ctx, "msm_segsum", k_xyzz_segsum, zk_div_up(rws * ns, 128), 128, 0, (const uint4*)src, rws, len, seg, ns, dst);
          src = dst; len = ns;
        }
        *res = src;
        return ZK_OK;
      };
      uint4 *R = nullptr, *C = nullptr, *I = nullptr;
      ZK_TRY(tree(part_r, P.count * nR, len_r, &R));
      ZK_TRY(tree(part_c, P.count * nC, len_c, &C));
      uint4* bits = take(P.count * items * kBitSegs);
      dim3 g2(zk_div_up((size_t)items * kBitSegs, 128), (unsigned)P.count);
      ZK_LAUNCH(ctx, "msm_reduce_rowcol", k_msm_rc_bits, g2, 128, 0, (const uint4*)R, (const uint4*)C, (uint32_t)nR, lb, items, bits);
      ZK_TRY(tree(bits, P.count * items, kBitSegs, &I));
      if (used * 128 > red_bytes) ZK_FAIL(ctx, ZK_ERR_STATE, "msm: reduction arena overflow");
      ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFinalMaxItems * 128)));
      ZK_LAUNCH(ctx, "msm_final", k_msm_final, (unsigned)P.count, 256, (size_t)items * 128, (const uint4*)I, items, 1u, (uint4*)d_out, ctx->msm_raw ? 1 : 0, (ctx->msm_raw && ctx->p2p) ? *ctx->p2p : zk_p2p_push());
      return ZK_OK;
    }
  }
  size_t rows = P.count * P.G * P.c;
  dim3 gr(zk_div_up((size_t)P.G * P.c * nseg1, 128), (unsigned)P.count);
  ZK_LAUNCH(ctx, "msm_reduce_bits", k_msm_reduce_bits, gr, 128, 0, (const uint4*)buckets, P, seg1, nseg1, r1);
  uint32_t len = nseg1;
  uint4 *src = r1, *dst = r2;
  while (len > 1) {
    uint32_t nseg = (len + seg - 1) / seg;
    ZK_LAUNCH(ctx, "msm_segsum", k_xyzz_segsum, zk_div_up(rows * nseg, 128), 128, 0, (const uint4*)src, rows, len, seg, nseg, dst);
    uint4* t = src; src = dst; dst = t;
    len = nseg;
  }
  {
    size_t fsmem = (size_t)P.G * P.c * 128;
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFinalMaxItems * 128)));
    ZK_LAUNCH(ctx, "msm_final", k_msm_final, (unsigned)P.count, 256, fsmem, (const uint4*)src, P.G * P.c, 1u, (uint4*)d_out, ctx->msm_raw ? 1 : 0, (ctx->msm_raw && ctx->p2p) ? *ctx->p2p : zk_p2p_push());
  }
  return ZK_OK;
}

int build_table(zk_ctx* ctx, zk_srs_set& s, size_t n, uint32_t k) {
  uint32_t c = table_window_bits(k), W = 254 / c + 1;
  size_t bytes = (size_t)W * n * 64;
  if (cudaMalloc(&s.table, bytes) != cudaSuccess) {
    cudaGetLastError();
    ZK_FAIL(ctx, ZK_ERR_OOM, "window table: cannot allocate %zu bytes", bytes);
  }
  s.table_c = (int)c; s.table_w = (int)W;
  ZK_LAUNCH(ctx, "msm_precompute", k_msm_precompute, zk_div_up(n, 128), 128, 0, (const uint4*)s.bases, n, c, W, (uint4*)s.table);
  return ZK_OK;
}

}  // namespace

// Both base sets (and their window tables) are built into locals and swapped into the context only when everything
// succeeded: a failed allocation leaves the previous SRS untouched (and nothing leaked).
static int srs_install(zk_ctx* ctx, uint32_t k, const zk_g1_affine* g, const zk_g1_affine* g_lagrange, int precompute_windows,
                       cudaMemcpyKind kind) {
  size_t n = (size_t)1 << k;
  zk_srs_set fresh[2];
  const zk_g1_affine* src[2] = {g, g_lagrange};
  int rc = ZK_OK;
  // the same array for both sets (single-MSM sweeps, sharded slices): one copy and one window table serve both
  const int nsets = (g == g_lagrange) ? 1 : 2;
  for (int i = 0; i < nsets && rc == ZK_OK; i++) {
    if (cudaMalloc(&fresh[i].bases, n * 64) != cudaSuccess) {
      cudaGetLastError();
      ctx->err = "zk_srs_upload: cannot allocate the base set"; rc = ZK_ERR_OOM; break;
    }
    if (cudaMemcpyAsync(fresh[i].bases, src[i], n * 64, kind, ctx->stream) != cudaSuccess) {
      ctx->err = std::string("zk_srs_upload: copy: ") + cudaGetErrorString(cudaGetLastError()); rc = ZK_ERR_CUDA; break;
    }
    if (precompute_windows) rc = build_table(ctx, fresh[i], n, k);
  }
  if (rc == ZK_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = std::string("zk_srs_upload: ") + cudaGetErrorString(cudaGetLastError()); rc = ZK_ERR_CUDA;
  }
  if (rc != ZK_OK) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& f : fresh) { if (f.bases) cudaFree(f.bases); if (f.table) cudaFree(f.table); }
    return rc;
  }
  if (nsets == 1) { fresh[1] = fresh[0]; fresh[1].alias = true; }
  zk_srs_set* sets[2] = {&ctx->srs_g, &ctx->srs_lagrange};
  for (int i = 0; i < 2; i++) {
    if (!ctx->srs_borrowed && !sets[i]->alias) { if (sets[i]->bases) cudaFree(sets[i]->bases); if (sets[i]->table) cudaFree(sets[i]->table); }
    *sets[i] = fresh[i];
  }
  ctx->srs_k = k; ctx->srs_n = n; ctx->srs_borrowed = false;
  return ZK_OK;
}

extern "C" int zk_srs_upload(zk_ctx* ctx, uint32_t k, const zk_g1_affine* g, const zk_g1_affine* g_lagrange,
                             int precompute_windows) {
  if (!ctx) return ZK_ERR_ARG;
  if (!g || !g_lagrange || k > 27) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_srs_upload: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  return srs_install(ctx, k, g, g_lagrange, precompute_windows, cudaMemcpyHostToDevice);
}

// Same as zk_srs_upload with the two base sets already in device memory (copied).
extern "C" int zk_srs_upload_dev(zk_ctx* ctx, uint32_t k, const zk_g1_affine* d_g, const zk_g1_affine* d_g_lagrange,
                                 int precompute_windows) {
  if (!ctx) return ZK_ERR_ARG;
  if (!d_g || !d_g_lagrange || k > 27) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_srs_upload_dev: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  return srs_install(ctx, k, d_g, d_g_lagrange, precompute_windows, cudaMemcpyDeviceToDevice);
}

// out[0] = bucket entries (non-zero signed digits = group additions of the accumulation), out[1] = MSMs run on this
// context since creation / the last reset.  Synchronises the context stream.
extern "C" int zk_ctx_counters(zk_ctx* ctx, uint64_t out[4], int reset) {
  if (!ctx) return ZK_ERR_ARG;
  if (!out) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_ctx_counters: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  memset(out, 0, 4 * sizeof(uint64_t));
  if (!ctx->d_counters) return ZK_OK;
  ZK_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_counters, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (reset) ZK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}

extern "C" int zk_srs_window(zk_ctx* ctx, int which, uint32_t* c_out, uint32_t* w_out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!c_out || !w_out || (which != ZK_BASES_G && which != ZK_BASES_G_LAGRANGE)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_srs_window: bad argument");
  const zk_srs_set& s = which == ZK_BASES_G ? ctx->srs_g : ctx->srs_lagrange;
  *c_out = s.table ? (uint32_t)s.table_c : 0u; *w_out = s.table ? (uint32_t)s.table_w : 0u;
  return ZK_OK;
}

extern "C" int zk_msm_batch_dev(zk_ctx* ctx, const zk_fr* d_scalars, int which, size_t n, size_t count, zk_g1* d_out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!d_out || (n && !d_scalars) || (which != ZK_BASES_G && which != ZK_BASES_G_LAGRANGE))
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_msm_batch_dev: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!count) return ZK_OK;
  zk_srs_set& s = which == ZK_BASES_G ? ctx->srs_g : ctx->srs_lagrange;
  if (!s.bases) ZK_FAIL(ctx, ZK_ERR_STATE, "msm: no SRS uploaded (zk_srs_upload)");
  if (n > ctx->srs_n) ZK_FAIL(ctx, ZK_ERR_ARG, "msm: n = %zu exceeds the SRS size %zu", n, ctx->srs_n);
  // the window table is indexed w * srs_n + i; shorter commitments use the first n columns
  bool table = s.table != nullptr;
  MsmPlan P = make_plan(n, count, table, (uint32_t)s.table_c, ctx->srs_n, ctx->sm_count);
  return run_msm(ctx, d_scalars, table ? s.table : s.bases, P, d_out);
}

extern "C" int zk_msm_dev(zk_ctx* ctx, const zk_fr* d_scalars, const zk_g1_affine* d_bases, int which, size_t n, zk_g1* d_out) {
  if (!ctx) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (which == ZK_BASES_HOST) {
    if (!d_out || (n && (!d_scalars || !d_bases))) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_msm_dev: bad argument");
    MsmPlan P = make_plan(n, 1, false, 0, n, ctx->sm_count);
    return run_msm(ctx, d_scalars, d_bases, P, d_out);
  }
  return zk_msm_batch_dev(ctx, d_scalars, which, n, 1, d_out);
}

extern "C" int zk_msm(zk_ctx* ctx, const zk_fr* scalars, const zk_g1_affine* bases, int which, size_t n, zk_g1* out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!out || (n && !scalars) || (which == ZK_BASES_HOST && n && !bases)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_msm: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->msm_scalars, n * 32 + 32));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->msm_out, sizeof(zk_g1)));
  if (which == ZK_BASES_HOST) {
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->msm_scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZK_TRY(zk_scratch_reserve(ctx, ctx->msm_bases, n * 64 + 64));
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->msm_bases.p, bases, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    ZK_TRY(zk_msm_dev(ctx, (const zk_fr*)ctx->msm_scalars.p, (const zk_g1_affine*)ctx->msm_bases.p, which, n, (zk_g1*)ctx->msm_out.p));
  } else {
    if (which != ZK_BASES_G && which != ZK_BASES_G_LAGRANGE) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_msm: bad base selector");
    zk_srs_set& s = which == ZK_BASES_G ? ctx->srs_g : ctx->srs_lagrange;
    if (!s.bases) ZK_FAIL(ctx, ZK_ERR_STATE, "msm: no SRS uploaded (zk_srs_upload)");
    if (n > ctx->srs_n) ZK_FAIL(ctx, ZK_ERR_ARG, "msm: n = %zu exceeds the SRS size %zu", n, ctx->srs_n);
    bool table = s.table != nullptr;
    MsmPlan P = make_plan(n, 1, table, (uint32_t)s.table_c, ctx->srs_n, ctx->sm_count);
    ZK_TRY(run_msm(ctx, ctx->msm_scalars.p, table ? s.table : s.bases, P, ctx->msm_out.p, scalars));
  }
  ZK_CUDA(ctx, cudaMemcpyAsync(out, ctx->msm_out.p, sizeof(zk_g1), cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}

// `count` commitments restricted to the bases [base_off, base_off + n_sub) of the resident SRS: the scalars of
// commitment m are d_scalars[m * sc_stride + base_off ..].  Internal (group-sharded single commitments).
int zk_msm_range_batch_dev(zk_ctx* ctx, const zk_fr* d_scalars, int which, size_t sc_stride, size_t base_off, size_t n_sub,
                           size_t count, zk_g1* d_out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!d_out || !d_scalars || (which != ZK_BASES_G && which != ZK_BASES_G_LAGRANGE)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_msm_range: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!count) return ZK_OK;
  zk_srs_set& s = which == ZK_BASES_G ? ctx->srs_g : ctx->srs_lagrange;
  if (!s.table) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_msm_range: needs the window table (zk_srs_upload with precompute_windows)");
  if (base_off + n_sub > ctx->srs_n) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_msm_range: range exceeds the SRS");
  MsmPlan P = make_plan(n_sub, count, true, (uint32_t)s.table_c, ctx->srs_n, ctx->sm_count);
  P.sc_stride = sc_stride; P.base_off = base_off;
  return run_msm(ctx, d_scalars, s.table, P, d_out);
}
int zk_g1_sum_strided_dev(zk_ctx* ctx, const zk_g1* d_points, size_t parts, size_t count, zk_g1* d_out) {
  ZK_LAUNCH(ctx, "g1_sum", k_g1_sum_strided, zk_div_up(count, 32), 32, 0, (const uint4*)d_points, parts, count, (uint4*)d_out, ctx->msm_raw ? 1 : 0);
  return ZK_OK;
}

// sum of `parts` XYZZ records (the partial results of a base-range-sharded MSM) -> one normalised G1
__global__ void k_xyzz_sum_normalise(const uint4* pts, size_t parts, uint4* out) {
  if (blockIdx.x || threadIdx.x) return;
  XYZZ acc = XYZZ::identity();
  for (size_t r = 0; r < parts; r++) acc = acc.add(XYZZ::load(pts + 8 * r));
  store_g1_normalised(out, acc.to_affine());
}
int zk_xyzz_sum_normalised_dev(zk_ctx* ctx, const void* d_xyzz, size_t parts, zk_g1* d_out) {
  ZK_LAUNCH(ctx, "g1_sum", k_xyzz_sum_normalise, 1, 32, 0, (const uint4*)d_xyzz, parts, (uint4*)d_out);
  return ZK_OK;
}

extern "C" int zk_g1_sum_dev(zk_ctx* ctx, const zk_g1* d_points, size_t count, zk_g1* d_out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!d_out || (count && !d_points)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_sum_dev: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_LAUNCH(ctx, "g1_sum", k_g1_sum, 1, 32, 0, (const uint4*)d_points, count, (uint4*)d_out);
  return ZK_OK;
}
extern "C" int zk_g1_sum(zk_ctx* ctx, const zk_g1* points, size_t count, zk_g1* out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!out || (count && !points)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_sum: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, (count + 1) * sizeof(zk_g1)));
  zk_g1* d = (zk_g1*)ctx->misc.p;
  ZK_CUDA(ctx, cudaMemcpyAsync(d + 1, points, count * sizeof(zk_g1), cudaMemcpyHostToDevice, ctx->stream));
  ZK_TRY(zk_g1_sum_dev(ctx, d + 1, count, d));
  ZK_CUDA(ctx, cudaMemcpyAsync(out, d, sizeof(zk_g1), cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}

// d_out[i] = d_scalars[i] * base as affine points (ParamsKZG::setup computes g[i] = s^i * G this way)
extern "C" int zk_g1_fixed_base_mul_dev(zk_ctx* ctx, const zk_g1_affine* base, const zk_fr* d_scalars, size_t n,
                                        zk_g1_affine* d_out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!base || (n && (!d_scalars || !d_out))) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_fixed_base_mul_dev: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!n) return ZK_OK;
  Affine b; memcpy(b.x.l, &base->x, 32); memcpy(b.y.l, &base->y, 32);
  ZK_LAUNCH(ctx, "fixed_base_mul", k_fixed_base_mul, zk_div_up(n, 128), 128, 0, (const uint4*)d_scalars, n, b, (uint4*)d_out);
  return ZK_OK;
}
