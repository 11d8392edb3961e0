// prover_kernels.cuh -- device kernels of the prover's n-sized loops that sit between the MSM / NTT /
// evaluate_h calls of create_proof.  They replace, on the device:
//   plonk/lookup/prover.rs       compress_expressions, permute_expression_pair, commit_product
//   plonk/permutation/prover.rs  Argument::commit (grand product of a column set)
//   arithmetic.rs                eval_polynomial, kate_division; shplonk/prover.rs linear combinations
// ([UPSTREAM] halo2_proofs 0.2.0 @ c254c75f, Cargo.lock:1314-1316; reached from create_proof,
// circuits/src/sgx_dcap_verifier.rs:814-822).  All of them are element-wise maps, scans, reductions or a stable
// radix sort over 2^k field elements: bound by HBM traffic and launch latency, not by the integer pipe.
#pragma once
#include "evalh.cuh"

namespace zk {

// ---- monoids for the three-phase scan --------------------------------------------------------------------
struct FrMulMonoid {
  typedef Fr T;
  __device__ static Fr identity() { return Fr::one(); }
  __device__ static Fr combine(const Fr& a, const Fr& b) { return a * b; }
  __device__ static Fr load(const void* p, size_t i) { return Fr::load((const uint4*)p + 2 * i); }
  __device__ static void store(void* p, size_t i, const Fr& v) { v.store((uint4*)p + 2 * i); }
};
struct FrAddMonoid {
  typedef Fr T;
  __device__ static Fr identity() { return Fr::zero(); }
  __device__ static Fr combine(const Fr& a, const Fr& b) { return a + b; }
  __device__ static Fr load(const void* p, size_t i) { return Fr::load((const uint4*)p + 2 * i); }
  __device__ static void store(void* p, size_t i, const Fr& v) { v.store((uint4*)p + 2 * i); }
};
struct U32AddMonoid {
  typedef uint32_t T;
  __device__ static uint32_t identity() { return 0; }
  __device__ static uint32_t combine(uint32_t a, uint32_t b) { return a + b; }
  __device__ static uint32_t load(const void* p, size_t i) { return ((const uint32_t*)p)[i]; }
  __device__ static void store(void* p, size_t i, uint32_t v) { ((uint32_t*)p)[i] = v; }
};

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

// Block-wide inclusive scan of one value per thread through shared memory (Hillis-Steele); returns the
// EXCLUSIVE prefix of the calling thread and leaves the block total in *total.
template <class M>
__device__ typename M::T block_exclusive(typename M::T v, typename M::T* sh, typename M::T* total) {
  typedef typename M::T T;
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int off = 1; off < kScanThreads; off <<= 1) {
    T x = sh[t];
    if (t >= off) x = M::combine(sh[t - off], x);
    __syncthreads();
    sh[t] = x;
    __syncthreads();
  }
  T excl = t ? sh[t - 1] : M::identity();
  *total = sh[kScanThreads - 1];
  __syncthreads();
  return excl;
}

// Logical element i of a scan lives at physical index (reverse ? n-1-i : i).
__device__ __forceinline__ size_t scan_phys(size_t i, size_t n, int reverse) { return reverse ? n - 1 - i : i; }

// The scans are SEGMENTED: blockIdx.y selects one of several independent arrays of n elements (segment g starts at
// element g * stride), so that the lookups' / permutation sets' products run in one launch.
// phase 1: reduction of every tile
template <class M>
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const void* in, size_t n, size_t stride, int reverse, void* tile_totals) {
  typedef typename M::T T;
  __shared__ T sh[kScanThreads];
  const size_t seg = (size_t)blockIdx.y * stride;
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  T acc = M::identity();
#pragma unroll
  for (int k = 0; k < kScanItems; k++)
    if (base + k < n) acc = M::combine(acc, M::load(in, seg + scan_phys(base + k, n, reverse)));
  T total;
  block_exclusive<M>(acc, sh, &total);
  if (threadIdx.x == 0) M::store(tile_totals, (size_t)blockIdx.y * gridDim.x + blockIdx.x, total);
}
// phase 2: exclusive scan of each segment's tile totals, in place, one block per segment; seeds (optional, device,
// one per segment) multiply in front
template <class M>
__global__ void __launch_bounds__(kScanThreads) k_scan_totals(void* tile_totals, size_t ntiles, const void* seeds) {
  typedef typename M::T T;
  __shared__ T sh[kScanThreads];
  const size_t seg = (size_t)blockIdx.x * ntiles;
  T carry = seeds ? M::load(seeds, blockIdx.x) : M::identity();
  for (size_t base = 0; base < ntiles; base += kScanThreads) {
    size_t i = base + threadIdx.x;
    T v = i < ntiles ? M::load(tile_totals, seg + i) : M::identity();
    T total;
    T excl = block_exclusive<M>(v, sh, &total);
    if (i < ntiles) M::store(tile_totals, seg + i, M::combine(carry, excl));
    carry = M::combine(carry, total);
  }
}
// phase 3: out[i] = prefix(tile) * prefix(thread) * local; exclusive => out[i] excludes element i
template <class M>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const void* in, void* out, size_t n, size_t stride_in, size_t stride_out,
                                                             int reverse, int exclusive, const void* tile_prefix) {
  typedef typename M::T T;
  __shared__ T sh[kScanThreads];
  const size_t seg_in = (size_t)blockIdx.y * stride_in, seg_out = (size_t)blockIdx.y * stride_out;
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  T v[kScanItems];
  T acc = M::identity();
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    v[k] = base + k < n ? M::load(in, seg_in + scan_phys(base + k, n, reverse)) : M::identity();
    acc = M::combine(acc, v[k]);
  }
  T total;
  T excl = block_exclusive<M>(acc, sh, &total);
  T run = M::combine(M::load(tile_prefix, (size_t)blockIdx.y * gridDim.x + blockIdx.x), excl);
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    T incl = M::combine(run, v[k]);
    if (base + k < n) M::store(out, seg_out + scan_phys(base + k, n, reverse), exclusive ? run : incl);
    run = incl;
  }
}

// ---- element-wise kernels ---------------------------------------------------------------------------------
// omega^i from the extended twiddle table (omega_ext^e, e < ext_n/2): omega = omega_ext^(ext_n/n)
__device__ __forceinline__ Fr omega_pow(const uint4* tw, size_t i, uint32_t shift, size_t half_ext) {
  size_t e = i << shift;
  Fr w = Fr::load(tw + 2 * (e & (half_ext - 1)));
  return (e & half_ext) ? w.neg() : w;
}

struct PermCols {
  const uint4* values[8];
  const uint4* sigma[8];
  uint32_t ncols;
  uint32_t pad;
};
struct PermNum { Fr dbeta[8]; };
// one column set per blockIdx.y:  mv[set][i] = prod_c (beta * sigma_c[i] + gamma + v_c[i])
__global__ void __launch_bounds__(256) k_perm_den(const PermCols* Cs, Fr beta, Fr gamma, uint4* mv, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PermCols& C = Cs[blockIdx.y];
  Fr m = Fr::one();
  for (uint32_t c = 0; c < C.ncols; c++)
    m = m * (beta * Fr::load(C.sigma[c] + 2 * i) + gamma + Fr::load(C.values[c] + 2 * i));
  m.store(mv + 2 * ((size_t)blockIdx.y * n + i));
}
// mv[set][i] *= prod_c (delta^(c0+c) * omega^i * beta + gamma + v_c[i]);  dbeta[c] = delta^(c0+c) * beta
__global__ void __launch_bounds__(256) k_perm_num(const PermCols* Cs, const PermNum* Ds, Fr gamma, const uint4* tw, uint32_t tw_shift,
                                                  size_t half_ext, uint4* mv, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PermCols& C = Cs[blockIdx.y];
  const PermNum& D = Ds[blockIdx.y];
  Fr w = omega_pow(tw, i, tw_shift, half_ext);
  uint4* p = mv + 2 * ((size_t)blockIdx.y * n + i);
  Fr m = Fr::load(p);
  for (uint32_t c = 0; c < C.ncols; c++) m = m * (D.dbeta[c] * w + gamma + Fr::load(C.values[c] + 2 * i));
  m.store(p);
}
// seeds[s] = last_z entering set s: seeds[0] = 1, seeds[s+1] = seeds[s] * zloc[s][usable]   (one thread)
__global__ void k_perm_seeds(const uint4* zloc, size_t n, size_t usable, uint32_t n_sets, uint4* seeds) {
  if (threadIdx.x || blockIdx.x) return;
  Fr s = Fr::one();
  for (uint32_t i = 0; i < n_sets; i++) {
    s.store(seeds + 2 * i);
    s = s * Fr::load(zloc + 2 * ((size_t)i * n + usable));
  }
}
// a[seg][i] *= seeds[seg]
__global__ void __launch_bounds__(256) k_scale_seg(uint4* a, const uint4* seeds, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint4* p = a + 2 * ((size_t)blockIdx.y * n + i);
  (Fr::load(p) * Fr::load(seeds + 2 * blockIdx.y)).store(p);
}
// lookup product denominators / numerators; one lookup per blockIdx.y, planes of L columns each
__global__ void __launch_bounds__(256) k_lookup_den(const uint4* pin, const uint4* ptab, Fr beta, Fr gamma, uint4* lp, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t o = 2 * ((size_t)blockIdx.y * n + i);
  ((beta + Fr::load(pin + o)) * (gamma + Fr::load(ptab + o))).store(lp + o);
}
__global__ void __launch_bounds__(256) k_lookup_num(const uint4* cin, const uint4* ctab, Fr beta, Fr gamma, uint4* lp, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t o = 2 * ((size_t)blockIdx.y * n + i);
  (Fr::load(lp + o) * (Fr::load(cin + o) + beta) * (Fr::load(ctab + o) + gamma)).store(lp + o);
}

// In-place batch inversion, zeros stay zero (ff batch_invert semantics).  Each thread owns kInvChunk consecutive
// elements: prefix products into `scratch`, one Fermat inversion (~380 multiplications, so the chunk is long enough to
// amortise it: 3 + 380/128 multiplications per element), back substitution.
constexpr int kInvChunk = 128;
__global__ void __launch_bounds__(128) k_batch_invert(uint4* v, uint4* scratch, size_t n) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * kInvChunk, hi = lo + kInvChunk < n ? lo + kInvChunk : n;
  if (lo >= n) return;
  Fr acc = Fr::one();
  for (size_t i = lo; i < hi; i++) {
    acc.store_raw(scratch + 2 * i);
    Fr x = Fr::load(v + 2 * i);
    if (!x.is_zero()) acc = acc * x;
  }
  acc = Fr::inv(acc);
  for (size_t i = hi; i-- > lo;) {
    Fr x = Fr::load(v + 2 * i);
    if (x.is_zero()) continue;
    (acc * Fr::load(scratch + 2 * i)).store(v + 2 * i);
    acc = acc * x;
  }
}

// table[i] = z^i, i < n: thread t starts at z^(t*kPowChunk) by square-and-multiply
constexpr int kPowChunk = 64;      // ~20 squarings / multiplications to reach the chunk's first power, then one multiplication per entry
__global__ void __launch_bounds__(128) k_pow_table(Fr z, uint4* table, size_t n) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * kPowChunk, hi = lo + kPowChunk < n ? lo + kPowChunk : n;
  if (lo >= n) return;
  Fr cur = Fr::pow_u64(z, (uint64_t)lo);
  for (size_t i = lo; i < hi; i++) { cur.store(table + 2 * i); cur = cur * z; }
}

// eval_polynomial for a list of polynomials at one point: partial[p][chunk] = sum_{i in chunk} poly_p[i] * table[i]
constexpr int kEvalThreads = 256;
__global__ void __launch_bounds__(kEvalThreads) k_eval_partial(const uint4* const* polys, const uint4* table, size_t n,
                                                               uint4* partial, uint32_t nchunks) {
  __shared__ Fr sh[kEvalThreads];
  const uint4* poly = polys[blockIdx.y];
  size_t per = (n + nchunks - 1) / nchunks;
  size_t lo = (size_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  Fr acc = Fr::zero();
  for (size_t i = lo + threadIdx.x; i < hi; i += kEvalThreads) acc = acc + Fr::load(poly + 2 * i) * Fr::load(table + 2 * i);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = kEvalThreads / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) sh[0].store(partial + 2 * ((size_t)blockIdx.y * nchunks + blockIdx.x));
}
__global__ void k_eval_final(const uint4* partial, uint32_t nchunks, uint4* out, uint32_t npolys) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npolys) return;
  Fr acc = Fr::zero();
  for (uint32_t c = 0; c < nchunks; c++) acc = acc + Fr::load(partial + 2 * ((size_t)p * nchunks + c));
  acc.store(out + 2 * p);
}

// cost[col] += number of non-zero 16-bit windows of the canonical values of column `col` (a proxy for the bucket additions a
// commitment of that column performs: the group assigns columns to GPUs by it)
__global__ void __launch_bounds__(256) k_column_cost(const uint4* cols, size_t n, unsigned int* cost) {
  const size_t col = blockIdx.y;
  unsigned int mine = 0;
  // every 8th row is enough for a RELATIVE cost (the columns of one batch hold the same kind of value in every row)
  for (size_t i = 8 * (blockIdx.x * (size_t)blockDim.x + threadIdx.x); i < n; i += 8 * (size_t)gridDim.x * blockDim.x) {
    const Fr v = Fr::load(cols + 2 * (col * n + i)).from_mont();
#pragma unroll
    for (int l = 0; l < 8; l++) mine += ((v.l[l] & 0xFFFFu) != 0) + ((v.l[l] >> 16) != 0);
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&cost[col], mine);
}
// out[map[j]] = src[j] for XYZZ records (128 bytes)
__global__ void k_scatter_points(const uint4* src, const unsigned int* map, uint4* out, unsigned int count) {
  unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * 8) return;
  const unsigned int dst = map[t >> 3];
  if (dst != 0xFFFFFFFFu) out[8 * (size_t)dst + (t & 7)] = src[t];       // padding slots of ranks with fewer columns
}

// out[i] = sum_j scalars[j] * polys[j][i]   (+ acc_in[i] when given)
__global__ void __launch_bounds__(256) k_lincomb(const uint4* const* polys, const uint4* scalars, uint32_t m, uint4* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // two independent accumulators and loads issued four terms ahead: the terms are independent products, only the sums chain
  // (field addition is exact, so the association order does not change the canonical result)
  Fr acc0 = Fr::zero(), acc1 = Fr::zero();
  uint32_t j = 0;
  for (; j + 4 <= m; j += 4) {
    const Fr p0 = Fr::load(polys[j] + 2 * i), p1 = Fr::load(polys[j + 1] + 2 * i), p2 = Fr::load(polys[j + 2] + 2 * i), p3 = Fr::load(polys[j + 3] + 2 * i);
    acc0 = acc0 + Fr::load(scalars + 2 * j) * p0;
    acc1 = acc1 + Fr::load(scalars + 2 * (j + 1)) * p1;
    acc0 = acc0 + Fr::load(scalars + 2 * (j + 2)) * p2;
    acc1 = acc1 + Fr::load(scalars + 2 * (j + 3)) * p3;
  }
  for (; j < m; j++) acc0 = acc0 + Fr::load(scalars + 2 * j) * Fr::load(polys[j] + 2 * i);
  (acc0 + acc1).store(out + 2 * i);
}
// a[i] -= low[i] for i < m (m <= 8): subtracts a low-degree polynomial in place
struct LowPoly { Fr c[8]; uint32_t m; };
__global__ void k_sub_low(uint4* a, LowPoly L) {
  uint32_t i = threadIdx.x;
  if (i < L.m) (Fr::load(a + 2 * i) - L.c[i]).store(a + 2 * i);
}
// out[i] = a[i] * b[i]
__global__ void __launch_bounds__(256) k_mul_vec(const uint4* a, const uint4* b, uint4* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  (Fr::load(a + 2 * i) * Fr::load(b + 2 * i)).store(out + 2 * i);
}
// kate_division epilogue: q[i-1] = S[i] * zinv^i for 1 <= i < n, q[n-1] = 0   (S = suffix sums of a[j] z^j)
__global__ void __launch_bounds__(256) k_kate_finish(const uint4* S, const uint4* zinv_table, uint4* q, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i + 1 < n) (Fr::load(S + 2 * (i + 1)) * Fr::load(zinv_table + 2 * (i + 1))).store(q + 2 * i);
  else Fr::zero().store(q + 2 * i);
}
// a[i] = a[i] * s
__global__ void __launch_bounds__(256) k_scale(uint4* a, Fr s, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  (Fr::load(a + 2 * i) * s).store(a + 2 * i);
}
// sigma[c][j] = delta^(c') * omega^(j') for mapping[c][j] = (c', j')   (permutation keygen)
__global__ void __launch_bounds__(256) k_sigma(const uint2* mapping, const uint4* delta_pows, const uint4* tw, uint32_t tw_shift,
                                               size_t half_ext, uint4* out, size_t total) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint2 m = mapping[i];
  (Fr::load(delta_pows + 2 * (size_t)m.x) * omega_pow(tw, m.y, tw_shift, half_ext)).store(out + 2 * i);
}
// Lagrange indicator columns: out[i] = (lo <= i < hi) ? 1 : 0
__global__ void __launch_bounds__(256) k_indicator(uint4* out, size_t n, size_t lo, size_t hi) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  ((i >= lo && i < hi) ? Fr::one() : Fr::zero()).store(out + 2 * i);
}

// ---- lookup: expressions over the Lagrange rows ---------------------------------------------------------------
// One graph of the program (the theta-compressed input or table expressions of a lookup) on rows [0, n).
template <int NI>
__global__ void __launch_bounds__(128) k_eval_rows(EvalProgram P, EvalScalars S, uint32_t graph, const uint32_t* code,
                                                   const uint4* const* cols, const uint4* consts, const uint4* challenges,
                                                   uint4* out, size_t n) {
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  EvalState<NI> st;
  Fr v = graph_eval<NI>(P, P.lookups[graph], code, cols, consts, challenges, S, st, Fr::zero(), idx);
  v.store(out + 2 * idx);
}

// ---- lookup: permute_expression_pair ----------------------------------------------------------------------------
// All sorts of one proof (2 per lookup argument: compressed inputs and compressed tables) run in the same launches:
// blockIdx.y = sort index.  Values / keys / sorted keys of sort s start at element s * stride; the index arrays are
// packed with stride n (the number of sorted elements), so that the concatenation of all sorts is ONE array whose
// radix offsets come from one global scan of hist[sort][digit][block].
// keys = canonical (non-Montgomery) values so that limb order = numeric order (Fr: Ord)
// The sorts cover the lookups [l0, l0 + Lloc) of Ltot: sort s < Lloc reads column l0 + s of the input plane, sort
// Lloc + s column l0 + s of the table plane (the planes are adjacent, Ltot columns each); keys / idx are dense in s.
__global__ void __launch_bounds__(256) k_canon_keys(const uint4* v, uint4* keys, uint32_t* idx, size_t n, size_t stride, uint32_t Lloc,
                                                    uint32_t l0, uint32_t Ltot) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = blockIdx.y;
  const size_t col = s < Lloc ? (size_t)l0 + s : (size_t)Ltot + l0 + (s - Lloc);
  Fr::load(v + 2 * (col * stride + i)).from_mont().store_raw(keys + 2 * ((size_t)s * stride + i));
  idx[(size_t)s * n + i] = (uint32_t)i;
}
// mask[sort][l] |= keys[i].l[l] ^ keys[0].l[l]: which bits differ anywhere in the array
__global__ void __launch_bounds__(256) k_key_diff_mask(const uint32_t* keys_all, size_t n, size_t stride, uint32_t* mask_all) {
  __shared__ uint32_t sh[8];
  const uint32_t* keys = keys_all + 8 * (size_t)blockIdx.y * stride;
  uint32_t* mask = mask_all + 8 * blockIdx.y;
  if (threadIdx.x < 8) sh[threadIdx.x] = 0;
  __syncthreads();
  uint32_t m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
#pragma unroll
    for (int l = 0; l < 8; l++) m[l] |= keys[8 * i + l] ^ keys[l];
#pragma unroll
  for (int l = 0; l < 8; l++) {
    uint32_t x = m[l];
    for (int off = 16; off > 0; off >>= 1) x |= __shfl_xor_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0 && x) atomicOr(&sh[l], x);
  }
  __syncthreads();
  if (threadIdx.x < 8 && sh[threadIdx.x]) atomicOr(&mask[threadIdx.x], sh[threadIdx.x]);
}

constexpr int kRadixThreads = 256;
constexpr int kRadixRows = 8;
constexpr int kRadixTile = kRadixThreads * kRadixRows;
constexpr uint32_t kNullDigit = 0xFFu;      // a padding pass: every key has digit 0 (stable identity permutation)
__device__ __forceinline__ uint32_t key_digit(const uint32_t* keys, uint32_t id, uint32_t d) {
  if (d == kNullDigit) return 0;
  return (keys[8 * (size_t)id + (d >> 2)] >> (8 * (d & 3))) & 0xFFu;
}
// hist[(sort * 256 + digit) * nblocks + block];  digits[sort] = digit index of this pass for that sort
__global__ void __launch_bounds__(kRadixThreads) k_radix_hist(const uint32_t* keys_all, const uint32_t* idx_all, size_t n, size_t stride,
                                                              const uint32_t* digits, uint32_t* hist) {
  __shared__ uint32_t sh[256];
  const uint32_t* keys = keys_all + 8 * (size_t)blockIdx.y * stride;
  const uint32_t* idx = idx_all + (size_t)blockIdx.y * n;
  const uint32_t d = digits[blockIdx.y];
  sh[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * kRadixTile;
  for (int r = 0; r < kRadixRows; r++) {
    size_t i = base + (size_t)r * kRadixThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[key_digit(keys, idx[i], d)], 1u);
  }
  __syncthreads();
  hist[((size_t)blockIdx.y * 256 + threadIdx.x) * gridDim.x + blockIdx.x] = sh[threadIdx.x];
}
// stable scatter; `offs` = exclusive scan of the whole hist array (offsets into the concatenated index array)
__global__ void __launch_bounds__(kRadixThreads) k_radix_scatter(const uint32_t* keys_all, const uint32_t* idx_in_all, uint32_t* idx_out_all,
                                                                 size_t n, size_t stride, const uint32_t* digits, const uint32_t* offs) {
  __shared__ uint32_t cnt[kRadixThreads / 32][256];
  const uint32_t* keys = keys_all + 8 * (size_t)blockIdx.y * stride;
  const uint32_t* idx_in = idx_in_all + (size_t)blockIdx.y * n;
  const uint32_t d = digits[blockIdx.y];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t running = offs[((size_t)blockIdx.y * 256 + threadIdx.x) * gridDim.x + blockIdx.x];     // thread t owns digit t
  size_t base = (size_t)blockIdx.x * kRadixTile;
  for (int r = 0; r < kRadixRows; r++) {
    for (int q = 0; q < kRadixThreads / 32; q++) cnt[q][threadIdx.x] = 0;
    __syncthreads();
    size_t i = base + (size_t)r * kRadixThreads + threadIdx.x;
    bool valid = i < n;
    uint32_t id = valid ? idx_in[i] : 0;
    uint32_t dg = valid ? key_digit(keys, id, d) : 256u;
    uint32_t peers = __match_any_sync(0xffffffffu, dg);
    uint32_t rank = __popc(peers & ((1u << lane) - 1));
    if (valid && rank == 0) cnt[w][dg] = __popc(peers);
    __syncthreads();
    {   // exclusive prefix over the warps for digit = threadIdx.x, seeded with the running global offset
      uint32_t acc = running;
      for (int q = 0; q < kRadixThreads / 32; q++) { uint32_t c = cnt[q][threadIdx.x]; cnt[q][threadIdx.x] = acc; acc += c; }
      running = acc;
    }
    __syncthreads();
    if (valid) idx_out_all[cnt[w][dg] + rank] = id;      // global offset: already includes sort * n
    __syncthreads();
  }
}

__device__ __forceinline__ int key_cmp(const uint32_t* a, const uint32_t* b) {
  for (int l = 7; l >= 0; l--) {
    if (a[l] != b[l]) return a[l] < b[l] ? -1 : 1;
  }
  return 0;
}
// sk[j] = keys[idx[j]]
__global__ void __launch_bounds__(256) k_gather_keys(const uint4* keys_all, const uint32_t* idx_all, uint4* sk_all, size_t n, size_t stride) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint4* keys = keys_all + 2 * (size_t)blockIdx.y * stride;
  uint4* sk = sk_all + 2 * (size_t)blockIdx.y * stride;
  size_t s = idx_all[(size_t)blockIdx.y * n + j];
  sk[2 * j] = keys[2 * s]; sk[2 * j + 1] = keys[2 * s + 1];
}
// flag[0] |= 1 when some adjacent pair is out of order under the FULL comparison (the top-digit passes did not
// produce a total order)
__global__ void __launch_bounds__(256) k_check_sorted(const uint32_t* sk_all, size_t n, size_t stride, uint32_t* flag) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j + 1 >= n) return;
  const uint32_t* sk = sk_all + 8 * (size_t)blockIdx.y * stride;
  if (key_cmp(sk + 8 * j, sk + 8 * (j + 1)) > 0) atomicOr(flag, 1u);
}
// lower_bound of key in sorted array
__device__ __forceinline__ size_t key_lower_bound(const uint32_t* sorted, size_t n, const uint32_t* key) {
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) >> 1;
    if (key_cmp(sorted + 8 * mid, key) < 0) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// One lookup per blockIdx.y; sorts [0, L) are the inputs, [L, 2L) the tables.  rep[j] = 1 when row j repeats the
// previous input value; left[j] = 1 when table entry j is NOT consumed by a first occurrence.  flag |= 2 when an
// input value is not in the table.
__global__ void __launch_bounds__(256) k_permute_flags(const uint32_t* sk_all, size_t n, size_t stride, uint32_t L, uint32_t* rep_all,
                                                       uint32_t* left_all, uint32_t* flag) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t* sa = sk_all + 8 * (size_t)blockIdx.y * stride;
  const uint32_t* st = sk_all + 8 * (size_t)(L + blockIdx.y) * stride;
  uint32_t* rep = rep_all + (size_t)blockIdx.y * n;
  uint32_t* left = left_all + (size_t)blockIdx.y * n;
  bool afirst = j == 0 || key_cmp(sa + 8 * j, sa + 8 * (j - 1)) != 0;
  rep[j] = afirst ? 0u : 1u;
  if (afirst) {
    size_t p = key_lower_bound(st, n, sa + 8 * j);
    if (p >= n || key_cmp(st + 8 * p, sa + 8 * j) != 0) atomicOr(flag, 2u);
  }
  bool tfirst = j == 0 || key_cmp(st + 8 * j, st + 8 * (j - 1)) != 0;
  bool used = false;
  if (tfirst) {
    size_t p = key_lower_bound(sa, n, st + 8 * j);
    used = p < n && key_cmp(sa + 8 * p, st + 8 * j) == 0;
  }
  left[j] = (tfirst && used) ? 0u : 1u;
}
// totals[l] = number of repeated rows of lookup l (from the exclusive ranks + last flags); flag |= 4 on a mismatch
__global__ void k_permute_totals(const uint32_t* rep, const uint32_t* rrank, const uint32_t* left, const uint32_t* lrank, size_t n,
                                 uint32_t L, uint32_t* totals, uint32_t* flag) {
  uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  size_t last = (size_t)l * n + n - 1;
  uint32_t a = rrank[last] + rep[last], b = lrank[last] + left[last];
  totals[l] = a;
  if (a != b) atomicOr(flag, 4u);
}
// rpos[rrank[j]] = j for repeated rows
__global__ void __launch_bounds__(256) k_permute_rpos(const uint32_t* rep_all, const uint32_t* rrank_all, uint32_t* rpos_all, size_t n) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  size_t o = (size_t)blockIdx.y * n;
  if (rep_all[o + j]) rpos_all[o + rrank_all[o + j]] = (uint32_t)j;
}
// A'[j] = input[ia[j]];  S'[j] = A'[j] on first occurrences; leftover table entry with rank q goes to the
// (m-1-q)-th repeated row (repeated_input_rows.pop()).  Values are gathered in Montgomery form from the sources.
// input / table / pin / ptab are planes of L columns with `stride` elements each.
__global__ void __launch_bounds__(256) k_permute_write(const uint4* input_all, const uint4* table_all, const uint32_t* idx_all,
                                                       const uint32_t* rep_all, const uint32_t* left_all, const uint32_t* lrank_all,
                                                       const uint32_t* rpos_all, const uint32_t* totals, uint32_t L, uint4* pin_all,
                                                       uint4* ptab_all, size_t n, size_t stride, uint32_t l0) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t l = blockIdx.y;                 // local lookup; its columns are l0 + l of each plane
  const size_t o = (size_t)l * n, so = 2 * (size_t)(l0 + l) * stride;
  const uint4* input = input_all + so;
  const uint4* table = table_all + so;
  uint4* pin = pin_all + so;
  uint4* ptab = ptab_all + so;
  size_t s = idx_all[o + j];
  uint4 a0 = input[2 * s], a1 = input[2 * s + 1];
  pin[2 * j] = a0; pin[2 * j + 1] = a1;
  if (!rep_all[o + j]) { ptab[2 * j] = a0; ptab[2 * j + 1] = a1; }
  if (left_all[o + j]) {
    uint32_t m = totals[l];                      // number of repeated rows
    uint32_t q = lrank_all[o + j];
    if (q < m) {
      size_t row = rpos_all[o + m - 1 - q], t = idx_all[(size_t)(L + l) * n + j];
      ptab[2 * row] = table[2 * t]; ptab[2 * row + 1] = table[2 * t + 1];
    }
  }
}

}  // namespace zk
