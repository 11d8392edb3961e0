// prover.cu -- keygen_pk and create_proof with every n-sized loop on the device.
//
// Restates the ORDER of halo2_proofs plonk/prover.rs `create_proof`, plonk/keygen.rs `keygen_vk`/`keygen_pk`,
// plonk/{lookup,permutation,vanishing}/prover.rs and poly/kzg/multiopen/shplonk/prover.rs ([UPSTREAM] @ c254c75f,
// Cargo.lock:1314-1316; called at circuits/src/sgx_dcap_verifier.rs:803, :807, :814-822): RNG draw order
// (SURVEY.md A.7), transcript order (3.2, A.8), query order and SHPLONK set construction (A.5).  Field and group
// results are canonical, so HOW each vector is computed is free: commitments are the Pippenger MSM of msm.cu,
// transforms the NTT of ntt.cu, h(X) the fused kernel of evalh.cu, and the loops in between the kernels of
// prover_kernels.cuh.  The host only hashes (Blake2b transcript) and does O(#columns) scalar arithmetic.
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <set>

#include "common.cuh"
#include "evalh_compile.hpp"
#include "prover_kernels.cuh"
#include "transcript.hpp"

using namespace zk;

// entry points of the other translation units
extern "C" int zk_lagrange_to_coeff_dev(zk_domain* d, zk_fr* d_a, size_t count);
extern "C" int zk_extended_to_coeff_dev(zk_domain* d, zk_fr* d_a_ext, zk_fr* d_out);
extern "C" int zk_divide_by_vanishing_poly_dev(zk_domain* d, zk_fr* d_a_ext);
extern "C" int zk_msm_batch_dev(zk_ctx* ctx, const zk_fr* d_scalars, int which, size_t n, size_t count, zk_g1* d_out);
extern "C" int zk_coeff_to_extended_dev(zk_domain* d, const zk_fr* d_a, zk_fr* d_out_ext, size_t count);
int zk_lagrange_to_coeff_oop(zk_domain* d, const zk_fr* d_src, zk_fr* d_dst, size_t count);
int zk_coeff_to_extended_strided(zk_domain* d, const zk_fr* d_a, zk_fr* d_out_ext, size_t count, size_t out_stride);
char* zk_pk_witness_slot(zk_pk* pk, size_t slot);
int zk_evaluate_h_launch(zk_pk* pk, const zk_fr* challenges, const zk_fr* y, const zk_fr* beta, const zk_fr* gamma,
                         const zk_fr* theta, zk_fr* d_out_ext, size_t row0, size_t row1);
struct zk_group;
int zk_group_allgather(zk_group* g, zk_ctx* on, void* d_buf, size_t bytes_per_rank);
bool zk_group_p2p_begin(zk_group* g, zk_ctx* on, size_t bytes_per_rank, zk_p2p_push* out);
int zk_group_p2p_finish(zk_group* g, zk_ctx* on, void* d_buf, size_t bytes_per_rank);
int zk_group_p2p_push(zk_group* g, zk_ctx* on, const zk_p2p_push& pp, const void* my_segment, size_t bytes);
int zk_group_p2p_error(zk_group* g, zk_ctx* on);
int zk_group_register(zk_group* g, void* base, size_t bytes);
size_t zk_pk_witness_bytes(zk_pk* pk);
int zk_group_exchange_rows(zk_group* g, zk_ctx* on, char* const* cols, const uint32_t* owners, size_t count, size_t height, size_t halo);
extern "C" uint32_t zk_group_rank(const zk_group* g);
extern "C" uint32_t zk_group_world(const zk_group* g);
extern "C" void zk_group_shard_plan(uint32_t rank, uint32_t world, size_t count, size_t n, int* mode, size_t* per, size_t* lo, size_t* hi);
extern "C" void zk_group_balance_plan(const uint32_t* cost, size_t count, uint32_t world, uint32_t* owner);
int zk_msm_range_batch_dev(zk_ctx* ctx, const zk_fr* d_scalars, int which, size_t sc_stride, size_t base_off, size_t n_sub,
                           size_t count, zk_g1* d_out);
int zk_g1_sum_strided_dev(zk_ctx* ctx, const zk_g1* d_points, size_t parts, size_t count, zk_g1* d_out);
const void* zk_domain_twiddles(const zk_domain* d);
size_t zk_domain_ext_n(const zk_domain* d);

struct zk_pkey {
  zk_ctx* ctx = nullptr;
  zk_pk* pk = nullptr;             // evaluate_h side (cosets, compiled program); owns its own domain
  zk_domain* dom = nullptr;
  int transcript_kind = 0;        // zk_transcript_kind
  zk_group* group = nullptr;      // not owned: the GPUs that share this proof's commitments and h(X) rows
  // second stream: every lagrange_to_coeff / coeff_to_extended of the witness polynomials runs here, overlapping the
  // commitments and the lookup / permutation kernels of the main stream (they only meet again at evaluate_h)
  zk_ctx* ctxB = nullptr;
  zk_domain* domB = nullptr;
  cudaEvent_t ev_in[4] = {nullptr, nullptr, nullptr, nullptr}, ev_done = nullptr;
  cudaEvent_t ev_copy[4] = {nullptr, nullptr, nullptr, nullptr};   // groups of advice columns uploaded on the copy stream
  char *d_advice_polys = nullptr, *d_lookup_polys = nullptr, *d_perm_z_polys = nullptr;
  uint32_t k = 0, degree = 0, bf = 0, n_fixed = 0, n_advice = 0, n_instance = 0, n_challenges = 0, n_lookups = 0;
  uint32_t n_perm = 0, n_sets = 0, chunk = 0;
  size_t n = 0, ext_n = 0;
  std::vector<zk_column> perm_columns;
  std::vector<zk_query> advice_queries, fixed_queries;
  Fr transcript_repr;
  Fr omega, omega_inv;
  uint32_t max_rotation = 0;       // largest |rotation| any expression, the permutation or the lookup arguments use
  // resident key material (Lagrange values and coefficient forms; the cosets live in `pk`)
  char *d_fixed_values = nullptr, *d_fixed_polys = nullptr, *d_sigma_values = nullptr, *d_sigma_polys = nullptr;
  std::vector<zk_g1_affine> fixed_commitments, perm_commitments;
  // lookup compression program: graphs 2l (inputs) and 2l+1 (tables) of lookup l
  CompiledEval CL;
  uint32_t* d_lcode = nullptr;
  void* d_lconsts = nullptr;
  const void** d_lcols = nullptr;
  void* d_challenges = nullptr;
  // per-proof work area, allocated once
  char *d_advice = nullptr, *d_instance = nullptr, *d_instance_polys = nullptr, *d_lookup = nullptr, *d_perm_z = nullptr;
  char *d_random = nullptr, *d_hext = nullptr, *d_hpieces = nullptr, *d_hpoly = nullptr, *d_hx = nullptr, *d_lx = nullptr;
  char *d_w0 = nullptr, *d_w1 = nullptr, *d_w2 = nullptr;      // n-element scratch vectors
  char *d_tables = nullptr;                                    // pow tables: slot s at d_tables + s*n*32
  uint32_t n_table_slots = 0;
  std::vector<char*> d_setacc;                                 // sum_j y^j P_ij per rotation set
  char *d_scan_totals = nullptr, *d_seed = nullptr;
  char *d_keys = nullptr, *d_sk = nullptr;                      // canonical keys / sorted keys of the 2L sorts
  uint32_t *d_idx[2] = {nullptr, nullptr}, *d_hist = nullptr, *d_u32[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t *d_masks = nullptr, *d_digits = nullptr, *d_totals = nullptr;
  char *d_permcols = nullptr, *d_permnum = nullptr, *d_seeds = nullptr;
  size_t scan_cap = 0;
  uint64_t sort_fallbacks = 0;
  uint32_t* d_flags = nullptr;                                  // [0] failure bits, [1..2] totals, [8..15] diff mask
  char* d_xchg = nullptr;                                       // group exchange area for commitment points
  char* d_balance = nullptr; size_t balance_cap = 0;            // group: gathered columns of a cost-balanced commitment
  unsigned int* d_cost = nullptr;                               // group: per-column costs / scatter maps (2 x 256 words)
  char* d_stage_p = nullptr;                                    // group staging of coefficient forms (allocated by set_group)
  size_t stage_cols = 0;
  size_t halo_rows = 0;                                          // extended rows a slice of h(X) reads beyond its own range (0: exchange whole columns)
  char *d_points = nullptr, *d_ptrs = nullptr, *d_scalars = nullptr, *d_partial = nullptr, *d_evals = nullptr;
  uint32_t max_list = 0, eval_chunks = 1;
  std::vector<void*> allocs;
};

namespace {

Fr fr_canon_words(const uint32_t c[8]) { Fr v; memcpy(v.l, c, 32); return v.to_mont(); }
Fr fr_delta() {
  const uint32_t c[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
  return fr_canon_words(c);
}
Fr fr_small(uint64_t x) { Fr v = Fr::zero(); v.l[0] = (uint32_t)x; v.l[1] = (uint32_t)(x >> 32); return v.to_mont(); }
bool fr_less(const Fr& a, const Fr& b) {        // Fr: Ord = numeric order of the canonical value
  Fr x = a.from_mont(), y = b.from_mont();
  for (int i = 7; i >= 0; i--) if (x.l[i] != y.l[i]) return x.l[i] < y.l[i];
  return false;
}
bool fr_same(const Fr& a, const Fr& b) { return (a - b).is_zero(); }

template <class T>
int dev_alloc(zk_pkey* K, T** p, size_t bytes) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes ? bytes : 32);
  if (e != cudaSuccess) {
    K->ctx->err = std::string("zk_keygen: cudaMalloc: ") + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? ZK_ERR_OOM : ZK_ERR_CUDA;
  }
  K->allocs.push_back(q);
  *p = (T*)q;
  return ZK_OK;
}

unsigned blocks_for(size_t n, unsigned per) { return zk_div_up(n, per); }

// ---- scans ----------------------------------------------------------------------------------------------------
template <class M>
int run_scan(zk_pkey* K, const char* name, const void* in, void* out, size_t n, int reverse, int exclusive, const void* d_seeds,
             unsigned nseg = 1, size_t stride_in = 0, size_t stride_out = 0) {
  zk_ctx* ctx = K->ctx;
  unsigned tiles = blocks_for(n, kScanTile);
  if ((size_t)tiles * nseg > K->scan_cap) ZK_FAIL(ctx, ZK_ERR_STATE, "prover: scan of %zu x %u elements exceeds the scratch", n, nseg);
  dim3 grid(tiles, nseg);
  ZK_LAUNCH(ctx, name, k_scan_tiles<M>, grid, kScanThreads, 0, in, n, stride_in, reverse, (void*)K->d_scan_totals);
  ZK_LAUNCH(ctx, name, k_scan_totals<M>, nseg, kScanThreads, 0, (void*)K->d_scan_totals, (size_t)tiles, d_seeds);
  ZK_LAUNCH(ctx, name, k_scan_apply<M>, grid, kScanThreads, 0, in, out, n, stride_in, stride_out, reverse, exclusive,
            (const void*)K->d_scan_totals);
  return ZK_OK;
}

int batch_invert(zk_pkey* K, char* v, char* scratch, size_t n) {
  zk_ctx* ctx = K->ctx;
  ZK_LAUNCH(ctx, "batch_invert", k_batch_invert, blocks_for(blocks_for(n, kInvChunk), 128), 128, 0, (uint4*)v, (uint4*)scratch, n);
  return ZK_OK;
}

// ---- pow tables --------------------------------------------------------------------------------------------------
char* table_slot(zk_pkey* K, uint32_t slot) { return K->d_tables + (size_t)slot * K->n * 32; }
int build_pow_table(zk_pkey* K, const Fr& z, uint32_t slot) {
  zk_ctx* ctx = K->ctx;
  ZK_LAUNCH(ctx, "pow_table", k_pow_table, blocks_for(blocks_for(K->n, kPowChunk), 128), 128, 0, z, (uint4*)table_slot(K, slot), K->n);
  return ZK_OK;
}

// ---- lists of device pointers / scalars ------------------------------------------------------------------------------
int upload_list(zk_pkey* K, const std::vector<const void*>& ptrs, const std::vector<Fr>* scalars) {
  zk_ctx* ctx = K->ctx;
  if (ptrs.size() > K->max_list) ZK_FAIL(ctx, ZK_ERR_ARG, "prover: pointer list of %zu exceeds %u", ptrs.size(), K->max_list);
  ZK_CUDA(ctx, cudaMemcpyAsync(K->d_ptrs, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  if (scalars) {
    std::vector<zk_fr> s(scalars->size());
    for (size_t i = 0; i < s.size(); i++) f_to_abi((*scalars)[i], &s[i]);
    ZK_CUDA(ctx, cudaMemcpyAsync(K->d_scalars, s.data(), s.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  // pageable sources are staged before the call returns
  return ZK_OK;
}
// out = sum_j scalars[j] * polys[j]
int lincomb(zk_pkey* K, const std::vector<const void*>& polys, const std::vector<Fr>& scalars, char* out) {
  zk_ctx* ctx = K->ctx;
  ZK_TRY(upload_list(K, polys, &scalars));
  ZK_LAUNCH(ctx, "poly_lincomb", k_lincomb, blocks_for(K->n, 256), 256, 0, (const uint4* const*)K->d_ptrs, (const uint4*)K->d_scalars,
            (uint32_t)polys.size(), (uint4*)out, K->n);
  return ZK_OK;
}
// evaluations of `polys` at the point whose pow table sits in `slot`, left in device memory at d_dst (canonical-range Fr)
int eval_list_to(zk_pkey* K, const std::vector<const void*>& polys, uint32_t slot, char* d_dst) {
  zk_ctx* ctx = K->ctx;
  if (polys.empty()) return ZK_OK;
  ZK_TRY(upload_list(K, polys, nullptr));
  dim3 grid(K->eval_chunks, (unsigned)polys.size());
  ZK_LAUNCH(ctx, "eval_polynomial", k_eval_partial, grid, kEvalThreads, 0, (const uint4* const*)K->d_ptrs, (const uint4*)table_slot(K, slot),
            K->n, (uint4*)K->d_partial, K->eval_chunks);
  ZK_LAUNCH(ctx, "eval_polynomial", k_eval_final, blocks_for(polys.size(), 64), 64, 0, (const uint4*)K->d_partial, K->eval_chunks,
            (uint4*)d_dst, (uint32_t)polys.size());
  return ZK_OK;
}
// evaluations of `polys` at the point whose pow table sits in `slot`; results appended to `out` (host, after sync)
int eval_list(zk_pkey* K, const std::vector<const void*>& polys, uint32_t slot, std::vector<Fr>* out) {
  zk_ctx* ctx = K->ctx;
  if (polys.empty()) return ZK_OK;
  ZK_TRY(eval_list_to(K, polys, slot, K->d_evals));
  std::vector<zk_fr> h(polys.size());
  ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), K->d_evals, polys.size() * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& e : h) out->push_back(f_from_abi<Fr>(&e));
  return ZK_OK;
}
// a <- a / (X - z) with the pow tables of z in slot_z and of 1/z in slot_zinv (kate_division; remainder dropped)
int kate_divide(zk_pkey* K, char* a, uint32_t slot_z, uint32_t slot_zinv) {
  zk_ctx* ctx = K->ctx;
  size_t n = K->n;
  ZK_LAUNCH(ctx, "kate_division", k_mul_vec, blocks_for(n, 256), 256, 0, (const uint4*)a, (const uint4*)table_slot(K, slot_z), (uint4*)K->d_w1, n);
  ZK_TRY(run_scan<FrAddMonoid>(K, "kate_division", K->d_w1, K->d_w2, n, 1, 0, nullptr));
  ZK_LAUNCH(ctx, "kate_division", k_kate_finish, blocks_for(n, 256), 256, 0, (const uint4*)K->d_w2, (const uint4*)table_slot(K, slot_zinv), (uint4*)a, n);
  return ZK_OK;
}

// commitments: `count` contiguous polynomials -> affine points on the host (normalised Jacobian from the MSM)
struct HostPoint { Fq x, y; bool identity; };
// MSM results inside the prover are XYZZ records (ctx->msm_raw): x = X / ZZ, y = Y / ZZZ, identity when ZZ = 0
struct PtRaw { zk_fq x, y, zz, zzz; };
struct RawGuard {
  zk_ctx* c; bool old;
  explicit RawGuard(zk_ctx* ctx) : c(ctx), old(ctx->msm_raw) { ctx->msm_raw = true; }
  ~RawGuard() { c->msm_raw = old; }
};
// host normalisation of a batch with ONE field inversion (Montgomery's trick)
void normalize_points(const PtRaw* raw, size_t count, std::vector<HostPoint>* out) {
  std::vector<Fq> den(count), pre(count);
  Fq acc = Fq::one();
  for (size_t i = 0; i < count; i++) {
    Fq zz = f_from_abi<Fq>(&raw[i].zz), zzz = f_from_abi<Fq>(&raw[i].zzz);
    den[i] = zz * zzz;
    pre[i] = acc;
    if (!zz.is_zero()) acc = acc * den[i];
  }
  Fq inv = Fq::inv(acc);
  std::vector<HostPoint> res(count);
  for (size_t i = count; i-- > 0;) {
    Fq zz = f_from_abi<Fq>(&raw[i].zz), zzz = f_from_abi<Fq>(&raw[i].zzz);
    HostPoint q;
    q.identity = zz.is_zero();
    if (q.identity) { q.x = Fq::zero(); q.y = Fq::zero(); }
    else {
      Fq t = inv * pre[i];                     // 1 / (zz * zzz)
      inv = inv * den[i];
      q.x = f_from_abi<Fq>(&raw[i].x) * (t * zzz);
      q.y = f_from_abi<Fq>(&raw[i].y) * (t * zz);
    }
    res[i] = q;
  }
  for (auto& q : res) out->push_back(q);
}
// Commitments of `count` contiguous polynomials into d_points[slot ..].  In a group the work is split (by column when
// there are at least as many polynomials as GPUs, by base range otherwise) and the points are all-gathered, so every
// rank continues with the same transcript.
int commit_balanced(zk_pkey* K, const char* d_polys, size_t count, int which, size_t slot);
int commit_enqueue(zk_pkey* K, const char* d_polys, size_t count, int which, size_t slot) {
  zk_ctx* ctx = K->ctx;
  PtRaw* out = (PtRaw*)K->d_points + slot;
  const uint32_t world = K->group ? zk_group_world(K->group) : 1;
  if (world <= 1 || !count) return zk_msm_batch_dev(ctx, (const zk_fr*)d_polys, which, K->n, count, (zk_g1*)out);
  const uint32_t rank = zk_group_rank(K->group);
  if (count >= world && count <= 256 && !getenv("ZK_GROUP_NO_BALANCE")) return commit_balanced(K, d_polys, count, which, slot);
  // A batch whose size is not a multiple of the group leaves whole columns of difference between the ranks (28 columns on 8
  // GPUs: four ranks commit 4, four commit 3).  The multiple of `world` goes by column, the remainder by base range, which
  // every rank shares equally.
  if (count > world && count % world != 0 && !getenv("ZK_GROUP_NO_HYBRID_SPLIT")) {
    const size_t whole = count - count % world;
    ZK_TRY(commit_enqueue(K, d_polys, whole, which, slot));
    return commit_enqueue(K, d_polys + whole * K->n * 32, count - whole, which, slot + whole);
  }
  int mode; size_t per, lo, hi;
  zk_group_shard_plan(rank, world, count, K->n, &mode, &per, &lo, &hi);
  PtRaw* xchg = (PtRaw*)K->d_xchg;
  // The exchange of the points is fused into the MSM: k_msm_final stores each result straight into every peer's window
  // and raises a flag when the rank's batch is complete (zk_p2p_push); the peers' wait kernel gathers them.  Without
  // mapped windows (ZK_GROUP_NO_P2P, IPC failure) the NCCL all-gather does the same.
  zk_p2p_push pp;
  const size_t seg_bytes = (mode == 0 ? per : count) * sizeof(PtRaw);
  const bool fused = zk_group_p2p_begin(K->group, ctx, seg_bytes, &pp);
  struct PushGuard { zk_ctx* c; ~PushGuard() { c->p2p = nullptr; } } guard{ctx};
  auto exchange = [&](bool launched_msm, PtRaw* mine) -> int {
    ctx->p2p = nullptr;
    if (!fused) return zk_group_allgather(K->group, ctx, xchg, seg_bytes);
    if (!launched_msm) ZK_TRY(zk_group_p2p_push(K->group, ctx, pp, mine, seg_bytes));      // nothing to commit here: push the zeros
    return zk_group_p2p_finish(K->group, ctx, xchg, seg_bytes);
  };
  if (mode == 0) {
    // by column, ROUND-ROBIN (column p belongs to rank p mod world): neighbouring columns of a region tend to hold the same
    // kind of values (bits and bytes are cheap commitments, field elements expensive), contiguous ranges do not balance
    auto owned = [&](uint32_t r) { return count > r ? (count - r + world - 1) / world : (size_t)0; };
    PtRaw* mine = xchg + (size_t)rank * per;
    if (fused && owned(rank) < per) ZK_CUDA(ctx, cudaMemsetAsync(mine, 0, per * sizeof(PtRaw), ctx->stream));
    if (owned(rank)) {
      if (fused) ctx->p2p = &pp;
      ZK_TRY(zk_msm_range_batch_dev(ctx, (const zk_fr*)(d_polys + (size_t)rank * K->n * 32), which, (size_t)world * K->n, 0, K->n, owned(rank),
                                    (zk_g1*)mine));
    }
    ZK_TRY(exchange(owned(rank) != 0, mine));
    for (uint32_t r = 0; r < world; r++)
      if (owned(r))
        ZK_CUDA(ctx, cudaMemcpy2DAsync(out + r, (size_t)world * sizeof(PtRaw), xchg + (size_t)r * per, sizeof(PtRaw), sizeof(PtRaw), owned(r),
                                       cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    PtRaw* mine = xchg + (size_t)rank * count;
    if (hi > lo) {
      if (fused) ctx->p2p = &pp;
      ZK_TRY(zk_msm_range_batch_dev(ctx, (const zk_fr*)d_polys, which, K->n, lo, hi - lo, count, (zk_g1*)mine));
    } else {
      ZK_CUDA(ctx, cudaMemsetAsync(mine, 0, count * sizeof(PtRaw), ctx->stream));   // zz = 0: identity
    }
    ZK_TRY(exchange(hi > lo, mine));
    ZK_TRY(zk_g1_sum_strided_dev(ctx, (const zk_g1*)xchg, world, count, (zk_g1*)out));
  }
  return ZK_OK;
}
// Cost-aware split of a batch of `count` >= world column commitments over the group.  Columns differ by an order of magnitude
// in cost (a column of bits has one non-zero window per scalar, a column of field elements sixteen): round-robin left
// msm_accumulate at 32.3 ms per proof on 2 GPUs and 10.7 ms on 8 against 28.5 / 7.2 for an even split.  Every rank measures the
// same per-column cost (non-zero 16-bit windows, one pass over the batch), so all ranks derive the same plan without talking
// (zk_group_balance_plan: longest first onto the least loaded rank).  A rank's columns are gathered into a contiguous scratch
// area when they are not adjacent (device copies, 16 MB per column), committed in ONE batch whose last kernel pushes the points
// to the peers, and a small kernel puts the points back in column order.  (Sharing the count mod G remainder by base range was
// tried first: the extra, tiny MSM pipeline per batch cost more latency than the even split saved.)
int commit_balanced(zk_pkey* K, const char* d_polys, size_t count, int which, size_t slot) {
  zk_ctx* ctx = K->ctx;
  const uint32_t world = zk_group_world(K->group), rank = zk_group_rank(K->group);
  const size_t nb = K->n * 32;
  PtRaw* out = (PtRaw*)K->d_points + slot;
  if (!K->d_cost) ZK_CUDA(ctx, cudaMalloc(&K->d_cost, 3 * 256 * sizeof(unsigned int)));
  ZK_CUDA(ctx, cudaMemsetAsync(K->d_cost, 0, 256 * sizeof(unsigned int), ctx->stream));
  {
    dim3 g(std::min<unsigned>(blocks_for(K->n, 256), 64), (unsigned)count);
    ZK_LAUNCH(ctx, "group_column_cost", k_column_cost, g, 256, 0, (const uint4*)d_polys, K->n, K->d_cost);
  }
  std::vector<unsigned int> cost(count);
  ZK_CUDA(ctx, cudaMemcpyAsync(cost.data(), K->d_cost, count * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<uint32_t> owner(count);
  zk_group_balance_plan(cost.data(), count, world, owner.data());
  std::vector<std::vector<uint32_t>> owned(world);
  for (size_t i = 0; i < count; i++) { if (owner[i] >= world) ZK_FAIL(ctx, ZK_ERR_STATE, "prover: balance plan"); owned[owner[i]].push_back((uint32_t)i); }
  size_t per = 0;
  for (auto& o : owned) per = std::max(per, o.size());
  // scratch: my columns, contiguous
  const size_t need = per * nb;
  if (K->balance_cap < need) {
    if (K->d_balance) { ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(K->d_balance); K->d_balance = nullptr; K->balance_cap = 0; }
    ZK_CUDA(ctx, cudaMalloc(&K->d_balance, need));
    K->balance_cap = need;
  }
  const std::vector<uint32_t>& mine_cols = owned[rank];
  const char* src = K->d_balance;
  bool contiguous = !mine_cols.empty();
  for (size_t j = 1; j < mine_cols.size(); j++) contiguous = contiguous && mine_cols[j] == mine_cols[j - 1] + 1;
  if (contiguous) {
    src = d_polys + (size_t)mine_cols[0] * nb;                  // already adjacent: commit in place
  } else {
    for (size_t j = 0; j < mine_cols.size();) {                 // runs of adjacent columns in one copy
      size_t e = j + 1;
      while (e < mine_cols.size() && mine_cols[e] == mine_cols[e - 1] + 1) e++;
      ZK_CUDA(ctx, cudaMemcpyAsync(K->d_balance + j * nb, d_polys + (size_t)mine_cols[j] * nb, (e - j) * nb, cudaMemcpyDeviceToDevice, ctx->stream));
      j = e;
    }
  }
  PtRaw* xchg = (PtRaw*)K->d_xchg;
  unsigned int* d_map = K->d_cost + 256;
  struct PushGuard { zk_ctx* c; ~PushGuard() { c->p2p = nullptr; } } guard{ctx};
  zk_p2p_push pp;
  const size_t seg = per * sizeof(PtRaw);
  const bool fused = zk_group_p2p_begin(K->group, ctx, seg, &pp);
  PtRaw* mine = xchg + (size_t)rank * per;
  if (!mine_cols.empty()) {
    if (fused) ctx->p2p = &pp;
    ZK_TRY(zk_msm_batch_dev(ctx, (const zk_fr*)src, which, K->n, mine_cols.size(), (zk_g1*)mine));
    ctx->p2p = nullptr;
  } else if (fused) {
    ZK_CUDA(ctx, cudaMemsetAsync(mine, 0, seg, ctx->stream));
    ZK_TRY(zk_group_p2p_push(K->group, ctx, pp, mine, seg));
  }
  if (fused) ZK_TRY(zk_group_p2p_finish(K->group, ctx, xchg, seg)); else ZK_TRY(zk_group_allgather(K->group, ctx, xchg, seg));
  // xchg[r * per + j] = the point of owned[r][j]
  std::vector<unsigned int> map((size_t)world * per, 0xFFFFFFFFu);
  for (uint32_t r = 0; r < world; r++) for (size_t j = 0; j < owned[r].size(); j++) map[(size_t)r * per + j] = owned[r][j];
  ZK_CUDA(ctx, cudaMemcpyAsync(d_map, map.data(), map.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
  ZK_LAUNCH(ctx, "group_scatter", k_scatter_points, blocks_for(map.size() * 8, 128), 128, 0, (const uint4*)xchg, (const unsigned int*)d_map, (uint4*)out,
            (unsigned int)map.size());
  return ZK_OK;
}
int commit_fetch(zk_pkey* K, size_t count, std::vector<HostPoint>* out) {
  zk_ctx* ctx = K->ctx;
  std::vector<PtRaw> h(count);
  ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), K->d_points, count * sizeof(PtRaw), cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  normalize_points(h.data(), count, out);
  return ZK_OK;
}
int commit(zk_pkey* K, const char* d_polys, size_t count, int which, std::vector<HostPoint>* out) {
  ZK_TRY(commit_enqueue(K, d_polys, count, which, 0));
  return commit_fetch(K, count, out);
}
bool is_device_pointer(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// lagrange_to_coeff + coeff_to_extended of witness columns on domain `d` (= on d's stream): values (Lagrange) -> poly
// (coefficients) and the coset in its evaluate_h slot.  In a group every column has an OWNER rank that transforms it;
// the coefficient forms are all-gathered through a staging buffer (every rank needs all of them for the openings), the
// cosets are written straight into their slots and only the row slices each rank's part of h(X) reads are exchanged.
struct ColJob { const char* values; char* poly; zk_fr* slot; uint32_t owner; };
// coeff_to_extended of jobs [i, j) whose coefficient forms are contiguous from `polys`: columns whose evaluate_h slots are
// a constant distance apart (adjacent slots, or every third slot for one plane of the lookup triples) go in ONE batch of
// launches -- single 2^21 transforms leave a 13 % tail on 148 SMs (3.46 waves), batches do not
int extend_run(zk_pkey* K, zk_domain* d, const std::vector<ColJob>& jobs, size_t i, size_t j, const char* polys, size_t nb) {
  const ptrdiff_t ext_bytes = (ptrdiff_t)K->ext_n * 32;
  for (size_t a = i; a < j;) {
    size_t b = a + 1;
    ptrdiff_t step = b < j ? (char*)jobs[b].slot - (char*)jobs[a].slot : ext_bytes;
    if (step >= ext_bytes && step % ext_bytes == 0) {
      while (b < j && (char*)jobs[b].slot - (char*)jobs[b - 1].slot == step) b++;
    } else {
      step = ext_bytes;
    }
    ZK_TRY(zk_coeff_to_extended_strided(d, (const zk_fr*)(polys + (a - i) * nb), jobs[a].slot, b - a, (size_t)step / 32));
    a = b;
  }
  return ZK_OK;
}
int transform_jobs(zk_pkey* K, zk_domain* d, zk_ctx* on, const std::vector<ColJob>& jobs) {
  const size_t nb = K->n * 32;
  const uint32_t world = K->group ? zk_group_world(K->group) : 1;
  const size_t count = jobs.size();
  if (!count) return ZK_OK;
  if (world <= 1 || !K->d_stage_p) {
    // batch runs of adjacent columns into one launch sequence (the cosets only when their slots are adjacent too)
    for (size_t i = 0; i < count;) {
      size_t j = i + 1;
      while (j < count && jobs[j].values == jobs[i].values + (j - i) * nb && jobs[j].poly == jobs[i].poly + (j - i) * nb) j++;
      ZK_TRY(zk_lagrange_to_coeff_oop(d, (const zk_fr*)jobs[i].values, (zk_fr*)jobs[i].poly, j - i));
      ZK_TRY(extend_run(K, d, jobs, i, j, jobs[i].poly, nb));
      i = j;
    }
    return ZK_OK;
  }
  const uint32_t rank = zk_group_rank(K->group);
  std::vector<size_t> owned(world, 0), kth(count);
  for (size_t i = 0; i < count; i++) { if (jobs[i].owner >= world) ZK_FAIL(on, ZK_ERR_STATE, "prover: column owner out of range"); kth[i] = owned[jobs[i].owner]++; }
  size_t maxown = 0;
  for (auto c : owned) maxown = std::max(maxown, c);
  if (maxown * world > K->stage_cols) ZK_FAIL(on, ZK_ERR_STATE, "prover: group staging holds %zu columns, %zu needed", K->stage_cols, maxown * world);
  char* sp = K->d_stage_p;
  for (size_t i = 0; i < count;) {
    if (jobs[i].owner != rank) { i++; continue; }
    // a run of my columns that are adjacent in memory (their staging places are adjacent by construction)
    size_t j = i + 1;
    while (j < count && jobs[j].owner == rank && kth[j] == kth[i] + (j - i) && jobs[j].values == jobs[i].values + (j - i) * nb) j++;
    char* mine = sp + ((size_t)rank * maxown + kth[i]) * nb;
    ZK_TRY(zk_lagrange_to_coeff_oop(d, (const zk_fr*)jobs[i].values, (zk_fr*)mine, j - i));
    ZK_TRY(extend_run(K, d, jobs, i, j, mine, nb));
    i = j;
  }
  ZK_TRY(zk_group_allgather(K->group, on, sp, maxown * nb));
  std::vector<char*> cols(count);
  std::vector<uint32_t> owners(count);
  for (size_t i = 0; i < count; i++) {
    ZK_CUDA(on, cudaMemcpyAsync(jobs[i].poly, sp + ((size_t)jobs[i].owner * maxown + kth[i]) * nb, nb, cudaMemcpyDeviceToDevice, on->stream));
    cols[i] = (char*)jobs[i].slot; owners[i] = jobs[i].owner;
  }
  return zk_group_exchange_rows(K->group, on, cols.data(), owners.data(), count, K->ext_n, K->halo_rows);
}
// contiguous columns split evenly over the group (advice, permutation z)
template <class SlotFn>
int transform_columns(zk_pkey* K, zk_domain* d, zk_ctx* on, const char* values, char* polys, size_t count, SlotFn slot) {
  const size_t nb = K->n * 32;
  const uint32_t world = K->group ? zk_group_world(K->group) : 1;
  const size_t per = (count + world - 1) / world;
  std::vector<ColJob> jobs(count);
  for (size_t i = 0; i < count; i++) jobs[i] = {values + i * nb, polys + i * nb, slot(i), (uint32_t)(i / per)};
  return transform_jobs(K, d, on, jobs);
}

// Commitments of arbitrary columns with an owner each: a rank commits the columns it owns (adjacent ones in one batch),
// the points are all-gathered and returned in list order.  With no group everything is local.
struct CommitJob { const char* poly; uint32_t owner; };
int commit_jobs(zk_pkey* K, const std::vector<CommitJob>& jobs, int which, std::vector<HostPoint>* out) {
  zk_ctx* ctx = K->ctx;
  const size_t nb = K->n * 32, count = jobs.size();
  const uint32_t world = K->group ? zk_group_world(K->group) : 1, rank = K->group ? zk_group_rank(K->group) : 0;
  if (!count) return ZK_OK;
  std::vector<size_t> owned(world, 0), kth(count);
  for (size_t i = 0; i < count; i++) { uint32_t o = world > 1 ? jobs[i].owner : 0; kth[i] = owned[o]++; }
  size_t maxown = 0;
  for (auto c : owned) maxown = std::max(maxown, c);
  PtRaw* xchg = (PtRaw*)K->d_xchg;
  for (size_t i = 0; i < count;) {
    uint32_t o = world > 1 ? jobs[i].owner : 0;
    size_t j = i + 1;
    while (j < count && (world > 1 ? jobs[j].owner : 0) == o && jobs[j].poly == jobs[i].poly + (j - i) * nb) j++;
    if (o == rank) ZK_TRY(zk_msm_batch_dev(ctx, (const zk_fr*)jobs[i].poly, which, K->n, j - i, (zk_g1*)(xchg + (size_t)rank * maxown + kth[i])));
    i = j;
  }
  if (world > 1) ZK_TRY(zk_group_allgather(K->group, ctx, xchg, maxown * sizeof(PtRaw)));
  std::vector<PtRaw> h((size_t)world * maxown), ordered(count);
  ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), xchg, h.size() * sizeof(PtRaw), cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < count; i++) ordered[i] = h[(size_t)(world > 1 ? jobs[i].owner : 0) * maxown + kth[i]];
  normalize_points(ordered.data(), count, out);
  return ZK_OK;
}

// ---- lookup: permute_expression_pair on the device, all lookup arguments at once --------------------------------------
// Sort s in [0, L) = compressed inputs of lookup s, [L, 2L) = compressed tables: planes 0 and 1 of d_lookup.
int radix_passes(zk_pkey* K, size_t usable, unsigned nsorts, const std::vector<std::vector<uint32_t>>& digits) {
  zk_ctx* ctx = K->ctx;
  size_t passes = 0;
  for (auto& d : digits) passes = std::max(passes, d.size());
  unsigned nb = blocks_for(usable, kRadixTile);
  dim3 grid(nb, nsorts);
  std::vector<uint32_t> tab(passes * nsorts);
  for (size_t p = 0; p < passes; p++)
    for (unsigned s_ = 0; s_ < nsorts; s_++) tab[p * nsorts + s_] = p < digits[s_].size() ? digits[s_][p] : kNullDigit;
  if (tab.size() > 32 * 2 * (size_t)kMaxLookups) ZK_FAIL(ctx, ZK_ERR_STATE, "prover: digit table");
  ZK_CUDA(ctx, cudaMemcpyAsync(K->d_digits, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t *idx = K->d_idx[0], *alt = K->d_idx[1];
  for (size_t p = 0; p < passes; p++) {
    const uint32_t* dg = K->d_digits + p * nsorts;
    ZK_LAUNCH(ctx, "lookup_sort", k_radix_hist, grid, kRadixThreads, 0, (const uint32_t*)K->d_keys, idx, usable, K->n, dg, K->d_hist);
    ZK_TRY(run_scan<U32AddMonoid>(K, "lookup_sort", K->d_hist, K->d_hist, (size_t)256 * nb * nsorts, 0, 1, nullptr));
    ZK_LAUNCH(ctx, "lookup_sort", k_radix_scatter, grid, kRadixThreads, 0, (const uint32_t*)K->d_keys, idx, alt, usable, K->n, dg, K->d_hist);
    std::swap(idx, alt);
  }
  K->d_idx[0] = idx; K->d_idx[1] = alt;          // d_idx[0] always holds the current order
  return ZK_OK;
}
// lookups [l0, l0 + Lloc) (all of them on one GPU; a rank's share in a group)
int permute_expression_pairs(zk_pkey* K, size_t usable, uint32_t l0, uint32_t Lloc) {
  zk_ctx* ctx = K->ctx;
  const uint32_t Ltot = K->n_lookups;
  if (!Lloc) return ZK_OK;
  const unsigned nsorts = 2 * Lloc;
  const size_t n = K->n;
  char* vals = K->d_lookup;                      // planes 0 (cin) and 1 (ctab) are adjacent: Ltot columns each
  dim3 g256(blocks_for(usable, 256), nsorts);
  ZK_LAUNCH(ctx, "lookup_sort", k_canon_keys, g256, 256, 0, (const uint4*)vals, (uint4*)K->d_keys, K->d_idx[0], usable, n, Lloc, l0, Ltot);
  ZK_CUDA(ctx, cudaMemsetAsync(K->d_masks, 0, (size_t)nsorts * 32, ctx->stream));
  ZK_LAUNCH(ctx, "lookup_sort", k_key_diff_mask, dim3(32, nsorts), 256, 0, (const uint32_t*)K->d_keys, usable, n, K->d_masks);
  std::vector<uint32_t> masks((size_t)nsorts * 8);
  ZK_CUDA(ctx, cudaMemcpyAsync(masks.data(), K->d_masks, masks.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // theta-compressed values look uniform: the 8 most significant varying digits (64 bits) order them unless two
  // distinct values agree on all of those digits.  Sort by them first and verify; fall back to every varying digit.
  const size_t kTop = 8;
  std::vector<std::vector<uint32_t>> varying(nsorts), first(nsorts);
  bool partial = false;
  for (unsigned s_ = 0; s_ < nsorts; s_++) {
    for (uint32_t d = 0; d < 32; d++) if ((masks[s_ * 8 + (d >> 2)] >> (8 * (d & 3))) & 0xFFu) varying[s_].push_back(d);
    if (varying[s_].size() > kTop) { partial = true; first[s_].assign(varying[s_].end() - kTop, varying[s_].end()); }
    else first[s_] = varying[s_];
  }
  ZK_TRY(radix_passes(K, usable, nsorts, first));
  ZK_LAUNCH(ctx, "lookup_sort", k_gather_keys, g256, 256, 0, (const uint4*)K->d_keys, K->d_idx[0], (uint4*)K->d_sk, usable, n);
  if (partial) {
    ZK_CUDA(ctx, cudaMemsetAsync(K->d_flags + 3, 0, 4, ctx->stream));
    ZK_LAUNCH(ctx, "lookup_sort", k_check_sorted, g256, 256, 0, (const uint32_t*)K->d_sk, usable, n, K->d_flags + 3);
    uint32_t bad = 0;
    ZK_CUDA(ctx, cudaMemcpyAsync(&bad, K->d_flags + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) {      // LSD needs the low digits first: redo from the identity order over every varying digit
      K->sort_fallbacks++;
      ZK_LAUNCH(ctx, "lookup_sort", k_canon_keys, g256, 256, 0, (const uint4*)vals, (uint4*)K->d_keys, K->d_idx[0], usable, n, Lloc, l0, Ltot);
      ZK_TRY(radix_passes(K, usable, nsorts, varying));
      ZK_LAUNCH(ctx, "lookup_sort", k_gather_keys, g256, 256, 0, (const uint4*)K->d_keys, K->d_idx[0], (uint4*)K->d_sk, usable, n);
    }
  }
  uint32_t *rep = K->d_u32[0], *left = K->d_u32[1], *rrank = K->d_u32[2], *lrank = K->d_u32[3], *rpos = K->d_u32[4];
  dim3 gl(blocks_for(usable, 256), Lloc);
  ZK_LAUNCH(ctx, "lookup_permute", k_permute_flags, gl, 256, 0, (const uint32_t*)K->d_sk, usable, n, Lloc, rep, left, K->d_flags);
  ZK_TRY(run_scan<U32AddMonoid>(K, "lookup_permute", rep, rrank, usable, 0, 1, nullptr, Lloc, usable, usable));
  ZK_TRY(run_scan<U32AddMonoid>(K, "lookup_permute", left, lrank, usable, 0, 1, nullptr, Lloc, usable, usable));
  ZK_LAUNCH(ctx, "lookup_permute", k_permute_totals, blocks_for(Lloc, 64), 64, 0, rep, rrank, left, lrank, usable, Lloc, K->d_totals, K->d_flags);
  ZK_LAUNCH(ctx, "lookup_permute", k_permute_rpos, gl, 256, 0, rep, rrank, rpos, usable);
  char* cin = K->d_lookup; char* ctab = cin + (size_t)Ltot * n * 32; char* pin = ctab + (size_t)Ltot * n * 32; char* ptab = pin + (size_t)Ltot * n * 32;
  ZK_LAUNCH(ctx, "lookup_permute", k_permute_write, gl, 256, 0, (const uint4*)cin, (const uint4*)ctab, K->d_idx[0], rep, left, lrank, rpos,
            K->d_totals, Lloc, (uint4*)pin, (uint4*)ptab, usable, n, l0);
  return ZK_OK;
}

// lookup compression: graph `gi` of the compress program over the n Lagrange rows
int eval_rows(zk_pkey* K, uint32_t gi, const EvalScalars& S, char* out) {
  zk_ctx* ctx = K->ctx;
  uint32_t ni = K->CL.max_intermediates;
  unsigned grid = blocks_for(K->n, 128);
#define ROWS_ARGS K->CL.prog, S, gi, (const uint32_t*)K->d_lcode, (const uint4* const*)K->d_lcols, (const uint4*)K->d_lconsts, \
                  (const uint4*)K->d_challenges, (uint4*)out, K->n
  if (ni <= 16) { ZK_LAUNCH(ctx, "lookup_compress", k_eval_rows<16>, grid, 128, 0, ROWS_ARGS); }
  else if (ni <= 128) { ZK_LAUNCH(ctx, "lookup_compress", k_eval_rows<128>, grid, 128, 0, ROWS_ARGS); }
  else { ZK_LAUNCH(ctx, "lookup_compress", k_eval_rows<512>, grid, 128, 0, ROWS_ARGS); }
#undef ROWS_ARGS
  return ZK_OK;
}

// ---- host-side helpers of SHPLONK ----------------------------------------------------------------------------------
// Montgomery's trick on the host (a host field inversion is ~380 emulated multiplications: worth sharing)
void host_batch_invert(std::vector<Fr>& v) {
  std::vector<Fr> pre(v.size());
  Fr acc = Fr::one();
  for (size_t i = 0; i < v.size(); i++) { pre[i] = acc; acc = acc * v[i]; }
  Fr inv = Fr::inv(acc);
  for (size_t i = v.size(); i-- > 0;) { Fr t = inv * pre[i]; inv = inv * v[i]; v[i] = t; }
}
// Lagrange basis of a point set: basis[j] = coefficients (low -> high) of prod_{i != j} (X - x_i) / prod_{i != j} (x_j - x_i).
// Every commitment of a rotation set interpolates through the SAME points, so the basis is built once per set.
std::vector<std::vector<Fr>> lagrange_basis(const std::vector<Fr>& pts) {
  size_t m = pts.size();
  std::vector<std::vector<Fr>> basis(m);
  std::vector<Fr> den(m, Fr::one());
  for (size_t j = 0; j < m; j++) {
    std::vector<Fr> num(1, Fr::one());
    for (size_t i = 0; i < m; i++) {
      if (i == j) continue;
      std::vector<Fr> nxt(num.size() + 1, Fr::zero());
      for (size_t t = 0; t < num.size(); t++) {
        nxt[t] = nxt[t] - num[t] * pts[i];
        nxt[t + 1] = nxt[t + 1] + num[t];
      }
      num.swap(nxt);
      den[j] = den[j] * (pts[j] - pts[i]);
    }
    basis[j] = num;
  }
  if (m > 1) host_batch_invert(den);          // m == 1: the denominator is the empty product
  for (size_t j = 0; j < m; j++)
    for (auto& c : basis[j]) c = c * den[j];
  return basis;
}
// coefficients (low -> high) of the polynomial of degree < m through (points, evals)
std::vector<Fr> lagrange_interpolate(const std::vector<std::vector<Fr>>& basis, const std::vector<Fr>& evals) {
  size_t m = basis.size();
  std::vector<Fr> out(m, Fr::zero());
  for (size_t j = 0; j < m; j++)
    for (size_t t = 0; t < m; t++) out[t] = out[t] + basis[j][t] * evals[j];
  return out;
}
Fr eval_small(const std::vector<Fr>& c, const Fr& x) {
  Fr acc = Fr::zero();
  for (size_t i = c.size(); i-- > 0;) acc = acc * x + c[i];
  return acc;
}

struct Query { const void* poly; uint32_t point; Fr eval; };      // point = index into the list of opening points
struct RotationSet {
  std::vector<uint32_t> points;                                    // ascending by field value
  std::vector<const void*> polys;
  std::vector<std::vector<Fr>> evals;
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
extern "C" void zk_pkey_destroy(zk_pkey* K) {
  if (!K) return;
  cudaSetDevice(K->ctx->device);
  cudaStreamSynchronize(K->ctx->stream);
  for (void* p : K->allocs) cudaFree(p);
  if (K->d_balance) cudaFree(K->d_balance);
  if (K->d_cost) cudaFree(K->d_cost);
  for (char* p : K->d_setacc) cudaFree(p);
  if (K->ctxB) cudaStreamSynchronize(K->ctxB->stream);
  for (auto& e : K->ev_in) if (e) cudaEventDestroy(e);
  for (auto& e : K->ev_copy) if (e) cudaEventDestroy(e);
  if (K->ev_done) cudaEventDestroy(K->ev_done);
  if (K->pk) zk_pk_destroy(K->pk);
  if (K->dom) zk_domain_destroy(K->dom);
  if (K->domB) zk_domain_destroy(K->domB);
  if (K->ctxB) zk_ctx_destroy(K->ctxB);
  delete K;
}
extern "C" zk_pk* zk_pkey_pk(zk_pkey* K) { return K ? K->pk : nullptr; }
extern "C" zk_ctx* zk_pkey_aux_ctx(zk_pkey* K) { return K ? K->ctxB : nullptr; }
extern "C" int zk_pkey_set_transcript(zk_pkey* K, int kind) {
  if (!K) return ZK_ERR_ARG;
  if (kind != ZK_TRANSCRIPT_BLAKE2B && kind != ZK_TRANSCRIPT_EVM && kind != ZK_TRANSCRIPT_POSEIDON) ZK_FAIL(K->ctx, ZK_ERR_ARG, "zk_pkey_set_transcript: unknown transcript kind %d", kind);
  K->transcript_kind = kind;
  return ZK_OK;
}
// The plan of a cost-balanced batch of commitments (host logic, also exercised by the CPU tests): owner[i] = the rank that
// commits column i.  Columns are dealt longest-first to the least loaded rank that still has room for one more column
// (ceil(count / world) per rank); ties go to the lower index / rank, so the plan is a pure function of the costs.  With costs
// spread over a factor of sixteen (columns of bits next to columns of field elements) the ranks that get one column more get
// cheap ones, which evens the loads out without splitting any column.
extern "C" void zk_group_balance_plan(const uint32_t* cost, size_t count, uint32_t world, uint32_t* owner) {
  if (!world) return;
  std::vector<uint32_t> order(count);
  for (size_t i = 0; i < count; i++) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost[a] > cost[b]; });
  const size_t cap = (count + world - 1) / world;
  std::vector<uint64_t> load(world, 0);
  std::vector<size_t> have(world, 0);
  for (size_t i = 0; i < count; i++) {
    uint32_t best = world;
    for (uint32_t r = 0; r < world; r++)
      if (have[r] < cap && (best == world || load[r] < load[best])) best = r;
    owner[order[i]] = best; have[best]++; load[best] += (uint64_t)cost[order[i]] + 1;
  }
}

extern "C" int zk_pkey_set_group(zk_pkey* K, zk_group* g) {
  if (!K) return ZK_ERR_ARG;
  if (g && zk_group_world(g) > 16) ZK_FAIL(K->ctx, ZK_ERR_ARG, "zk_pkey_set_group: at most 16 GPUs");
  K->group = g;
  if (g && zk_group_world(g) > 1 && !K->d_stage_p && !getenv("ZK_GROUP_NO_NTT_SHARD")) {
    zk_ctx* ctx = K->ctx;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t cols = std::max<size_t>(std::max<size_t>(K->n_advice, 2 * (size_t)K->n_lookups), K->n_sets) + 2 * zk_group_world(g);
    void* p = nullptr;
    ZK_CUDA(ctx, cudaMalloc(&p, cols * K->n * 32));
    K->allocs.push_back(p);
    K->d_stage_p = (char*)p; K->stage_cols = cols;
    // rows of the extended domain that evaluate_h reads relative to its own row: rotation * 2^(ext_k - k)
    K->halo_rows = (size_t)K->max_rotation * (K->ext_n / K->n);
  }
  // the evaluate_h witness slots have the same layout on every rank: map the peers' copies, so that the row exchange of the
  // cosets becomes plain device-to-device copies into the peers' slots (collective: every rank makes this call)
  if (g && zk_group_world(g) > 1 && K->pk) ZK_TRY(zk_group_register(g, zk_pk_witness_slot(K->pk, 0), zk_pk_witness_bytes(K->pk)));
  // likewise the staging area of the coefficient forms and the h(X) rows (their all-gathers become peer copies)
  if (g && zk_group_world(g) > 1 && K->d_stage_p) ZK_TRY(zk_group_register(g, K->d_stage_p, K->stage_cols * K->n * 32));
  if (g && zk_group_world(g) > 1 && K->d_hext) ZK_TRY(zk_group_register(g, K->d_hext, K->ext_n * 32));
  if (g && zk_group_world(g) > 1 && K->d_lookup) ZK_TRY(zk_group_register(g, K->d_lookup, ((size_t)K->n_lookups * 5 + K->n_sets) * K->n * 32));
  return ZK_OK;
}
extern "C" size_t zk_proof_randoms(const zk_pkey* K) {
  if (!K) return 0;
  size_t a = K->n_advice, l = K->n_lookups, bf = K->bf;
  return a * (bf + 1) + a + l * (2 * (bf + 1) + 2) + (size_t)K->n_sets * (bf + 1) + l * (bf + 1) + K->n + 1 + (K->degree - 1);
}
extern "C" size_t zk_proof_max_bytes(const zk_pkey* K) {
  if (!K) return 0;
  size_t points = K->n_advice + 2 * (size_t)K->n_lookups + K->n_sets + K->n_lookups + 1 + (K->degree - 1) + 2;
  size_t scalars = K->advice_queries.size() + K->fixed_queries.size() + 1 + K->n_perm + 3 * (size_t)K->n_sets + 5 * (size_t)K->n_lookups;
  return (K->transcript_kind == ZK_TRANSCRIPT_EVM ? 64 : 32) * points + 32 * scalars;
}
extern "C" int zk_pkey_vk_commitments(const zk_pkey* K, zk_g1_affine* fixed_commitments, zk_g1_affine* perm_commitments) {
  if (!K) return ZK_ERR_ARG;
  if (fixed_commitments && !K->fixed_commitments.empty()) memcpy(fixed_commitments, K->fixed_commitments.data(), K->fixed_commitments.size() * sizeof(zk_g1_affine));
  if (perm_commitments && !K->perm_commitments.empty()) memcpy(perm_commitments, K->perm_commitments.data(), K->perm_commitments.size() * sizeof(zk_g1_affine));
  return ZK_OK;
}

extern "C" int zk_keygen(zk_ctx* ctx, const zk_circuit_desc* cd, const zk_fr* const* fixed_values, const uint32_t* perm_mapping,
                         zk_pkey** out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!cd || !out || (cd->ev.n_fixed && !fixed_values) || (cd->ev.n_perm_columns && !perm_mapping) ||
      (cd->ev.n_lookups && (!cd->lookup_inputs || !cd->lookup_tables)) || (cd->n_advice_queries && !cd->advice_queries) ||
      (cd->n_fixed_queries && !cd->fixed_queries))
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: bad argument");
  const zk_evaluator_desc& D = cd->ev;
  if (D.degree < 3) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: cs.degree() must be >= 3");
  if (D.n_lookups * 2 > (uint32_t)kMaxLookups) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: more than %d lookup arguments", kMaxLookups / 2);
  if (D.degree - 2 > 8) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: permutation chunk of %u columns exceeds 8", D.degree - 2);
  if (ctx->srs_k < D.k || !ctx->srs_g.bases || !ctx->srs_lagrange.bases) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_keygen: upload an SRS with k >= %u first", D.k);
  if (ctx->srs_k != D.k) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_keygen: the resident SRS has k = %u, the circuit k = %u (g_lagrange is per k)", ctx->srs_k, D.k);
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  zk_pkey* K = new zk_pkey();
  K->ctx = ctx;
#define KG(expr) do { int _r = (expr); if (_r != ZK_OK) { zk_pkey_destroy(K); return _r; } } while (0)
#define KG_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string("zk_keygen: ") + #expr + ": " + cudaGetErrorString(_e); zk_pkey_destroy(K); return _e == cudaErrorMemoryAllocation ? ZK_ERR_OOM : ZK_ERR_CUDA; } } while (0)
  K->k = D.k; K->degree = D.degree; K->bf = D.blinding_factors;
  K->n_fixed = D.n_fixed; K->n_advice = D.n_advice; K->n_instance = D.n_instance; K->n_challenges = D.n_challenges;
  K->n_lookups = D.n_lookups; K->n_perm = D.n_perm_columns; K->chunk = D.degree - 2;
  K->n_sets = K->n_perm ? (K->n_perm + K->chunk - 1) / K->chunk : 0;
  K->n = (size_t)1 << D.k;
  if (K->n < (size_t)K->bf + 3) { zk_pkey_destroy(K); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: not enough rows (n = %zu, minimum_rows = %u)", (size_t)1 << D.k, D.blinding_factors + 3); }
  K->perm_columns.assign(D.perm_columns, D.perm_columns + D.n_perm_columns);
  K->advice_queries.assign(cd->advice_queries, cd->advice_queries + cd->n_advice_queries);
  K->fixed_queries.assign(cd->fixed_queries, cd->fixed_queries + cd->n_fixed_queries);
  for (auto& q : K->advice_queries) if (q.column >= D.n_advice) { zk_pkey_destroy(K); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: advice query column out of range"); }
  for (auto& q : K->fixed_queries) if (q.column >= D.n_fixed) { zk_pkey_destroy(K); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: fixed query column out of range"); }
  K->transcript_repr = f_from_abi<Fr>(&cd->transcript_repr);
  {
    auto scan = [&](const zk_graph& g) {
      for (uint32_t i = 0; i < g.n_rotations; i++) K->max_rotation = std::max<uint32_t>(K->max_rotation, (uint32_t)abs(g.rotations[i]));
    };
    scan(D.custom_gates);
    for (uint32_t l = 0; l < D.n_lookups; l++) scan(D.lookups[l]);
    if (D.n_lookups) K->max_rotation = std::max<uint32_t>(K->max_rotation, 1);
    if (D.n_perm_columns) K->max_rotation = std::max<uint32_t>(K->max_rotation, D.blinding_factors + 1);
  }
  KG(zk_domain_create(ctx, D.degree, D.k, &K->dom));
  K->ext_n = zk_domain_ext_n(K->dom);
  for (auto& e : K->ev_copy) KG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (!(getenv("ZK_PROVER_STREAMS") && atoi(getenv("ZK_PROVER_STREAMS")) == 1)) {
    int r = zk_ctx_create(ctx->device, nullptr, &K->ctxB);
    if (r != ZK_OK) { ctx->err = "zk_keygen: cannot create the second stream"; zk_pkey_destroy(K); return r; }
    r = zk_domain_create(K->ctxB, D.degree, D.k, &K->domB);
    if (r != ZK_OK) { ctx->err = "zk_keygen: " + K->ctxB->err; zk_pkey_destroy(K); return r; }
    for (auto& e : K->ev_in) KG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    KG_CUDA(cudaEventCreateWithFlags(&K->ev_done, cudaEventDisableTiming));
  }
  zk_fr dc[8];
  zk_domain_constants(K->dom, dc);
  K->omega = f_from_abi<Fr>(&dc[0]); K->omega_inv = f_from_abi<Fr>(&dc[1]);
  const size_t n = K->n, nb = n * 32;

  // ---- lookup compression program: slots = fixed | advice | instance Lagrange columns -----------------------------
  {
    CompiledEval& C = K->CL;
    C = CompiledEval();
    EvalProgram& P = C.prog;
    memset(&P, 0, sizeof P);
    P.k = D.k; P.ext_k = D.k;
    P.off_fixed = 0; P.off_advice = D.n_fixed; P.off_instance = D.n_fixed + D.n_advice;
    C.n_slots = D.n_fixed + D.n_advice + D.n_instance;
    std::set<std::pair<uint32_t, uint32_t>> cols;
    for (uint32_t l = 0; l < D.n_lookups; l++) {
      if (!detail::compile_graph(cd->lookup_inputs[l], D, C, &P.lookups[2 * l], cols, true) ||
          !detail::compile_graph(cd->lookup_tables[l], D, C, &P.lookups[2 * l + 1], cols, true)) {
        std::string e = C.error; zk_pkey_destroy(K); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: lookup %u: %s", l, e.c_str());
      }
    }
    if (C.max_intermediates > 512) { zk_pkey_destroy(K); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: lookup graph needs more than 512 intermediates"); }
    if (C.consts.empty()) { zk_fr z; memset(&z, 0, sizeof z); C.consts.push_back(z); }
    if (C.code.empty()) C.code.push_back(0);
    KG(dev_alloc(K, &K->d_lcode, C.code.size() * 4));
    KG(dev_alloc(K, &K->d_lconsts, C.consts.size() * 32));
    KG(dev_alloc(K, &K->d_lcols, (C.n_slots + 1) * sizeof(void*)));
    KG(dev_alloc(K, &K->d_challenges, (D.n_challenges + 1) * 32));
    KG_CUDA(cudaMemcpyAsync(K->d_lcode, C.code.data(), C.code.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    KG_CUDA(cudaMemcpyAsync(K->d_lconsts, C.consts.data(), C.consts.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
  }

  // ---- buffers ---------------------------------------------------------------------------------------------------------
  KG(dev_alloc(K, &K->d_fixed_values, D.n_fixed * nb)); KG(dev_alloc(K, &K->d_fixed_polys, D.n_fixed * nb));
  KG(dev_alloc(K, &K->d_sigma_values, K->n_perm * nb)); KG(dev_alloc(K, &K->d_sigma_polys, K->n_perm * nb));
  KG(dev_alloc(K, &K->d_advice, D.n_advice * nb));
  KG(dev_alloc(K, &K->d_advice_polys, D.n_advice * nb));
  KG(dev_alloc(K, &K->d_lookup_polys, (size_t)D.n_lookups * 3 * nb));
  KG(dev_alloc(K, &K->d_perm_z_polys, K->n_sets * nb));
  KG(dev_alloc(K, &K->d_instance, D.n_instance * nb)); KG(dev_alloc(K, &K->d_instance_polys, D.n_instance * nb));
  // the permutation products sit right behind plane 4 (Z) of the lookup area, so that [lookup Z | permutation z] is ONE
  // contiguous batch for the commitment (no challenge separates the two: both follow beta / gamma)
  KG(dev_alloc(K, &K->d_lookup, ((size_t)D.n_lookups * 5 + K->n_sets) * nb));
  K->d_perm_z = K->d_lookup + (size_t)D.n_lookups * 5 * nb;
  KG(dev_alloc(K, &K->d_random, nb)); KG(dev_alloc(K, &K->d_hext, K->ext_n * 32));
  KG(dev_alloc(K, &K->d_hpieces, (size_t)(D.degree - 1) * nb));
  KG(dev_alloc(K, &K->d_hpoly, nb)); KG(dev_alloc(K, &K->d_hx, nb)); KG(dev_alloc(K, &K->d_lx, nb));
  {   // w0 / w1 hold one vector per lookup argument or permutation set in the batched product kernels
    size_t seg = std::max<size_t>(std::max<size_t>(D.n_lookups, K->n_sets), 1);
    KG(dev_alloc(K, &K->d_w0, seg * nb)); KG(dev_alloc(K, &K->d_w1, seg * nb)); KG(dev_alloc(K, &K->d_w2, nb));
  }
  // opening points: one per distinct rotation (+ u), each with a table of powers and of inverse powers
  {
    std::set<int32_t> rots = {0, 1, -1, -(int32_t)(K->bf + 1)};
    for (auto& q : K->advice_queries) rots.insert(q.rotation);
    for (auto& q : K->fixed_queries) rots.insert(q.rotation);
    K->n_table_slots = 2 * ((uint32_t)rots.size() + 1);
  }
  KG(dev_alloc(K, &K->d_tables, (size_t)K->n_table_slots * nb));
  {
    const size_t L = D.n_lookups, seg = std::max<size_t>(std::max<size_t>(L, K->n_sets), 1);
    size_t radix_hist = (size_t)256 * blocks_for(n, kRadixTile) * 2 * L;
    K->scan_cap = std::max<size_t>(blocks_for(radix_hist, kScanTile), (size_t)blocks_for(n, kScanTile) * seg) + 1;
    KG(dev_alloc(K, &K->d_scan_totals, K->scan_cap * 32)); KG(dev_alloc(K, &K->d_seed, 32));
    KG(dev_alloc(K, &K->d_keys, 2 * L * nb)); KG(dev_alloc(K, &K->d_sk, 2 * L * nb));
    for (int i = 0; i < 2; i++) KG(dev_alloc(K, &K->d_idx[i], 2 * L * n * 4));
    for (int i = 0; i < 5; i++) KG(dev_alloc(K, &K->d_u32[i], L * n * 4));
    KG(dev_alloc(K, &K->d_hist, radix_hist * 4));
    KG(dev_alloc(K, &K->d_masks, 2 * L * 32 + 32)); KG(dev_alloc(K, &K->d_digits, 32 * 2 * (size_t)kMaxLookups * 4));
    KG(dev_alloc(K, &K->d_totals, L * 4 + 4));
    KG(dev_alloc(K, &K->d_permcols, (K->n_sets + 1) * sizeof(PermCols))); KG(dev_alloc(K, &K->d_permnum, (K->n_sets + 1) * sizeof(PermNum)));
    KG(dev_alloc(K, &K->d_seeds, (K->n_sets + 1) * 32));
  }
  KG(dev_alloc(K, &K->d_flags, 64));
  K->max_list = D.n_advice + D.n_fixed + D.n_instance + K->n_perm + K->n_sets + 3 * D.n_lookups + D.degree + 16;
  K->eval_chunks = (uint32_t)std::min<size_t>(std::max<size_t>(n / 4096, 1), 64);
  KG(dev_alloc(K, &K->d_points, (size_t)K->max_list * sizeof(PtRaw)));
  KG(dev_alloc(K, &K->d_xchg, (size_t)(K->max_list + 64) * 16 * sizeof(PtRaw)));
  KG(dev_alloc(K, &K->d_ptrs, (size_t)K->max_list * sizeof(void*)));
  KG(dev_alloc(K, &K->d_scalars, (size_t)K->max_list * 32));
  KG(dev_alloc(K, &K->d_partial, (size_t)K->max_list * K->eval_chunks * 32));
  KG(dev_alloc(K, &K->d_evals, (size_t)K->max_list * 32));

  // ---- fixed columns: values, coefficient forms ----------------------------------------------------------------------------
  for (uint32_t i = 0; i < D.n_fixed; i++) {
    if (!fixed_values[i]) { zk_pkey_destroy(K); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: null fixed column"); }
    KG_CUDA(cudaMemcpyAsync(K->d_fixed_values + i * nb, fixed_values[i], nb, cudaMemcpyHostToDevice, ctx->stream));
  }
  KG_CUDA(cudaMemcpyAsync(K->d_fixed_polys, K->d_fixed_values, D.n_fixed * nb, cudaMemcpyDeviceToDevice, ctx->stream));
  KG(zk_lagrange_to_coeff_dev(K->dom, (zk_fr*)K->d_fixed_polys, D.n_fixed));
  // ---- permutation: sigma_i[j] = delta^(i') * omega^(j') ----------------------------------------------------------------------
  const uint4* tw = (const uint4*)zk_domain_twiddles(K->dom);
  const uint32_t tw_shift = zk_domain_extended_k(K->dom) - D.k;
  const size_t half_ext = K->ext_n >> 1;
  if (K->n_perm) {
    for (size_t i = 0; i < (size_t)K->n_perm * n; i++)
      if (perm_mapping[2 * i] >= K->n_perm || perm_mapping[2 * i + 1] >= n) { zk_pkey_destroy(K); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_keygen: permutation mapping out of range"); }
    std::vector<zk_fr> dp(K->n_perm);
    Fr cur = Fr::one(), delta = fr_delta();
    for (uint32_t i = 0; i < K->n_perm; i++) { f_to_abi(cur, &dp[i]); cur = cur * delta; }
    KG_CUDA(cudaMemcpyAsync(K->d_scalars, dp.data(), dp.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    // the mapping is staged through the (still unused) sigma polynomial buffer: 8 bytes per cell
    KG_CUDA(cudaMemcpyAsync(K->d_sigma_polys, perm_mapping, (size_t)K->n_perm * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    zk_timing_begin(ctx, "keygen_sigma");
    k_sigma<<<blocks_for((size_t)K->n_perm * n, 256), 256, 0, ctx->stream>>>((const uint2*)K->d_sigma_polys, (const uint4*)K->d_scalars, tw, tw_shift,
                                                                           half_ext, (uint4*)K->d_sigma_values, (size_t)K->n_perm * n);
    zk_timing_end(ctx);
    KG_CUDA(cudaGetLastError());
    KG_CUDA(cudaMemcpyAsync(K->d_sigma_polys, K->d_sigma_values, K->n_perm * nb, cudaMemcpyDeviceToDevice, ctx->stream));
    KG(zk_lagrange_to_coeff_dev(K->dom, (zk_fr*)K->d_sigma_polys, K->n_perm));
  }
  // ---- l0, l_last, l_blind --------------------------------------------------------------------------------------------------------
  {
    char* ind = K->d_w0;   // three indicator columns need 3n: use w0, w1, w2
    char* cols3[3] = {K->d_w0, K->d_w1, K->d_w2};
    (void)ind;
    size_t lo[3] = {0, n - K->bf - 1, n - K->bf}, hi[3] = {1, n - K->bf, n};
    for (int i = 0; i < 3; i++) {
      zk_timing_begin(ctx, "keygen_indicator");
      k_indicator<<<blocks_for(n, 256), 256, 0, ctx->stream>>>((uint4*)cols3[i], n, lo[i], hi[i]);
      zk_timing_end(ctx);
      KG(zk_lagrange_to_coeff_dev(K->dom, (zk_fr*)cols3[i], 1));
    }
    std::vector<const zk_fr*> fx(D.n_fixed + 1), sg(K->n_perm + 1);
    for (uint32_t i = 0; i < D.n_fixed; i++) fx[i] = (const zk_fr*)(K->d_fixed_polys + i * nb);
    for (uint32_t i = 0; i < K->n_perm; i++) sg[i] = (const zk_fr*)(K->d_sigma_polys + i * nb);
    // coefficient forms (device pointers; the copies inside are cudaMemcpyDefault) -> cosets, l_active_row
    KG(zk_pk_upload(ctx, &D, fx.data(), sg.data(), (const zk_fr*)K->d_w0, (const zk_fr*)K->d_w1, (const zk_fr*)K->d_w2, 1, &K->pk));
  }
  // ---- keygen_vk: commitments of the fixed and permutation columns ------------------------------------------------------------------
  {
    RawGuard raw_guard(ctx);
    std::vector<HostPoint> pts;
    if (D.n_fixed) KG(commit(K, K->d_fixed_values, D.n_fixed, ZK_BASES_G_LAGRANGE, &pts));
    if (K->n_perm) KG(commit(K, K->d_sigma_values, K->n_perm, ZK_BASES_G_LAGRANGE, &pts));
    for (size_t i = 0; i < pts.size(); i++) {
      zk_g1_affine a; memset(&a, 0, sizeof a);
      if (!pts[i].identity) { f_to_abi(pts[i].x, &a.x); f_to_abi(pts[i].y, &a.y); }
      (i < D.n_fixed ? K->fixed_commitments : K->perm_commitments).push_back(a);
    }
  }
  KG_CUDA(cudaStreamSynchronize(ctx->stream));
#undef KG
#undef KG_CUDA
  *out = K;
  return ZK_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
extern "C" int zk_create_proof(zk_pkey* K, const zk_fr* const* advice_values, const zk_fr* const* instance_values,
                               const uint32_t* instance_lens, const zk_fr* randoms, size_t n_randoms, uint8_t* proof,
                               size_t proof_cap, size_t* proof_len) {
  if (!K) return ZK_ERR_ARG;
  zk_ctx* ctx = K->ctx;
  if (!proof || !proof_len || !randoms || (K->n_advice && !advice_values) || (K->n_instance && (!instance_values || !instance_lens)))
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: bad argument");
  if (n_randoms < zk_proof_randoms(K)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: %zu randoms supplied, %zu needed", n_randoms, zk_proof_randoms(K));
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  RawGuard raw_guard(ctx);                              // commitments come back as XYZZ records, normalised on the host
  const size_t n = K->n, nb = n * 32;
  const uint32_t bf = K->bf, A = K->n_advice, L = K->n_lookups;
  const size_t usable = n - (bf + 1);
  size_t rpos = 0;                                    // cursor in `randoms`
  auto draw_to = [&](char* dst, size_t count) -> int {
    ZK_CUDA(ctx, cudaMemcpyAsync(dst, randoms + rpos, count * 32, cudaMemcpyDefault, ctx->stream));
    rpos += count;
    return ZK_OK;
  };
  TranscriptWrite tr(K->transcript_kind);
  // ZK_PROVER_TRACE=1: host wall-clock at the section boundaries (after a stream sync), to find host-bound gaps
  const bool trace = getenv("ZK_PROVER_TRACE") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(ctx->stream);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    fprintf(stderr, "[zk_create_proof] %-28s %8.3f ms\n", what, ms);
  };
  auto write_points = [&](const std::vector<HostPoint>& pts) -> int {
    for (auto& p : pts)
      if (!tr.write_point(p.x, p.y, p.identity)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: cannot write points at infinity to the transcript");
    return ZK_OK;
  };
  zk_domain* dom = K->dom;
  const uint4* tw = (const uint4*)zk_domain_twiddles(dom);
  const uint32_t tw_shift = zk_domain_extended_k(dom) - K->k;
  const size_t half_ext = K->ext_n >> 1;
  // Work for the second stream: `body(domain)` is enqueued there after everything queued so far on the main stream.
  // With per-kernel timing on (the breakdown pass of bench.py) the two streams take turns so that kernel times do
  // not overlap; with ZK_PROVER_STREAMS=1 there is no second stream and the work runs in line.
  zk_ctx* cB = K->ctxB;
  const bool serialize = ctx->timing;
  if (cB) cB->timing = ctx->timing;
  auto side = [&](int ev, auto&& fn) -> int {
    if (!cB) return fn(dom, ctx);
    if (serialize) ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZK_CUDA(ctx, cudaEventRecord(K->ev_in[ev], ctx->stream));
    ZK_CUDA(ctx, cudaStreamWaitEvent(cB->stream, K->ev_in[ev], 0));
    int r = fn(K->domB, cB);
    if (r != ZK_OK) { ctx->err = cB->err; return r; }
    if (serialize) ZK_CUDA(ctx, cudaStreamSynchronize(cB->stream));
    return ZK_OK;
  };
  const uint32_t I_ = K->n_instance;
  auto slot_adv = [&](uint32_t i) { return (zk_fr*)zk_pk_witness_slot(K->pk, i); };
  auto slot_ins = [&](uint32_t i) { return (zk_fr*)zk_pk_witness_slot(K->pk, (size_t)K->n_advice + i); };
  auto slot_z = [&](uint32_t s_) { return (zk_fr*)zk_pk_witness_slot(K->pk, (size_t)K->n_advice + I_ + s_); };
  auto slot_lk = [&](uint32_t l, uint32_t t) { return (zk_fr*)zk_pk_witness_slot(K->pk, (size_t)K->n_advice + I_ + K->n_sets + 3 * (size_t)l + t); };

  // vk.hash_into; instances (KZG: the values are absorbed as scalars)
  tr.common_scalar(K->transcript_repr);
  for (uint32_t i = 0; i < K->n_instance; i++) {
    if (instance_lens[i] > usable) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: InstanceTooLarge");
    ZK_CUDA(ctx, cudaMemsetAsync(K->d_instance + i * nb, 0, nb, ctx->stream));
    if (instance_lens[i]) {
      if (!instance_values[i]) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: null instance column");
      ZK_CUDA(ctx, cudaMemcpyAsync(K->d_instance + i * nb, instance_values[i], (size_t)instance_lens[i] * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (uint32_t j = 0; j < instance_lens[i]; j++) tr.common_scalar(f_from_abi<Fr>(&instance_values[i][j]));
  }
  if (K->n_instance) {
    ZK_CUDA(ctx, cudaMemcpyAsync(K->d_instance_polys, K->d_instance, K->n_instance * nb, cudaMemcpyDeviceToDevice, ctx->stream));
    ZK_TRY(zk_lagrange_to_coeff_dev(dom, (zk_fr*)K->d_instance_polys, K->n_instance));
    ZK_TRY(zk_coeff_to_extended_dev(dom, (const zk_fr*)K->d_instance_polys, slot_ins(0), K->n_instance));
  }

  // ---- advice: upload, blind, commit ---------------------------------------------------------------------------------
  // Host-resident columns are uploaded in groups on the copy stream; the commitments (main stream) and the
  // coefficient forms / cosets (second stream) of a group start as soon as that group has arrived.
  for (uint32_t i = 0; i < A; i++)
    if (!advice_values[i]) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: null advice column");
  {
    const bool split = A >= 8 && n >= ((size_t)1 << 14) && !is_device_pointer(advice_values[0]) && !getenv("ZK_PROVER_NO_SPLIT_UPLOAD");
    const uint32_t G = split ? 4 : 1;
    cudaStream_t up = ctx->stream;
    if (split) { ZK_TRY(zk_copy_stream(ctx)); up = ctx->copy_stream; }
    const size_t r_adv = rpos;
    rpos += (size_t)A * (bf + 1) + A;                    // blinding rows of every column, then one (unused) Blind per column
    auto gb = [&](uint32_t g) { return (uint32_t)((uint64_t)A * g / G); };
    for (uint32_t g = 0; g < G; g++) {
      for (uint32_t i = gb(g); i < gb(g + 1); i++) {
        ZK_CUDA(ctx, cudaMemcpyAsync(K->d_advice + i * nb, advice_values[i], usable * 32, cudaMemcpyDefault, up));
        ZK_CUDA(ctx, cudaMemcpyAsync(K->d_advice + i * nb + usable * 32, randoms + r_adv + (size_t)i * (bf + 1), (size_t)(bf + 1) * 32,
                                     cudaMemcpyDefault, up));
      }
      if (split) ZK_CUDA(ctx, cudaEventRecord(K->ev_copy[g], up));
    }
    for (uint32_t g = 0; g < G; g++) {
      const uint32_t i0 = gb(g), cnt = gb(g + 1) - gb(g);
      if (!cnt) continue;
      if (split) {
        ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, K->ev_copy[g], 0));
        if (serialize) ZK_CUDA(ctx, cudaStreamSynchronize(up));
      }
      ZK_TRY(side(0, [&](zk_domain* d, zk_ctx* on) -> int {   // advice_polys and their cosets, concurrently with the commitments
        return transform_columns(K, d, on, K->d_advice + i0 * nb, K->d_advice_polys + i0 * nb, cnt, [&](size_t i) { return slot_adv(i0 + (uint32_t)i); });
      }));
      ZK_TRY(commit_enqueue(K, K->d_advice + i0 * nb, cnt, ZK_BASES_G_LAGRANGE, i0));
    }
    std::vector<HostPoint> pts;
    if (A) ZK_TRY(commit_fetch(K, A, &pts));
    ZK_TRY(write_points(pts));
  }
  std::vector<Fr> challenges(K->n_challenges);
  for (auto& c : challenges) c = tr.squeeze_challenge();
  if (K->n_challenges) {
    std::vector<zk_fr> ch(K->n_challenges);
    for (uint32_t i = 0; i < K->n_challenges; i++) f_to_abi(challenges[i], &ch[i]);
    ZK_CUDA(ctx, cudaMemcpyAsync(K->d_challenges, ch.data(), ch.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  Fr theta = tr.squeeze_challenge();
  mark("advice committed");

  // ---- lookups: compress, permute, commit A' and S' ---------------------------------------------------------------------
  // d_lookup holds five planes of L columns: 0 compressed inputs, 1 compressed tables, 2 A', 3 S', 4 Z -- so that the
  // commitments, transforms and sorts of all lookup arguments are single batched calls.
  auto lk = [&](uint32_t l, int which) { return K->d_lookup + ((size_t)which * L + l) * nb; };
  auto lkp = [&](uint32_t l, int which) { return K->d_lookup_polys + ((size_t)which * L + l) * nb; };   // 0 A', 1 S', 2 Z (coefficients)
  const uint32_t gworld = (K->group && K->d_stage_p) ? zk_group_world(K->group) : 1, grank = gworld > 1 ? zk_group_rank(K->group) : 0;
  const uint32_t l_per = L ? (L + gworld - 1) / gworld : 1;
  auto lookup_owner = [&](uint32_t l) { return gworld > 1 ? l / l_per : 0u; };
  const uint32_t l_lo = std::min(L, grank * l_per), l_hi = std::min(L, (grank + 1) * l_per);
  const bool share_lookup_columns = gworld > 1 && L && !getenv("ZK_GROUP_NO_LOOKUP_SHARE");
  if (L) {
    std::vector<const void*> cols(K->CL.n_slots);
    for (uint32_t i = 0; i < K->n_fixed; i++) cols[K->CL.prog.off_fixed + i] = K->d_fixed_values + i * nb;
    for (uint32_t i = 0; i < A; i++) cols[K->CL.prog.off_advice + i] = K->d_advice + i * nb;
    for (uint32_t i = 0; i < K->n_instance; i++) cols[K->CL.prog.off_instance + i] = K->d_instance + i * nb;
    ZK_CUDA(ctx, cudaMemcpyAsync((void*)K->d_lcols, cols.data(), cols.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    EvalScalars S;
    S.beta = Fr::zero(); S.gamma = Fr::zero(); S.y = Fr::zero(); S.theta = theta;
    S.delta_start = Fr::zero(); S.delta = Fr::zero(); S.one = Fr::one();
    ZK_CUDA(ctx, cudaMemsetAsync(K->d_flags, 0, 64, ctx->stream));
    // in a group every lookup argument has an owner rank that does all of its work: compression, sort / permutation,
    // the commitments of A', S' and later Z, their transforms and the grand product
    for (uint32_t l = l_lo; l < l_hi; l++) {
      ZK_TRY(eval_rows(K, 2 * l, S, lk(l, 0)));
      ZK_TRY(eval_rows(K, 2 * l + 1, S, lk(l, 1)));
    }
    ZK_TRY(permute_expression_pairs(K, usable, l_lo, l_hi - l_lo));
    for (uint32_t l = 0; l < L; l++) {                 // RNG order: A' rows, S' rows, then the two Blinds, per lookup
      if (l >= l_lo && l < l_hi) {
        ZK_TRY(draw_to(lk(l, 2) + usable * 32, bf + 1));
        ZK_TRY(draw_to(lk(l, 3) + usable * 32, bf + 1));
      } else {
        rpos += 2 * (size_t)(bf + 1);
      }
      rpos += 2;
    }
    ZK_TRY(side(1, [&](zk_domain* d, zk_ctx* on) -> int {   // A', S' -> coefficient forms and cosets
      std::vector<ColJob> jobs;
      for (uint32_t l = 0; l < L; l++) jobs.push_back({lk(l, 2), lkp(l, 0), slot_lk(l, 1), lookup_owner(l)});
      for (uint32_t l = 0; l < L; l++) jobs.push_back({lk(l, 3), lkp(l, 1), slot_lk(l, 2), lookup_owner(l)});
      return transform_jobs(K, d, on, jobs);
    }));
    // status of the permutations: every rank must learn about a failure anywhere (nobody may wait in a collective)
    uint32_t flags = 0;
    ZK_CUDA(ctx, cudaMemcpyAsync(&flags, K->d_flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (gworld > 1) {
      uint32_t* xf = (uint32_t*)K->d_xchg;
      std::vector<uint32_t> mine(4, 0), all(4 * (size_t)gworld);
      mine[0] = flags;
      ZK_CUDA(ctx, cudaMemcpyAsync(xf + 4 * grank, mine.data(), 16, cudaMemcpyHostToDevice, ctx->stream));
      ZK_TRY(zk_group_allgather(K->group, ctx, xf, 16));
      ZK_CUDA(ctx, cudaMemcpyAsync(all.data(), xf, all.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
      ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      for (uint32_t r = 0; r < gworld; r++) flags |= all[4 * r];
    }
    if (flags & 6u) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: ConstraintSystemFailure (a lookup input is not in its table)");
    std::vector<HostPoint> pts;
    if (share_lookup_columns) {
      // A lookup's permuted columns exist only on its owner, and owners hold 1 or 2 lookups each (11 lookups on 8 GPUs): their
      // commitments -- the most expensive columns of the proof -- would follow that imbalance.  Every owner copies its columns
      // into all peers' lookup areas (16 MB each over NVLink, one flag round), and the 2 L columns are then committed as a
      // replicated batch with the cost-balanced split.
      std::vector<char*> cols; std::vector<uint32_t> owners;
      for (int plane = 2; plane <= 3; plane++)
        for (uint32_t l = 0; l < L; l++) { cols.push_back(lk(l, plane)); owners.push_back(lookup_owner(l)); }
      ZK_TRY(zk_group_exchange_rows(K->group, ctx, cols.data(), owners.data(), cols.size(), K->n, K->n));
      ZK_TRY(commit(K, lk(0, 2), 2 * (size_t)L, ZK_BASES_G_LAGRANGE, &pts));
    } else {
      std::vector<CommitJob> cj;
      for (uint32_t l = 0; l < L; l++) cj.push_back({lk(l, 2), lookup_owner(l)});
      for (uint32_t l = 0; l < L; l++) cj.push_back({lk(l, 3), lookup_owner(l)});
      ZK_TRY(commit_jobs(K, cj, ZK_BASES_G_LAGRANGE, &pts));
    }
    for (uint32_t l = 0; l < L; l++) ZK_TRY(write_points({pts[l], pts[L + l]}));
  }
  Fr beta = tr.squeeze_challenge();
  mark("lookups permuted");
  Fr gamma = tr.squeeze_challenge();

  // ---- permutation.commit --------------------------------------------------------------------------------------------------
  auto column_values = [&](const zk_column& c) -> const char* {
    return c.column_type == ZK_COL_ADVICE ? K->d_advice + c.index * nb : c.column_type == ZK_COL_FIXED ? K->d_fixed_values + c.index * nb
                                                                                                      : K->d_instance + c.index * nb;
  };
  // single GPU: the permutation and lookup products are committed together (17 columns for the QE3 shape) -- one call's
  // fixed latency less; a group keeps its two differently sharded commitments
  const bool merged_products = K->n_sets && L && (!(K->group && zk_group_world(K->group) > 1) || share_lookup_columns) && !getenv("ZK_PROVER_SPLIT_PRODUCTS");
  if (K->n_sets) {
    // every column set at once: z_s = seed_s * (exclusive prefix products of num/den), seed_{s+1} = z_s[usable]
    const uint32_t NS = K->n_sets;
    Fr delta = fr_delta(), dpow = Fr::one();
    std::vector<PermCols> pc(NS);
    std::vector<PermNum> pn(NS);
    for (uint32_t s = 0; s < NS; s++) {
      uint32_t c0 = s * K->chunk, c1 = std::min(c0 + K->chunk, K->n_perm);
      memset(&pc[s], 0, sizeof(PermCols));
      pc[s].ncols = c1 - c0;
      for (uint32_t c = c0; c < c1; c++) {
        pc[s].values[c - c0] = (const uint4*)column_values(K->perm_columns[c]);
        pc[s].sigma[c - c0] = (const uint4*)(K->d_sigma_values + c * nb);
        pn[s].dbeta[c - c0] = dpow * beta;
        dpow = dpow * delta;
      }
      for (uint32_t c = pc[s].ncols; c < 8; c++) pn[s].dbeta[c] = Fr::zero();
    }
    ZK_CUDA(ctx, cudaMemcpyAsync(K->d_permcols, pc.data(), NS * sizeof(PermCols), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(K->d_permnum, pn.data(), NS * sizeof(PermNum), cudaMemcpyHostToDevice, ctx->stream));
    dim3 g(blocks_for(n, 256), NS);
    ZK_LAUNCH(ctx, "permutation_product", k_perm_den, g, 256, 0, (const PermCols*)K->d_permcols, beta, gamma, (uint4*)K->d_w0, n);
    ZK_TRY(batch_invert(K, K->d_w0, K->d_w1, (size_t)NS * n));
    ZK_LAUNCH(ctx, "permutation_product", k_perm_num, g, 256, 0, (const PermCols*)K->d_permcols, (const PermNum*)K->d_permnum, gamma, tw,
              tw_shift, half_ext, (uint4*)K->d_w0, n);
    ZK_TRY(run_scan<FrMulMonoid>(K, "permutation_product", K->d_w0, K->d_perm_z, n, 0, 1, nullptr, NS, n, n));
    ZK_LAUNCH(ctx, "permutation_product", k_perm_seeds, 1, 32, 0, (const uint4*)K->d_perm_z, n, usable, NS, (uint4*)K->d_seeds);
    ZK_LAUNCH(ctx, "permutation_product", k_scale_seg, g, 256, 0, (uint4*)K->d_perm_z, (const uint4*)K->d_seeds, n);
    for (uint32_t s = 0; s < NS; s++) {
      ZK_TRY(draw_to(K->d_perm_z + s * nb + (n - bf) * 32, bf));
      rpos += 1;
    }
    ZK_TRY(side(2, [&](zk_domain* d, zk_ctx* on) -> int {
      return transform_columns(K, d, on, K->d_perm_z, K->d_perm_z_polys, NS, [&](size_t i) { return slot_z((uint32_t)i); });
    }));
    if (!merged_products) {
      std::vector<HostPoint> pts;
      ZK_TRY(commit(K, K->d_perm_z, NS, ZK_BASES_G_LAGRANGE, &pts));
      ZK_TRY(write_points(pts));
    }
  }

  // ---- lookups: commit_product (own lookups only in a group) -------------------------------------------------------------------
  if (L) {
    const uint32_t Lloc = l_hi - l_lo;
    if (Lloc) {
      dim3 g(blocks_for(n, 256), Lloc);
      ZK_LAUNCH(ctx, "lookup_product", k_lookup_den, g, 256, 0, (const uint4*)lk(l_lo, 2), (const uint4*)lk(l_lo, 3), beta, gamma, (uint4*)K->d_w0, n);
      ZK_TRY(batch_invert(K, K->d_w0, K->d_w1, (size_t)Lloc * n));
      ZK_LAUNCH(ctx, "lookup_product", k_lookup_num, g, 256, 0, (const uint4*)lk(l_lo, 0), (const uint4*)lk(l_lo, 1), beta, gamma, (uint4*)K->d_w0, n);
      ZK_TRY(run_scan<FrMulMonoid>(K, "lookup_product", K->d_w0, lk(l_lo, 4), n, 0, 1, nullptr, Lloc, n, n));
    }
    for (uint32_t l = 0; l < L; l++) {
      if (l >= l_lo && l < l_hi) ZK_TRY(draw_to(lk(l, 4) + (n - bf) * 32, bf)); else rpos += bf;
      rpos += 1;
    }
    ZK_TRY(side(3, [&](zk_domain* d, zk_ctx* on) -> int {
      std::vector<ColJob> jobs;
      for (uint32_t l = 0; l < L; l++) jobs.push_back({lk(l, 4), lkp(l, 2), slot_lk(l, 0), lookup_owner(l)});
      return transform_jobs(K, d, on, jobs);
    }));
    if (merged_products) {
      // one batch of L + n_sets columns; the transcript takes the permutation points first
      if (share_lookup_columns) {      // the owners' Z columns go to every rank first (see commit_permuted above)
        std::vector<char*> cols; std::vector<uint32_t> owners;
        for (uint32_t l = 0; l < L; l++) { cols.push_back(lk(l, 4)); owners.push_back(lookup_owner(l)); }
        ZK_TRY(zk_group_exchange_rows(K->group, ctx, cols.data(), owners.data(), cols.size(), K->n, K->n));
      }
      std::vector<HostPoint> pts;
      ZK_TRY(commit(K, lk(0, 4), L + K->n_sets, ZK_BASES_G_LAGRANGE, &pts));
      std::vector<HostPoint> perm(pts.begin() + L, pts.end()), look(pts.begin(), pts.begin() + L);
      ZK_TRY(write_points(perm));
      ZK_TRY(write_points(look));
    } else {
      std::vector<CommitJob> cj;
      for (uint32_t l = 0; l < L; l++) cj.push_back({lk(l, 4), lookup_owner(l)});
      std::vector<HostPoint> pts;
      ZK_TRY(commit_jobs(K, cj, ZK_BASES_G_LAGRANGE, &pts));
      ZK_TRY(write_points(pts));
    }
  }

  // ---- vanishing.commit: random polynomial ------------------------------------------------------------------------------------
  ZK_TRY(draw_to(K->d_random, n));
  rpos += 1;
  {
    std::vector<HostPoint> pts;
    ZK_TRY(commit(K, K->d_random, 1, ZK_BASES_G, &pts));
    ZK_TRY(write_points(pts));
  }
  Fr y = tr.squeeze_challenge();
  mark("products committed");

  // ---- h(X): every coset is in place once the second stream has drained ----------------------------------------------
  {
    if (cB) {
      ZK_CUDA(ctx, cudaEventRecord(K->ev_done, cB->stream));
      ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, K->ev_done, 0));
    }
    std::vector<zk_fr> ch(K->n_challenges + 1);
    for (uint32_t i = 0; i < K->n_challenges; i++) f_to_abi(challenges[i], &ch[i]);
    zk_fr yh, bh, gh, th;
    f_to_abi(y, &yh); f_to_abi(beta, &bh); f_to_abi(gamma, &gh); f_to_abi(theta, &th);
    const uint32_t world = K->group ? zk_group_world(K->group) : 1;
    if (world > 1 && (K->ext_n % world) == 0) {          // every GPU evaluates its slice of the extended rows
      const size_t rows = K->ext_n / world, r0 = rows * zk_group_rank(K->group);
      ZK_TRY(zk_evaluate_h_launch(K->pk, ch.data(), &yh, &bh, &gh, &th, (zk_fr*)K->d_hext, r0, r0 + rows));
      ZK_TRY(zk_group_allgather(K->group, ctx, K->d_hext, rows * 32));
    } else {
      ZK_TRY(zk_evaluate_h_launch(K->pk, ch.data(), &yh, &bh, &gh, &th, (zk_fr*)K->d_hext, 0, K->ext_n));
    }
  }
  // vanishing.construct
  ZK_TRY(zk_divide_by_vanishing_poly_dev(dom, (zk_fr*)K->d_hext));
  ZK_TRY(zk_extended_to_coeff_dev(dom, (zk_fr*)K->d_hext, (zk_fr*)K->d_hpieces));
  const uint32_t n_pieces = K->degree - 1;
  rpos += n_pieces;
  {
    std::vector<HostPoint> pts;
    ZK_TRY(commit(K, K->d_hpieces, n_pieces, ZK_BASES_G, &pts));
    ZK_TRY(write_points(pts));
  }
  Fr x = tr.squeeze_challenge();
  mark("h committed");
  Fr xn = Fr::pow_u64(x, (uint64_t)n);

  // ---- opening points and their power tables -----------------------------------------------------------------------------------
  std::vector<int32_t> rot_of;                        // point index -> rotation
  std::vector<Fr> pt_val;
  // Points are keyed by their field VALUE, as halo2's construct_intermediate_sets (BTreeSet of points) and the verifier do:
  // two rotations that are congruent mod n (small circuits: -(bf + 1) or a query rotation wrapping onto another) are one
  // opening point.
  std::vector<int32_t> rot_seen; std::vector<uint32_t> rot_point;      // rotation -> point index (cache)
  auto point_of = [&](int32_t r) -> uint32_t {
    for (uint32_t i = 0; i < rot_seen.size(); i++) if (rot_seen[i] == r) return rot_point[i];
    Fr v = x * (r >= 0 ? Fr::pow_u64(K->omega, (uint64_t)r) : Fr::pow_u64(K->omega_inv, (uint64_t)(-(int64_t)r)));
    uint32_t idx = (uint32_t)pt_val.size();
    for (uint32_t i = 0; i < pt_val.size(); i++) if (fr_same(pt_val[i], v)) { idx = i; break; }
    if (idx == pt_val.size()) { rot_of.push_back(r); pt_val.push_back(v); }
    rot_seen.push_back(r); rot_point.push_back(idx);
    return idx;
  };
  struct EvalReq { const void* poly; uint32_t point; };
  std::vector<EvalReq> reqs;
  auto req = [&](const void* poly, int32_t r) { reqs.push_back({poly, point_of(r)}); };
  for (auto& q : K->advice_queries) req(K->d_advice_polys + q.column * nb, q.rotation);
  for (auto& q : K->fixed_queries) req(K->d_fixed_polys + q.column * nb, q.rotation);
  // vanishing.evaluate: h(X) = sum_i xn^i h_i(X) (opened, not written); random_eval
  {
    std::vector<const void*> pp; std::vector<Fr> sc;
    Fr cur = Fr::one();
    for (uint32_t i = 0; i < n_pieces; i++) { pp.push_back(K->d_hpieces + i * nb); sc.push_back(cur); cur = cur * xn; }
    ZK_TRY(lincomb(K, pp, sc, K->d_hpoly));
  }
  req(K->d_random, 0);
  for (uint32_t i = 0; i < K->n_perm; i++) req(K->d_sigma_polys + i * nb, 0);
  const int32_t last_rot = -(int32_t)(bf + 1);
  for (uint32_t s = 0; s < K->n_sets; s++) {
    req(K->d_perm_z_polys + s * nb, 0); req(K->d_perm_z_polys + s * nb, 1);
    if (s + 1 < K->n_sets) req(K->d_perm_z_polys + s * nb, last_rot);
  }
  for (uint32_t l = 0; l < L; l++) {
    req(lkp(l, 2), 0); req(lkp(l, 2), 1); req(lkp(l, 0), 0); req(lkp(l, 0), -1); req(lkp(l, 1), 0);
  }
  const size_t n_written = reqs.size();
  req(K->d_hpoly, 0);                                  // the prover's own evaluation of h (ProverQuery::get_eval)
  if (2 * (rot_of.size() + 1) > K->n_table_slots) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: opening point tables");
  for (uint32_t p = 0; p < rot_of.size(); p++) ZK_TRY(build_pow_table(K, pt_val[p], 2 * p));
  std::vector<Fr> evals(reqs.size());
  const uint32_t eworld = (K->group && !getenv("ZK_GROUP_NO_EVAL_SHARD")) ? zk_group_world(K->group) : 1;
  if (eworld > 1) {
    // a group shares the evaluations: request i belongs to rank i mod world (every rank holds every coefficient form); each
    // rank evaluates its requests point by point into its segment of the exchange area, one small exchange, then every rank
    // replays the same enumeration to put the values back in request order
    const uint32_t erank = zk_group_rank(K->group);
    auto mine_of = [&](uint32_t r, uint32_t p) {
      std::vector<size_t> idx;
      for (size_t i = 0; i < reqs.size(); i++) if (reqs[i].point == p && i % eworld == r) idx.push_back(i);
      return idx;
    };
    size_t cap = 0;
    for (uint32_t r = 0; r < eworld; r++) { size_t c = 0; for (uint32_t p = 0; p < rot_of.size(); p++) c += mine_of(r, p).size(); cap = std::max(cap, c); }
    if (cap > K->max_list || cap * 32 * eworld > (size_t)(K->max_list + 64) * 16 * sizeof(PtRaw)) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: evaluation exchange area");
    char* xe = K->d_xchg;
    size_t pos = 0;
    for (uint32_t p = 0; p < rot_of.size(); p++) {
      std::vector<size_t> idx = mine_of(erank, p);
      if (idx.empty()) continue;
      std::vector<const void*> polys;
      for (size_t i : idx) polys.push_back(reqs[i].poly);
      ZK_TRY(eval_list_to(K, polys, 2 * p, xe + ((size_t)erank * cap + pos) * 32));
      pos += idx.size();
    }
    ZK_TRY(zk_group_allgather(K->group, ctx, xe, cap * 32));
    std::vector<zk_fr> h(cap * eworld);
    ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), xe, h.size() * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t r = 0; r < eworld; r++) {
      size_t q = 0;
      for (uint32_t p = 0; p < rot_of.size(); p++)
        for (size_t i : mine_of(r, p)) evals[i] = f_from_abi<Fr>(&h[(size_t)r * cap + q++]);
    }
  } else
  for (uint32_t p = 0; p < rot_of.size(); p++) {
    std::vector<const void*> polys; std::vector<size_t> where;
    for (size_t i = 0; i < reqs.size(); i++) if (reqs[i].point == p) { polys.push_back(reqs[i].poly); where.push_back(i); }
    for (size_t off = 0; off < polys.size(); off += K->max_list) {
      size_t cnt = std::min<size_t>(K->max_list, polys.size() - off);
      std::vector<const void*> part(polys.begin() + off, polys.begin() + off + cnt);
      std::vector<Fr> got;
      ZK_TRY(eval_list(K, part, 2 * p, &got));
      for (size_t i = 0; i < cnt; i++) evals[where[off + i]] = got[i];
    }
  }
  for (size_t i = 0; i < n_written; i++) tr.write_scalar(evals[i]);
  mark("evaluations");

  // ---- queries in create_proof's order ---------------------------------------------------------------------------------------------
  std::vector<Query> queries;
  {
    size_t e = 0;
    std::vector<Fr> adv_e(evals.begin(), evals.begin() + K->advice_queries.size()); e += K->advice_queries.size();
    std::vector<Fr> fix_e(evals.begin() + e, evals.begin() + e + K->fixed_queries.size()); e += K->fixed_queries.size();
    Fr random_eval = evals[e++];
    std::vector<Fr> sig_e(evals.begin() + e, evals.begin() + e + K->n_perm); e += K->n_perm;
    struct PE { Fr x, next, last; };
    std::vector<PE> pe(K->n_sets);
    for (uint32_t s = 0; s < K->n_sets; s++) { pe[s].x = evals[e++]; pe[s].next = evals[e++]; if (s + 1 < K->n_sets) pe[s].last = evals[e++]; }
    std::vector<std::vector<Fr>> le(L);
    for (uint32_t l = 0; l < L; l++) { le[l].assign(evals.begin() + e, evals.begin() + e + 5); e += 5; }
    Fr h_eval = evals[e++];
    const uint32_t p0 = point_of(0);
    for (size_t i = 0; i < K->advice_queries.size(); i++)
      queries.push_back({K->d_advice_polys + K->advice_queries[i].column * nb, point_of(K->advice_queries[i].rotation), adv_e[i]});
    for (uint32_t s = 0; s < K->n_sets; s++) {
      queries.push_back({K->d_perm_z_polys + s * nb, p0, pe[s].x});
      queries.push_back({K->d_perm_z_polys + s * nb, point_of(1), pe[s].next});
    }
    for (uint32_t s = K->n_sets; s-- > 0;) if (s + 1 < K->n_sets) queries.push_back({K->d_perm_z_polys + s * nb, point_of(last_rot), pe[s].last});
    for (uint32_t l = 0; l < L; l++) {
      queries.push_back({lkp(l, 2), p0, le[l][0]}); queries.push_back({lkp(l, 0), p0, le[l][2]}); queries.push_back({lkp(l, 1), p0, le[l][4]});
      queries.push_back({lkp(l, 0), point_of(-1), le[l][3]}); queries.push_back({lkp(l, 2), point_of(1), le[l][1]});
    }
    for (size_t i = 0; i < K->fixed_queries.size(); i++)
      queries.push_back({K->d_fixed_polys + K->fixed_queries[i].column * nb, point_of(K->fixed_queries[i].rotation), fix_e[i]});
    for (uint32_t i = 0; i < K->n_perm; i++) queries.push_back({K->d_sigma_polys + i * nb, p0, sig_e[i]});
    queries.push_back({K->d_hpoly, p0, h_eval});
    queries.push_back({K->d_random, p0, random_eval});
  }

  // ---- ProverSHPLONK::create_proof --------------------------------------------------------------------------------------------------
  // construct_intermediate_sets: points are compared / ordered by their field value
  std::vector<uint32_t> pt_rank(pt_val.size());      // rank of point p in ascending field order
  {
    std::vector<uint32_t> ord(pt_val.size());
    for (uint32_t i = 0; i < ord.size(); i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return fr_less(pt_val[a], pt_val[b]); });
    for (uint32_t i = 0; i < ord.size(); i++) pt_rank[ord[i]] = i;
  }
  std::vector<std::pair<const void*, std::set<uint32_t>>> commitment_sets;          // (poly, ranks of its points)
  for (auto& q : queries) {
    bool found = false;
    for (auto& cs_ : commitment_sets) if (cs_.first == q.poly) { cs_.second.insert(pt_rank[q.point]); found = true; break; }
    if (!found) commitment_sets.push_back({q.poly, {pt_rank[q.point]}});
  }
  std::vector<uint32_t> by_rank(pt_val.size());
  for (uint32_t p = 0; p < pt_val.size(); p++) by_rank[pt_rank[p]] = p;
  std::vector<RotationSet> sets;
  {
    std::vector<std::set<uint32_t>> keys;
    for (auto& cs_ : commitment_sets) {
      size_t si = 0;
      for (; si < keys.size(); si++) if (keys[si] == cs_.second) break;
      if (si == keys.size()) {
        keys.push_back(cs_.second);
        RotationSet rs;
        for (uint32_t r : cs_.second) rs.points.push_back(by_rank[r]);               // ascending
        sets.push_back(rs);
      }
      std::vector<Fr> ev;
      for (uint32_t p : sets[si].points) {
        bool ok = false;
        for (auto& q : queries) if (q.poly == cs_.first && q.point == p) { ev.push_back(q.eval); ok = true; break; }
        if (!ok) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: missing query evaluation");
      }
      sets[si].polys.push_back(cs_.first);
      sets[si].evals.push_back(ev);
    }
  }
  mark("shplonk sets (host)");
  Fr ys = tr.squeeze_challenge();
  Fr v = tr.squeeze_challenge();
  while (K->d_setacc.size() < sets.size()) {
    void* p = nullptr;
    ZK_CUDA(ctx, cudaMalloc(&p, nb));
    K->d_setacc.push_back((char*)p);
  }
  // inverse power tables of the opening points (kate_division by suffix sums)
  {
    std::vector<Fr> inv_pts = pt_val;
    for (auto& p : inv_pts) if (p.is_zero()) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: opening point is zero");
    host_batch_invert(inv_pts);
    for (uint32_t p = 0; p < rot_of.size(); p++) ZK_TRY(build_pow_table(K, inv_pts[p], 2 * p + 1));
  }
  std::vector<std::vector<Fr>> ypow(sets.size());
  std::vector<std::vector<std::vector<Fr>>> rpoly(sets.size());
  {
    Fr vpow = Fr::one();
    for (size_t i = 0; i < sets.size(); i++) {
      RotationSet& rs = sets[i];
      std::vector<Fr> pts;
      for (uint32_t p : rs.points) pts.push_back(pt_val[p]);
      Fr cur = Fr::one();
      for (size_t j = 0; j < rs.polys.size(); j++) { ypow[i].push_back(cur); cur = cur * ys; }
      for (size_t off = 0; off < rs.polys.size(); off += K->max_list - 1) {         // sum_j y^j P_ij, in slices
        size_t cnt = std::min<size_t>(K->max_list - 1, rs.polys.size() - off);
        std::vector<const void*> pp(rs.polys.begin() + off, rs.polys.begin() + off + cnt);
        std::vector<Fr> sc(ypow[i].begin() + off, ypow[i].begin() + off + cnt);
        if (off) { pp.push_back(K->d_setacc[i]); sc.push_back(Fr::one()); }
        ZK_TRY(lincomb(K, pp, sc, K->d_setacc[i]));
      }
      const std::vector<std::vector<Fr>> basis = lagrange_basis(pts);
      LowPoly low; memset(&low, 0, sizeof low);
      low.m = (uint32_t)pts.size();
      if (low.m > 8) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: a polynomial is opened at more than 8 points");
      for (uint32_t t = 0; t < 8; t++) low.c[t] = Fr::zero();
      for (size_t j = 0; j < rs.polys.size(); j++) {
        rpoly[i].push_back(lagrange_interpolate(basis, rs.evals[j]));                // low_degree_equivalent
        for (uint32_t t = 0; t < low.m; t++) low.c[t] = low.c[t] + ypow[i][j] * rpoly[i][j][t];
      }
      ZK_CUDA(ctx, cudaMemcpyAsync(K->d_w0, K->d_setacc[i], nb, cudaMemcpyDeviceToDevice, ctx->stream));
      ZK_LAUNCH(ctx, "shplonk", k_sub_low, 1, 32, 0, (uint4*)K->d_w0, low);
      for (uint32_t p : rs.points) ZK_TRY(kate_divide(K, K->d_w0, 2 * p, 2 * p + 1));  // div_by_vanishing
      if (i == 0) {
        ZK_CUDA(ctx, cudaMemcpyAsync(K->d_hx, K->d_w0, nb, cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        ZK_TRY(lincomb(K, {K->d_hx, K->d_w0}, {Fr::one(), vpow}, K->d_hx));
      }
      vpow = vpow * v;
    }
  }
  {
    std::vector<HostPoint> pts;
    ZK_TRY(commit(K, K->d_hx, 1, ZK_BASES_G, &pts));
    ZK_TRY(write_points(pts));
  }
  Fr u = tr.squeeze_challenge();
  mark("shplonk h committed");
  {
    std::vector<const void*> pp; std::vector<Fr> sc;
    Fr vpow = Fr::one(), c0 = Fr::zero(), z0_diff = Fr::one();
    for (size_t i = 0; i < sets.size(); i++) {
      Fr z_i = Fr::one();                              // evaluate_vanishing_polynomial(super_point_set \ points, u)
      for (uint32_t p = 0; p < pt_val.size(); p++)
        if (std::find(sets[i].points.begin(), sets[i].points.end(), p) == sets[i].points.end()) z_i = z_i * (u - pt_val[p]);
      if (i == 0) z0_diff = z_i;
      Fr r_u = Fr::zero();
      for (size_t j = 0; j < sets[i].polys.size(); j++) r_u = r_u + ypow[i][j] * eval_small(rpoly[i][j], u);
      pp.push_back(K->d_setacc[i]); sc.push_back(vpow * z_i);
      c0 = c0 + vpow * z_i * r_u;
      vpow = vpow * v;
    }
    Fr zt = Fr::one();
    for (auto& p : pt_val) zt = zt * (u - p);
    pp.push_back(K->d_hx); sc.push_back(zt.neg());
    ZK_TRY(lincomb(K, pp, sc, K->d_lx));
    LowPoly low; memset(&low, 0, sizeof low);
    low.m = 1; low.c[0] = c0;
    for (uint32_t t = 1; t < 8; t++) low.c[t] = Fr::zero();
    ZK_LAUNCH(ctx, "shplonk", k_sub_low, 1, 32, 0, (uint4*)K->d_lx, low);
    if (u.is_zero() || z0_diff.is_zero()) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: degenerate challenge u");
    uint32_t su = 2 * (uint32_t)rot_of.size();
    ZK_TRY(build_pow_table(K, u, su));
    ZK_TRY(build_pow_table(K, Fr::inv(u), su + 1));
    ZK_TRY(kate_divide(K, K->d_lx, su, su + 1));
    ZK_LAUNCH(ctx, "shplonk", k_scale, blocks_for(n, 256), 256, 0, (uint4*)K->d_lx, Fr::inv(z0_diff), n);
    std::vector<HostPoint> pts;
    ZK_TRY(commit(K, K->d_lx, 1, ZK_BASES_G, &pts));
    ZK_TRY(write_points(pts));
  }
  mark("done");
  // a peer that never raised its flag leaves the wait kernels with a timeout and an error word: never hand out such a proof
  if (K->group && zk_group_world(K->group) > 1) {
    if (zk_group_p2p_error(K->group, ctx) || (cB && zk_group_p2p_error(K->group, cB)))
      ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: a peer-to-peer exchange of the group timed out");
  }
  if (rpos != zk_proof_randoms(K)) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_create_proof: drew %zu randoms, expected %zu", rpos, zk_proof_randoms(K));
  if (tr.proof.size() > proof_cap) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_create_proof: proof needs %zu bytes, buffer holds %zu", tr.proof.size(), proof_cap);
  memcpy(proof, tr.proof.data(), tr.proof.size());
  *proof_len = tr.proof.size();
  return ZK_OK;
}
