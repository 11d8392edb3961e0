// common.cuh -- context object, error plumbing, kernel launch bookkeeping.
// One zk_ctx = one device + one stream (SURVEY.md 8b "Threading"); no global state.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/zkdcap_b200.h"

struct zk_srs_set {
  void* bases = nullptr;       // n affine points (64 B each)
  void* table = nullptr;       // optional window table: W x n affine points, table[w][i] = 2^(c*w) * bases[i]
  int table_c = 0, table_w = 0;
  bool alias = false;          // same memory as the other base set (g == g_lagrange was uploaded): never freed on its own
};

// Peer-to-peer exchange of small results inside a group of GPUs (group.cu): every rank maps every peer's exchange window
// (CUDA IPC over NVLink).  A producer kernel stores its results straight into the peers' windows and, once all of its CTAs
// are done, raises a flag there; the consumer kernel on each peer spins on its own flags.  Passed BY VALUE to kernels.
constexpr int kZkMaxGroup = 16;
struct zk_p2p_push {
  char* peer_data[kZkMaxGroup];        // peer q's window half for this exchange + my segment offset (nullptr for myself)
  uint32_t* peer_flag[kZkMaxGroup];    // peer q's flag slot for ME
  uint32_t world = 0, rank = 0, epoch = 0;
  unsigned int* done = nullptr;        // local counter of finished producer CTAs
};

struct zk_timing_rec { const char* name; cudaEvent_t a, b; };
struct zk_stat { uint64_t launches = 0; double ms = 0; };

struct zk_scratch {
  void* p = nullptr;
  size_t cap = 0;
};

struct zk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;   // H2D staging that overlaps compute inside one host-pointer call
  cudaEvent_t copy_event = nullptr;
  int sm_count = 148;
  std::string err;
  uint64_t launches = 0;
  // timing
  bool timing = false;
  std::vector<zk_timing_rec> pending;
  std::vector<cudaEvent_t> event_pool;
  std::map<std::string, zk_stat> stats;
  // SRS (ParamsKZG)
  uint32_t srs_k = 0;
  size_t srs_n = 0;
  zk_srs_set srs_g, srs_lagrange;
  bool srs_borrowed = false;    // the base sets belong to another context (zk_ctx_share_srs)
  // internal: MSM results are left as XYZZ records (x, y, zz, zzz; 128 B) instead of normalised G1 (96 B): the prover
  // normalises on the host (one shared inversion per batch) rather than spending a ~135 us single-thread Fermat
  // inversion on the device after every commitment batch
  bool msm_raw = false;
  // grow-only scratch buffers, reused across calls on this context
  zk_scratch msm_scalars, msm_bases, msm_work, msm_out, ntt_buf, ntt_buf2, misc;
  // cached twiddle tables: key = (log_n, first 8 bytes of omega...) -> device table of omega^e, e < n/2
  struct tw_entry { uint32_t log_n; zk_fr omega; void* table; };
  std::vector<tw_entry> twiddles;
  // device-side work counters (zk_ctx_counters): [0] bucket entries = non-zero signed digits sorted into buckets, i.e. the
  // group additions the accumulation performs, [1] MSMs
  unsigned long long* d_counters = nullptr;
  // set by the prover around a sharded commitment: k_msm_final then also stores its XYZZ results into the peers' windows
  const zk_p2p_push* p2p = nullptr;
};

#define ZK_FAIL(ctx, code, ...)                                  \
  do {                                                           \
    char _b[512];                                                \
    snprintf(_b, sizeof _b, __VA_ARGS__);                        \
    (ctx)->err = _b;                                             \
    return (code);                                               \
  } while (0)

#define ZK_CUDA(ctx, expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      ZK_FAIL(ctx, _e == cudaErrorMemoryAllocation ? ZK_ERR_OOM : ZK_ERR_CUDA, "%s:%d: %s: %s", \
              __FILE__, __LINE__, #expr, cudaGetErrorString(_e));                            \
  } while (0)

#define ZK_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != ZK_OK) return _r; \
  } while (0)

static inline int zk_scratch_reserve(zk_ctx* ctx, zk_scratch& s, size_t bytes) {
  if (s.cap >= bytes) return ZK_OK;
  if (s.p) {
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZK_CUDA(ctx, cudaFree(s.p));
    s.p = nullptr; s.cap = 0;
  }
  size_t cap = bytes + (bytes >> 3);
  ZK_CUDA(ctx, cudaMalloc(&s.p, cap));
  s.cap = cap;
  return ZK_OK;
}

static inline cudaEvent_t zk_get_event(zk_ctx* ctx) {
  if (!ctx->event_pool.empty()) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
static inline void zk_timing_begin(zk_ctx* ctx, const char* name) {
  ctx->launches++;
  if (!ctx->timing) return;
  zk_timing_rec r; r.name = name; r.a = zk_get_event(ctx); r.b = zk_get_event(ctx);
  cudaEventRecord(r.a, ctx->stream);
  ctx->pending.push_back(r);
}
static inline void zk_timing_end(zk_ctx* ctx) {
  if (!ctx->timing) return;
  cudaEventRecord(ctx->pending.back().b, ctx->stream);
}

// Launch `kernel` on the context stream; `name` is the kernel family used for statistics.
#define ZK_LAUNCH(ctx, name, kernel, grid, block, smem, ...)                         \
  do {                                                                               \
    zk_timing_begin(ctx, name);                                                      \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                 \
    zk_timing_end(ctx);                                                              \
    cudaError_t _le = cudaGetLastError();                                            \
    if (_le != cudaSuccess)                                                          \
      ZK_FAIL(ctx, ZK_ERR_CUDA, "%s:%d: launch %s: %s", __FILE__, __LINE__, name, cudaGetErrorString(_le)); \
  } while (0)

static inline int zk_copy_stream(zk_ctx* ctx) {
  if (ctx->copy_stream) return ZK_OK;
  ZK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  ZK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
  return ZK_OK;
}

static inline unsigned zk_div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// entry points implemented across translation units
int zk_twiddles_get(zk_ctx* ctx, const zk_fr* omega, uint32_t log_n, const void** table);
