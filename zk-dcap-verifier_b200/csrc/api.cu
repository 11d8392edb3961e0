// api.cu -- context lifetime, timing statistics, device memory helpers and the integer-pipe
// micro-benchmarks used as roofline denominators.
#include "common.cuh"
#include "ec.cuh"

using namespace zk;

extern "C" const char* zk_version(void) { return "zkdcap_b200 0.1 (sm_100a)"; }

extern "C" int zk_ctx_create(int device, void* cuda_stream, zk_ctx** out) {
  if (!out) return ZK_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ZK_ERR_CUDA;
  zk_ctx* ctx = new zk_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return ZK_ERR_CUDA; }
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ZK_ERR_CUDA; }
    ctx->own_stream = true;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = ctx;
  return ZK_OK;
}

static void free_scratch(zk_scratch& s) { if (s.p) cudaFree(s.p); s.p = nullptr; s.cap = 0; }
static void free_srs(zk_srs_set& s) { if (!s.alias) { if (s.bases) cudaFree(s.bases); if (s.table) cudaFree(s.table); } s = zk_srs_set(); }

extern "C" void zk_ctx_destroy(zk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  free_scratch(ctx->msm_scalars); free_scratch(ctx->msm_bases); free_scratch(ctx->msm_work);
  free_scratch(ctx->msm_out); free_scratch(ctx->ntt_buf); free_scratch(ctx->ntt_buf2); free_scratch(ctx->misc);
  if (!ctx->srs_borrowed) { free_srs(ctx->srs_g); free_srs(ctx->srs_lagrange); }
  for (auto& t : ctx->twiddles) cudaFree(t.table);
  if (ctx->d_counters) cudaFree(ctx->d_counters);
  if (ctx->copy_event) cudaEventDestroy(ctx->copy_event);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// Lets `dst` (another stream on the same device) commit against the SRS that `src` holds resident.
extern "C" int zk_ctx_share_srs(zk_ctx* dst, const zk_ctx* src) {
  if (!dst || !src) return ZK_ERR_ARG;
  if (dst->device != src->device) ZK_FAIL(dst, ZK_ERR_ARG, "zk_ctx_share_srs: contexts are on different devices");
  if (!src->srs_g.bases) ZK_FAIL(dst, ZK_ERR_STATE, "zk_ctx_share_srs: source context holds no SRS");
  if (!dst->srs_borrowed) { free_srs(dst->srs_g); free_srs(dst->srs_lagrange); }
  dst->srs_g = src->srs_g; dst->srs_lagrange = src->srs_lagrange;
  dst->srs_k = src->srs_k; dst->srs_n = src->srs_n;
  dst->srs_borrowed = true;
  return ZK_OK;
}

extern "C" const char* zk_last_error(const zk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int zk_ctx_sync(zk_ctx* ctx) {
  if (!ctx) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}

extern "C" int zk_ctx_timing(zk_ctx* ctx, int enable) {
  if (!ctx) return ZK_ERR_ARG;
  ctx->timing = enable != 0;
  return ZK_OK;
}
static int drain_timing(zk_ctx* ctx) {
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& r : ctx->pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      zk_stat& s = ctx->stats[r.name];
      s.launches++; s.ms += ms;
    }
    ctx->event_pool.push_back(r.a); ctx->event_pool.push_back(r.b);
  }
  ctx->pending.clear();
  return ZK_OK;
}
extern "C" int zk_ctx_timing_reset(zk_ctx* ctx) {
  if (!ctx) return ZK_ERR_ARG;
  ZK_TRY(drain_timing(ctx));
  ctx->stats.clear();
  return ZK_OK;
}
extern "C" int zk_ctx_timing_read(zk_ctx* ctx, zk_kernel_stat* rows, int cap) {
  if (!ctx || (cap > 0 && !rows)) return ZK_ERR_ARG;
  ZK_TRY(drain_timing(ctx));
  int n = 0;
  for (auto& kv : ctx->stats) {
    if (n >= cap) break;
    memset(&rows[n], 0, sizeof rows[n]);
    strncpy(rows[n].name, kv.first.c_str(), sizeof(rows[n].name) - 1);
    rows[n].launches = kv.second.launches;
    rows[n].total_ms = kv.second.ms;
    n++;
  }
  return n;
}
extern "C" uint64_t zk_ctx_launch_count(const zk_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int zk_dev_alloc(zk_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
  return ZK_OK;
}
extern "C" int zk_dev_free(zk_ctx* ctx, void* dptr) {
  if (!ctx) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ZK_CUDA(ctx, cudaFree(dptr));
  return ZK_OK;
}
extern "C" int zk_h2d(zk_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx || (bytes && (!dst || !src))) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}
extern "C" int zk_d2h(zk_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx || (bytes && (!dst || !src))) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}

// ---- micro-benchmarks ------------------------------------------------------------------------

template <class F>
__global__ void __launch_bounds__(256) k_bench_modmul(uint32_t iters, uint32_t* sink) {
  // four independent chains per thread so the FMA pipe, not latency, is the limit
  F a = F::one(), b = F::r2(), c = F::one() + F::r2(), d = F::r2().dbl();
  a.l[0] ^= threadIdx.x; b.l[1] ^= blockIdx.x;
  for (uint32_t i = 0; i < iters; i++) {
    a = a * b; b = b * c; c = c * d; d = d * a;
  }
  F s = a + b + c + d;
  if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) sink[0] = s.l[3];
}
extern "C" int zk_bench_modmul(zk_ctx* ctx, int field, uint32_t iters, double* gmul_per_s) {
  if (!ctx || !gmul_per_s) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, 256));
  unsigned grid = ctx->sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  ZK_CUDA(ctx, cudaEventCreate(&e0)); ZK_CUDA(ctx, cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    if (rep == 1) ZK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (field == 0) { ZK_LAUNCH(ctx, "bench_modmul", k_bench_modmul<Fr>, grid, block, 0, iters, (uint32_t*)ctx->misc.p); }
    else { ZK_LAUNCH(ctx, "bench_modmul", k_bench_modmul<Fq>, grid, block, 0, iters, (uint32_t*)ctx->misc.p); }
  }
  ZK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  ZK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double muls = (double)grid * block * 4.0 * iters;
  *gmul_per_s = muls / (ms * 1e-3) / 1e9;
  return ZK_OK;
}

template <int KIND>
__global__ void __launch_bounds__(256) k_bench_imad(uint32_t iters, uint32_t* sink) {
  uint32_t x0 = threadIdx.x, x1 = blockIdx.x, x2 = 3, x3 = 5, x4 = 7, x5 = 11, x6 = 13, x7 = 17;
  uint32_t m = threadIdx.x * 2654435761u + 1;
  if (KIND == 0) {
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        x0 = x0 * m + x1; x1 = x1 * m + x2; x2 = x2 * m + x3; x3 = x3 * m + x4;
        x4 = x4 * m + x5; x5 = x5 * m + x6; x6 = x6 * m + x7; x7 = x7 * m + x0;
      }
    }
  } else {
    uint64_t y0 = x0, y1 = x1, y2 = x2, y3 = x3, y4 = x4, y5 = x5, y6 = x6, y7 = x7;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y0) : "r"((uint32_t)y1), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y1) : "r"((uint32_t)y2), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y2) : "r"((uint32_t)y3), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y3) : "r"((uint32_t)y4), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y4) : "r"((uint32_t)y5), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y5) : "r"((uint32_t)y6), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y6) : "r"((uint32_t)y7), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y7) : "r"((uint32_t)y0), "r"(m));
      }
    }
    x0 = (uint32_t)(y0 ^ (y0 >> 32)); x1 = (uint32_t)(y1 ^ (y1 >> 32)); x2 = (uint32_t)y2; x3 = (uint32_t)y3;
    x4 = (uint32_t)y4; x5 = (uint32_t)y5; x6 = (uint32_t)y6; x7 = (uint32_t)y7;
  }
  uint32_t s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  if (s == 0x13572468u) sink[0] = s;
}
extern "C" int zk_bench_imad(zk_ctx* ctx, int kind, uint32_t iters, double* timad_per_s) {
  if (!ctx || !timad_per_s) return ZK_ERR_ARG;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, 256));
  unsigned grid = ctx->sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  ZK_CUDA(ctx, cudaEventCreate(&e0)); ZK_CUDA(ctx, cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    if (rep == 1) ZK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (kind == 0) { ZK_LAUNCH(ctx, "bench_imad", k_bench_imad<0>, grid, block, 0, iters, (uint32_t*)ctx->misc.p); }
    else { ZK_LAUNCH(ctx, "bench_imad", k_bench_imad<1>, grid, block, 0, iters, (uint32_t*)ctx->misc.p); }
  }
  ZK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  ZK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double ops = (double)grid * block * 64.0 * iters;
  *timad_per_s = ops / (ms * 1e-3) / 1e12;
  return ZK_OK;
}

// ---- element-wise field / curve self-test kernels (parity tests call these through the ABI) ---
// op codes: see zk_field_op in include/zkdcap_b200.h
template <class F>
__global__ void k_field_op(int op, const uint4* a, const uint4* b, uint4* o, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = F::load(a + 2 * i), y = b ? F::load(b + 2 * i) : F::zero(), z;
  switch (op) {
    case 0: z = x + y; break;
    case 1: z = x - y; break;
    case 2: z = x * y; break;
    case 3: z = x.neg(); break;
    case 4: z = F::inv(x); break;
    case 5: z = x.sqr(); break;
    case 6: z = x.to_mont(); break;
    case 7: z = x.from_mont(); break;
    default: z = x.dbl(); break;
  }
  z.store(o + 2 * i);
}
extern "C" int zk_field_op(zk_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n) {
  if (!ctx || !a || !out) return ZK_ERR_ARG;
  if (!n) return ZK_OK;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  void *da = nullptr, *db = nullptr, *dout = nullptr;
  ZK_CUDA(ctx, cudaMalloc(&da, n * 32)); ZK_CUDA(ctx, cudaMalloc(&dout, n * 32));
  ZK_CUDA(ctx, cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (b) { ZK_CUDA(ctx, cudaMalloc(&db, n * 32)); ZK_CUDA(ctx, cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream)); }
  if (field == 0) { ZK_LAUNCH(ctx, "field_op", k_field_op<Fr>, zk_div_up(n, 128), 128, 0, op, (const uint4*)da, (const uint4*)db, (uint4*)dout, n); }
  else { ZK_LAUNCH(ctx, "field_op", k_field_op<Fq>, zk_div_up(n, 128), 128, 0, op, (const uint4*)da, (const uint4*)db, (uint4*)dout, n); }
  ZK_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return ZK_OK;
}
