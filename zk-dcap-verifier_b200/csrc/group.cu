// group.cu -- one proof over several GPUs of a box: a group of zk_ctx (one process per GPU) joined by NCCL.
//
// the independent column commitments of one create_proof shard across GPUs with no data-path
// collective other than an all-gather of the resulting 96-byte points; a single commitment shards by base range with
// an all-gather of one partial point per GPU; h(X) shards by row range with an all-gather of the rows.  The reference
// has no counterpart (it is a single-process rayon program); this is the build's own scaling axis for
// "DCAP proof wall-time at 1/2/4/8 B200" (BASELINE.json).
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy torch has already loaded is reused, so that one process
// never holds two NCCL instances); the library itself has no link-time dependency on it.
#include <dlfcn.h>

#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

// One peer-mapped exchange window per communicator (= per stream that issues exchanges): two halves used alternately
// (a rank may be one exchange ahead of a peer, never two -- it cannot finish exchange e + 1 before every peer has pushed
// e + 1, i.e. has consumed e), each half = world segments of kP2PSegment bytes, then one 128-byte flag line per source rank.
constexpr size_t kP2PSegment = 64 * 1024;
struct P2PWindow {
  char* local = nullptr;
  char* peer[kZkMaxGroup] = {};
  unsigned int* done = nullptr;
  uint32_t epoch = 0;
  bool ok = false;
  size_t half_bytes(uint32_t world) const { return world * kP2PSegment + world * 128; }
};

struct zk_group {
  zk_ctx* ctx = nullptr;
  uint32_t rank = 0, world = 1;
  void* comm = nullptr;      // collectives of the context's own stream
  void* comm2 = nullptr;     // collectives issued on another stream of the same device (the prover's transform stream)
  P2PWindow win[2];          // [0] exchanges on the context's stream, [1] on the second stream
  char* msm_xchg = nullptr;  // world XYZZ records: the partial points of zk_msm_sharded_dev
  // buffers with the same layout on every rank whose peers are mapped (CUDA IPC): rows are copied straight into them
  struct RegBuf { char* local; size_t bytes; char* peer[kZkMaxGroup]; };
  std::vector<RegBuf> regs;
  char* d_dummy = nullptr;   // 16 bytes pushed with the "rows have arrived" flag
};

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_comm_init)(void**, int, nccl_uid, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_err)(int);
typedef int (*fn_sendrecv)(void*, size_t, int, int, void*, cudaStream_t);     // ncclSend (const void*) / ncclRecv share this shape
typedef int (*fn_group)(void);

struct Nccl {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_comm_init comm_init = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_allgather allgather = nullptr;
  fn_err err = nullptr;
  fn_sendrecv send = nullptr, recv = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
};
Nccl* nccl() {
  static Nccl N;
  static bool tried = false;
  if (!tried) {
    tried = true;
    N.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (N.lib) {
      N.get_uid = (fn_get_uid)dlsym(N.lib, "ncclGetUniqueId");
      N.comm_init = (fn_comm_init)dlsym(N.lib, "ncclCommInitRank");
      N.comm_destroy = (fn_comm_destroy)dlsym(N.lib, "ncclCommDestroy");
      N.allgather = (fn_allgather)dlsym(N.lib, "ncclAllGather");
      N.err = (fn_err)dlsym(N.lib, "ncclGetErrorString");
      N.send = (fn_sendrecv)dlsym(N.lib, "ncclSend"); N.recv = (fn_sendrecv)dlsym(N.lib, "ncclRecv");
      N.group_start = (fn_group)dlsym(N.lib, "ncclGroupStart"); N.group_end = (fn_group)dlsym(N.lib, "ncclGroupEnd");
      if (!N.get_uid || !N.comm_init || !N.comm_destroy || !N.allgather) N.lib = nullptr;
    }
  }
  return N.lib ? &N : nullptr;
}

// ---- peer-to-peer exchange kernels -------------------------------------------------------------------------------------------
// spin until every peer's flag for this exchange shows `epoch` (bounded: ~4 s at 2 GHz, then *err = 1)
__device__ __forceinline__ bool p2p_wait_flags(const char* local_flags, uint32_t world, uint32_t rank, uint32_t epoch) {
  bool ok = true;
  if (threadIdx.x < world && threadIdx.x != rank) {
    const volatile uint32_t* f = (const volatile uint32_t*)(local_flags + 128 * threadIdx.x);
    const long long t0 = clock64();
    while (*f != epoch) {
      if (clock64() - t0 > 8000000000ll) { ok = false; break; }
    }
  }
  __threadfence_system();
  return ok;
}
// consumer: wait, then copy the peers' segments from the local window into d_buf (segment q at d_buf + q * bytes)
__global__ void __launch_bounds__(256) k_p2p_wait_gather(const char* half, const char* flags, uint32_t world, uint32_t rank, uint32_t epoch,
                                                         char* d_buf, size_t bytes, int* err) {
  if (!p2p_wait_flags(flags, world, rank, epoch)) *err = 1;
  __syncthreads();
  const size_t vec = bytes / 16;
  for (uint32_t q = 0; q < world; q++) {
    if (q == rank) continue;
    const uint4* src = (const uint4*)(half + q * kP2PSegment);
    uint4* dst = (uint4*)(d_buf + q * bytes);
    for (size_t i = threadIdx.x; i < vec; i += blockDim.x) dst[i] = __ldcg(src + i);
  }
}
// producer for data that is already in d_buf (generic small all-gather): push my segment, raise my flag everywhere
__global__ void __launch_bounds__(256) k_p2p_push(zk_p2p_push pp, const char* my_segment, size_t bytes) {
  const size_t vec = bytes / 16;
  const uint4* src = (const uint4*)my_segment;
  for (uint32_t q = 0; q < pp.world; q++) {
    if (q == pp.rank) continue;
    uint4* dst = (uint4*)pp.peer_data[q];
    for (size_t i = threadIdx.x; i < vec; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pp.world && threadIdx.x != pp.rank) *(volatile uint32_t*)pp.peer_flag[threadIdx.x] = pp.epoch;
}

int p2p_setup(zk_group* g, Nccl* N, int which) {
  zk_ctx* ctx = g->ctx;
  P2PWindow& W = g->win[which];
  const size_t total = 2 * W.half_bytes(g->world);
  void* comm = which == 0 ? g->comm : g->comm2;
  cudaIpcMemHandle_t* d_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> handles(g->world);
  if (cudaMalloc(&W.local, total) != cudaSuccess) { cudaGetLastError(); return ZK_ERR_OOM; }
  if (cudaMalloc(&W.done, 128) != cudaSuccess || cudaMalloc(&d_handles, g->world * sizeof(cudaIpcMemHandle_t)) != cudaSuccess) { cudaGetLastError(); return ZK_ERR_OOM; }
  cudaMemsetAsync(W.local, 0, total, ctx->stream);
  cudaMemsetAsync(W.done, 0, 128, ctx->stream);
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, W.local) != cudaSuccess) { cudaGetLastError(); cudaFree(d_handles); return ZK_ERR_CUDA; }
  cudaMemcpyAsync(d_handles + g->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream);
  int r = N->allgather(d_handles + g->rank, d_handles, sizeof mine, /*ncclChar*/ 0, comm, ctx->stream);
  if (r != 0 || cudaMemcpyAsync(handles.data(), d_handles, g->world * sizeof mine, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(d_handles); return ZK_ERR_CUDA; }
  cudaFree(d_handles);
  for (uint32_t q = 0; q < g->world; q++) {
    if (q == g->rank) { W.peer[q] = W.local; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, handles[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return ZK_ERR_CUDA; }
    W.peer[q] = (char*)p;
  }
  // nobody may push before every window is zeroed and mapped: one more (tiny) collective as a barrier
  int* d_tok = nullptr;
  if (cudaMalloc(&d_tok, g->world * 4) != cudaSuccess) return ZK_ERR_OOM;
  r = N->allgather(d_tok + g->rank, d_tok, 4, 0, comm, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_tok);
  if (r != 0) return ZK_ERR_CUDA;
  W.ok = true;
  return ZK_OK;
}
void p2p_teardown(zk_group* g) {
  for (auto& W : g->win) {
    for (uint32_t q = 0; q < g->world; q++)
      if (q != g->rank && W.peer[q]) cudaIpcCloseMemHandle(W.peer[q]);
    if (W.local) cudaFree(W.local);
    if (W.done) cudaFree(W.done);
    W = P2PWindow();
  }
}

}  // namespace

// Descriptor of the next exchange on `on`'s stream for a producer kernel (k_msm_final's epilogue, k_p2p_push); false when the
// peer windows are not available or a segment would not fit (the caller then uses the NCCL all-gather).
bool zk_group_p2p_begin(zk_group* g, zk_ctx* on, size_t bytes_per_rank, zk_p2p_push* out) {
  if (!g || g->world <= 1) return false;
  P2PWindow& W = g->win[(on && on != g->ctx) ? 1 : 0];
  if (!W.ok || bytes_per_rank > kP2PSegment || (bytes_per_rank & 15)) return false;
  W.epoch++;
  const size_t half = (W.epoch & 1) * W.half_bytes(g->world);
  zk_p2p_push pp;
  pp.world = g->world; pp.rank = g->rank; pp.epoch = W.epoch; pp.done = W.done;
  for (uint32_t q = 0; q < g->world; q++) {
    pp.peer_data[q] = W.peer[q] + half + g->rank * kP2PSegment;
    pp.peer_flag[q] = (uint32_t*)(W.peer[q] + half + g->world * kP2PSegment + 128 * g->rank);
  }
  *out = pp;
  return true;
}
// Completes the exchange begun by zk_group_p2p_begin: waits for every peer's flag and copies their segments into d_buf.
int zk_group_p2p_finish(zk_group* g, zk_ctx* on, void* d_buf, size_t bytes_per_rank) {
  zk_ctx* ctx = on ? on : g->ctx;
  P2PWindow& W = g->win[ctx != g->ctx ? 1 : 0];
  const size_t half = (W.epoch & 1) * W.half_bytes(g->world);
  int* err = (int*)(W.done + 16);
  ZK_LAUNCH(ctx, "group_p2p", k_p2p_wait_gather, 1, 256, 0, (const char*)(W.local + half), (const char*)(W.local + half + g->world * kP2PSegment),
            g->world, g->rank, W.epoch, (char*)d_buf, bytes_per_rank, err);
  return ZK_OK;
}
// Pushes a segment that is already complete in local memory (ranks without work in a sharded commitment, generic gathers).
int zk_group_p2p_push(zk_group* g, zk_ctx* on, const zk_p2p_push& pp, const void* my_segment, size_t bytes) {
  zk_ctx* ctx = on ? on : g->ctx;
  ZK_LAUNCH(ctx, "group_p2p", k_p2p_push, 1, 256, 0, pp, (const char*)my_segment, bytes);
  return ZK_OK;
}
// 1 when an exchange timed out since the last call (a peer never raised its flag); synchronises the stream.
int zk_group_p2p_error(zk_group* g, zk_ctx* on) {
  zk_ctx* ctx = on ? on : g->ctx;
  P2PWindow& W = g->win[ctx != g->ctx ? 1 : 0];
  if (!W.ok) return 0;
  int e = 0;
  cudaMemcpyAsync(&e, (int*)(W.done + 16), 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  return e;
}

extern "C" int zk_group_unique_id(uint8_t id[256]) {     // two NCCL ids: one communicator per stream
  Nccl* N = nccl();
  if (!N || !id) return ZK_ERR_STATE;
  for (int i = 0; i < 2; i++) {
    nccl_uid u;
    memset(&u, 0, sizeof u);
    if (N->get_uid(&u) != 0) return ZK_ERR_CUDA;
    memcpy(id + 128 * i, &u, 128);
  }
  return ZK_OK;
}

extern "C" int zk_group_create(zk_ctx* ctx, uint32_t rank, uint32_t world, const uint8_t id[256], zk_group** out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!out || !id || world == 0 || rank >= world) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_group_create: bad argument");
  if (world & (world - 1)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_group_create: the group size must be a power of two (got %u)", world);
  Nccl* N = nccl();
  if (!N) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_group_create: libnccl.so.2 not found");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  zk_group* g = new zk_group();
  g->ctx = ctx; g->rank = rank; g->world = world;
  for (int i = 0; i < 2; i++) {
    nccl_uid u;
    memcpy(&u, id + 128 * i, 128);
    int r = N->comm_init(i == 0 ? &g->comm : &g->comm2, (int)world, u, (int)rank);
    if (r != 0) {
      std::string e = N->err ? N->err(r) : "nccl error";
      if (g->comm) N->comm_destroy(g->comm);
      delete g;
      ZK_FAIL(ctx, ZK_ERR_CUDA, "zk_group_create: ncclCommInitRank: %s", e.c_str());
    }
  }
  // peer windows for the small exchanges (commitment points, status words); ZK_GROUP_NO_P2P=1 keeps NCCL all-gathers
  if (world > 1 && !getenv("ZK_GROUP_NO_P2P")) {
    for (int i = 0; i < 2; i++)
      if (p2p_setup(g, N, i) != ZK_OK) { g->win[i].ok = false; break; }
    // all ranks must agree (a rank that failed would deadlock the others): both windows or none, agreed through NCCL
    int* d_ok = nullptr;
    std::vector<int> oks(world, 0);
    int mine = (g->win[0].ok && g->win[1].ok) ? 1 : 0;
    if (cudaMalloc(&d_ok, world * 4) == cudaSuccess) {
      cudaMemcpyAsync(d_ok + rank, &mine, 4, cudaMemcpyHostToDevice, ctx->stream);
      N->allgather(d_ok + rank, d_ok, 4, 0, g->comm, ctx->stream);
      cudaMemcpyAsync(oks.data(), d_ok, world * 4, cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      cudaFree(d_ok);
    }
    bool all = true;
    for (int v : oks) all = all && v == 1;
    if (!all) { g->win[0].ok = g->win[1].ok = false; }
  }
  *out = g;
  return ZK_OK;
}
extern "C" int zk_group_p2p_enabled(const zk_group* g) { return g && g->win[0].ok && g->win[1].ok ? 1 : 0; }

int zk_xyzz_sum_normalised_dev(zk_ctx* ctx, const void* d_xyzz, size_t parts, zk_g1* d_out);      // msm.cu
int zk_group_allgather(zk_group* g, zk_ctx* on, void* d_buf, size_t bytes_per_rank);

// ONE multi-scalar multiplication over the GPUs of a group, sharded by contiguous base range: every rank
// keeps ITS slice of the bases resident (zk_srs_upload of the slice, with its own window table) and passes its slice of the
// scalars; the partial point of each rank is produced by the usual Pippenger pipeline, whose last kernel stores it straight
// into every peer's exchange window and raises a flag (no collective call); every rank then sums the `world` partial points
// and normalises.  d_out: one normalised G1 (device), identical on every rank.
extern "C" int zk_msm_sharded_dev(zk_group* g, const zk_fr* d_scalars, int which, size_t n_local, zk_g1* d_out) {
  if (!g) return ZK_ERR_ARG;
  zk_ctx* ctx = g->ctx;
  if (!d_out || (n_local && !d_scalars)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_msm_sharded_dev: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!g->msm_xchg) ZK_CUDA(ctx, cudaMalloc(&g->msm_xchg, (size_t)kZkMaxGroup * 128));
  char* mine = g->msm_xchg + (size_t)g->rank * 128;
  zk_p2p_push pp;
  const bool fused = g->world > 1 && zk_group_p2p_begin(g, ctx, 128, &pp);
  const bool old_raw = ctx->msm_raw;
  ctx->msm_raw = true;
  if (fused && n_local) ctx->p2p = &pp;
  int r = zk_msm_batch_dev(ctx, d_scalars, which, n_local, 1, (zk_g1*)mine);
  ctx->p2p = nullptr; ctx->msm_raw = old_raw;
  if (r != ZK_OK) return r;
  if (g->world > 1) {
    if (fused) {
      if (!n_local) ZK_TRY(zk_group_p2p_push(g, ctx, pp, mine, 128));       // the empty sum was written locally only
      ZK_TRY(zk_group_p2p_finish(g, ctx, g->msm_xchg, 128));
    } else {
      ZK_TRY(zk_group_allgather(g, ctx, g->msm_xchg, 128));
    }
  }
  return zk_xyzz_sum_normalised_dev(ctx, g->msm_xchg, g->world, d_out);
}

extern "C" void zk_group_destroy(zk_group* g) {
  if (!g) return;
  Nccl* N = nccl();
  cudaSetDevice(g->ctx->device); cudaDeviceSynchronize();
  if (g->msm_xchg) cudaFree(g->msm_xchg);
  if (g->d_dummy) cudaFree(g->d_dummy);
  for (auto& r : g->regs)
    for (uint32_t q = 0; q < g->world; q++) if (q != g->rank && r.peer[q]) cudaIpcCloseMemHandle(r.peer[q]);
  g->regs.clear();
  p2p_teardown(g);
  if (N && g->comm) {
    cudaSetDevice(g->ctx->device); cudaDeviceSynchronize();
    N->comm_destroy(g->comm);
    if (g->comm2) N->comm_destroy(g->comm2);
  }
  delete g;
}
extern "C" uint32_t zk_group_rank(const zk_group* g) { return g ? g->rank : 0; }
extern "C" uint32_t zk_group_world(const zk_group* g) { return g ? g->world : 1; }

// In-place all-gather on the stream of `on`: rank r's contribution sits at d_buf + r * bytes_per_rank.  `on` is the
// group's own context or another context of the same device (then the second communicator is used).
int zk_group_allgather(zk_group* g, zk_ctx* on, void* d_buf, size_t bytes_per_rank) {
  zk_ctx* ctx = on ? on : g->ctx;
  Nccl* N = nccl();
  if (!N) ZK_FAIL(ctx, ZK_ERR_STATE, "group: NCCL not loaded");
  void* comm = ctx == g->ctx ? g->comm : g->comm2;
  zk_p2p_push pp;
  if (zk_group_p2p_begin(g, ctx, bytes_per_rank, &pp)) {          // small: stores into the peers' windows + flags
    ZK_TRY(zk_group_p2p_push(g, ctx, pp, (const char*)d_buf + (size_t)g->rank * bytes_per_rank, bytes_per_rank));
    return zk_group_p2p_finish(g, ctx, d_buf, bytes_per_rank);
  }
  // large, inside a registered buffer (coefficient forms, h(X) rows): my segment is copied straight into every peer's copy of
  // the buffer (copy engines over NVLink), then the "arrived" flags go round
  for (auto& r : g->regs) {
    char* b = (char*)d_buf;
    if (b < r.local || b + (size_t)g->world * bytes_per_rank > r.local + r.bytes) continue;
    // one-sided writes need the receivers' consent: a flag round first ("my previous use of this buffer is over" -- it is
    // stream-ordered behind this rank's reads of the buffer), then the copies, then the "arrived" round
    if (!zk_group_p2p_begin(g, ctx, 16, &pp)) break;
    ZK_TRY(zk_group_p2p_push(g, ctx, pp, g->d_dummy + 16 * g->rank, 16));
    ZK_TRY(zk_group_p2p_finish(g, ctx, g->d_dummy, 16));
    if (!zk_group_p2p_begin(g, ctx, 16, &pp)) ZK_FAIL(ctx, ZK_ERR_STATE, "group: exchange window");
    zk_timing_begin(ctx, "group_allgather");
    char* mine = b + (size_t)g->rank * bytes_per_rank;
    cudaError_t ce = cudaSuccess;
    for (uint32_t q = 0; q < g->world && ce == cudaSuccess; q++)
      if (q != g->rank) ce = cudaMemcpyAsync(r.peer[q] + (mine - r.local), mine, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream);
    zk_timing_end(ctx);
    if (ce != cudaSuccess) ZK_FAIL(ctx, ZK_ERR_CUDA, "group: peer copy: %s", cudaGetErrorString(ce));
    ZK_TRY(zk_group_p2p_push(g, ctx, pp, g->d_dummy + 16 * g->rank, 16));
    return zk_group_p2p_finish(g, ctx, g->d_dummy, 16);
  }
  zk_timing_begin(ctx, "group_allgather");
  int r = N->allgather((const char*)d_buf + (size_t)g->rank * bytes_per_rank, d_buf, bytes_per_rank, /*ncclChar*/ 0, comm, ctx->stream);
  zk_timing_end(ctx);
  if (r != 0) ZK_FAIL(ctx, ZK_ERR_CUDA, "group: ncclAllGather: %s", N->err ? N->err(r) : "error");
  return ZK_OK;
}

// Collective: every rank registers ITS buffer of the same size and layout (the evaluate_h witness slots); afterwards
// zk_group_exchange_rows writes the peers' rows straight into their copies (device-to-device copies over NVLink through the
// IPC mapping) instead of going through ncclSend / ncclRecv.  No-op without peer windows.
int zk_group_register(zk_group* g, void* base, size_t bytes) {
  if (!g || g->world <= 1 || !g->win[0].ok || !g->win[1].ok || getenv("ZK_GROUP_NO_P2P_ROWS")) return ZK_OK;
  zk_ctx* ctx = g->ctx;
  for (auto& r : g->regs) if (r.local == (char*)base) return ZK_OK;
  Nccl* N = nccl();
  if (!N) return ZK_OK;
  zk_group::RegBuf rb; memset(&rb, 0, sizeof rb);
  rb.local = (char*)base; rb.bytes = bytes;
  cudaIpcMemHandle_t mine, *d_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> handles(g->world);
  int ok = cudaIpcGetMemHandle(&mine, base) == cudaSuccess ? 1 : 0;
  if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof mine); }
  // the handle and a "mapping worked" word travel together; the registration only counts when EVERY rank succeeded
  ZK_CUDA(ctx, cudaMalloc(&d_handles, g->world * sizeof mine));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_handles + g->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
  int r = N->allgather(d_handles + g->rank, d_handles, sizeof mine, 0, g->comm, ctx->stream);
  ZK_CUDA(ctx, cudaMemcpyAsync(handles.data(), d_handles, g->world * sizeof mine, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d_handles);
  if (r != 0) ok = 0;
  for (uint32_t q = 0; q < g->world && ok; q++) {
    if (q == g->rank) { rb.peer[q] = rb.local; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, handles[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    rb.peer[q] = (char*)p;
  }
  int* d_ok = nullptr;
  std::vector<int> oks(g->world, 0);
  ZK_CUDA(ctx, cudaMalloc(&d_ok, g->world * 4));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_ok + g->rank, &ok, 4, cudaMemcpyHostToDevice, ctx->stream));
  N->allgather(d_ok + g->rank, d_ok, 4, 0, g->comm, ctx->stream);
  ZK_CUDA(ctx, cudaMemcpyAsync(oks.data(), d_ok, g->world * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d_ok);
  bool all = true;
  for (int v : oks) all = all && v == 1;
  if (!all) {
    for (uint32_t q = 0; q < g->world; q++) if (q != g->rank && rb.peer[q]) cudaIpcCloseMemHandle(rb.peer[q]);
    return ZK_OK;
  }
  if (!g->d_dummy) { ZK_CUDA(ctx, cudaMalloc(&g->d_dummy, 256)); ZK_CUDA(ctx, cudaMemsetAsync(g->d_dummy, 0, 256, ctx->stream)); }
  g->regs.push_back(rb);
  return ZK_OK;
}

// Row-slice exchange of full-height columns.  Column c (c < count; cols[c] = its device address on THIS rank, the same
// layout on every rank) is owned by rank owners[c] and complete there; afterwards every rank q also holds rows
// [q * rows - halo, (q + 1) * rows + halo) (mod height) of every column -- what its slice of h(X) reads.  One grouped
// ncclSend / ncclRecv batch; nothing is staged.
int zk_group_exchange_rows(zk_group* g, zk_ctx* on, char* const* cols, const uint32_t* owners, size_t count, size_t height, size_t halo) {
  zk_ctx* ctx = on ? on : g->ctx;
  const size_t rows = height / g->world;
  // every column inside ONE registered buffer: the owner copies each peer's rows straight into the peer's copy of the buffer
  // (copy engines over NVLink), then raises the "rows have arrived" flag in every peer; everybody waits for all flags
  const zk_group::RegBuf* rb = nullptr;
  for (auto& r : g->regs) {
    bool inside = count > 0;
    for (size_t c = 0; c < count && inside; c++) inside = cols[c] >= r.local && cols[c] + height * 32 <= r.local + r.bytes;
    if (inside) { rb = &r; break; }
  }
  zk_p2p_push pp;
  if (rb && zk_group_p2p_begin(g, ctx, 16, &pp)) {
    zk_timing_begin(ctx, "group_exchange_rows");
    cudaError_t ce = cudaSuccess;
    for (size_t c = 0; c < count && ce == cudaSuccess; c++) {
      if (owners[c] != g->rank) continue;
      for (uint32_t q = 0; q < g->world && ce == cudaSuccess; q++) {
        if (q == g->rank) continue;
        size_t lo = (size_t)q * rows + height - halo, len = rows + 2 * halo;
        if (len >= height) { lo = 0; len = height; }
        lo %= height;
        size_t first = std::min(len, height - lo);
        for (int piece = 0; piece < 2 && ce == cudaSuccess; piece++) {
          size_t off = piece == 0 ? lo : 0, nrow = piece == 0 ? first : len - first;
          if (!nrow) continue;
          char* src = cols[c] + off * 32;
          ce = cudaMemcpyAsync(rb->peer[q] + (src - rb->local), src, nrow * 32, cudaMemcpyDeviceToDevice, ctx->stream);
        }
      }
    }
    zk_timing_end(ctx);
    if (ce != cudaSuccess) ZK_FAIL(ctx, ZK_ERR_CUDA, "group: peer copy of rows: %s", cudaGetErrorString(ce));
    ZK_TRY(zk_group_p2p_push(g, ctx, pp, g->d_dummy + 16 * g->rank, 16));
    return zk_group_p2p_finish(g, ctx, g->d_dummy, 16);
  }
  Nccl* N = nccl();
  if (!N || !N->send || !N->recv || !N->group_start || !N->group_end) ZK_FAIL(ctx, ZK_ERR_STATE, "group: NCCL point-to-point not available");
  void* comm = ctx == g->ctx ? g->comm : g->comm2;
  zk_timing_begin(ctx, "group_exchange_rows");
  int r = N->group_start();
  for (size_t c = 0; c < count && r == 0; c++) {
    const uint32_t owner = owners[c];
    for (uint32_t q = 0; q < g->world && r == 0; q++) {
      // the peer's rows as up to two contiguous pieces of the (cyclic) column
      if (owner != g->rank && q != g->rank) continue;
      if (owner == q) continue;                                  // the owner already has everything
      const uint32_t peer = owner == g->rank ? q : owner;        // I send to q, or I (= q) receive from the owner
      size_t lo = (size_t)q * rows + height - halo, len = rows + 2 * halo;      // start taken mod height below
      if (len >= height) { lo = 0; len = height; }
      lo %= height;
      size_t first = std::min(len, height - lo);
      for (int piece = 0; piece < 2 && r == 0; piece++) {
        size_t off = piece == 0 ? lo : 0, n = piece == 0 ? first : len - first;
        if (!n) continue;
        char* p = cols[c] + off * 32;
        r = owner == g->rank ? N->send(p, n * 32, /*ncclChar*/ 0, (int)peer, comm, ctx->stream)
                             : N->recv(p, n * 32, 0, (int)peer, comm, ctx->stream);
      }
    }
  }
  int r2 = N->group_end();
  zk_timing_end(ctx);
  if (r != 0 || r2 != 0) ZK_FAIL(ctx, ZK_ERR_CUDA, "group: row exchange: %s", N->err ? N->err(r ? r : r2) : "error");
  return ZK_OK;
}

// How `count` commitments are split over the group (also used by the CPU tests of the host logic):
// mode 0 = by column: rank r commits columns [lo, hi); mode 1 = by base range: every rank commits all `count`
// polynomials over bases [lo, hi) of n.
extern "C" void zk_group_shard_plan(uint32_t rank, uint32_t world, size_t count, size_t n, int* mode, size_t* per, size_t* lo, size_t* hi) {
  if (world <= 1) { *mode = 0; *per = count; *lo = 0; *hi = count; return; }
  if (count >= world) {
    size_t p = (count + world - 1) / world;
    *mode = 0; *per = p;
    *lo = std::min(count, (size_t)rank * p); *hi = std::min(count, (size_t)(rank + 1) * p);
  } else {
    size_t p = (n + world - 1) / world;
    *mode = 1; *per = p;
    *lo = std::min(n, (size_t)rank * p); *hi = std::min(n, (size_t)(rank + 1) * p);
  }
}
