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

#include <algorithm>
#include <string>

#include "common.cuh"

struct zk_group {
  zk_ctx* ctx = nullptr;
  uint32_t rank = 0, world = 1;
  void* comm = nullptr;      // collectives of the context's own stream
  void* comm2 = nullptr;     // collectives issued on another stream of the same device (the prover's transform stream)
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

}  // namespace

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
  *out = g;
  return ZK_OK;
}

extern "C" void zk_group_destroy(zk_group* g) {
  if (!g) return;
  Nccl* N = nccl();
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
  zk_timing_begin(ctx, "group_allgather");
  int r = N->allgather((const char*)d_buf + (size_t)g->rank * bytes_per_rank, d_buf, bytes_per_rank, /*ncclChar*/ 0, comm, ctx->stream);
  zk_timing_end(ctx);
  if (r != 0) ZK_FAIL(ctx, ZK_ERR_CUDA, "group: ncclAllGather: %s", N->err ? N->err(r) : "error");
  return ZK_OK;
}

// Row-slice exchange of full-height columns.  Column c (c < count; cols[c] = its device address on THIS rank, the same
// layout on every rank) is owned by rank owners[c] and complete there; afterwards every rank q also holds rows
// [q * rows - halo, (q + 1) * rows + halo) (mod height) of every column -- what its slice of h(X) reads.  One grouped
// ncclSend / ncclRecv batch; nothing is staged.
int zk_group_exchange_rows(zk_group* g, zk_ctx* on, char* const* cols, const uint32_t* owners, size_t count, size_t height, size_t halo) {
  zk_ctx* ctx = on ? on : g->ctx;
  Nccl* N = nccl();
  if (!N || !N->send || !N->recv || !N->group_start || !N->group_end) ZK_FAIL(ctx, ZK_ERR_STATE, "group: NCCL point-to-point not available");
  void* comm = ctx == g->ctx ? g->comm : g->comm2;
  const size_t rows = height / g->world;
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
