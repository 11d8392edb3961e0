// evalh.cu -- Evaluator::evaluate_h on the device: proving-key residency, the extended-domain
// transforms of the witness polynomials and the fused per-row kernel (evalh.cuh).
// Reference: halo2_proofs plonk::evaluation::Evaluator::evaluate_h ([UPSTREAM] src/plonk/evaluation.rs
// @ c254c75f), called from create_proof (circuits/src/sgx_dcap_verifier.rs:814).
#include <stdlib.h>

#include <dlfcn.h>

#include "common.cuh"
#include "evalh_codegen.hpp"

using namespace zk;

// csrc/ff.cuh as text (written by build.py): the prelude of the generated evaluate_h kernels
static const char kFfSource[] =
#include "../build/ff_embed.inc"
    ;

struct zk_pk {
  zk_ctx* ctx = nullptr;
  zk_domain* dom = nullptr;
  CompiledEval C;
  zk_evaluator_desc shape;        // scalar fields only (pointers are not retained)
  uint32_t* d_code = nullptr;
  void* d_consts = nullptr;
  void* d_resident = nullptr;     // fixed cosets | sigma cosets | l0 | l_last | l_active  (ext_n each)
  void* d_witness = nullptr;      // advice | instance | z sets | lookup (Z, A', S') cosets
  void* d_stage = nullptr;        // coefficient-form staging for the host-pointer entry point
  const void** d_cols = nullptr;  // device copy of the column pointer table
  void* d_challenges = nullptr;
  size_t n_resident = 0, n_witness = 0;
  // the circuit's own evaluate_h kernel: generated source (evalh_codegen.hpp) compiled with NVRTC at key generation
  cudaLibrary_t gen_lib = nullptr;
  cudaKernel_t gen_kernel = nullptr;
  void* d_gen_sc = nullptr;       // beta, gamma, theta, y, delta_start, delta, one
  std::string gen_log;            // why the generated kernel is not in use (empty when it is)
  double gen_compile_s = 0;
};

namespace {

#ifndef ZK_EVALH_MIN_CTAS
#define ZK_EVALH_MIN_CTAS 4
#endif
template <int NI>
__global__ void __launch_bounds__(128, ZK_EVALH_MIN_CTAS) k_evalh(EvalProgram P, EvalScalars S, const uint32_t* code,
                                               const uint4* const* cols, const uint4* consts,
                                               const uint4* challenges, const uint4* tw, uint4* out, size_t row0, size_t row1) {
  size_t idx = row0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t ext = (size_t)1 << P.ext_k;
  if (idx >= row1) return;
  EvalState<NI> st;
  Fr bt = S.one;
  if (P.ext_k > 0) {
    size_t half = ext >> 1;
    bt = Fr::load(tw + 2 * (idx & (half - 1)));     // extended_omega^idx = +-table[idx mod ext/2]
    if (idx >= half) bt = bt.neg();
  }
  Fr v = evalh_row<NI>(P, code, cols, consts, challenges, S, st, idx, bt);
  v.store(out + 2 * idx);
}

// l_active_row = 1 - (l_last + l_blind) on the extended domain (keygen.rs)
__global__ void k_l_active(const uint4* l_last, uint4* l_blind_inout, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  (Fr::one() - (Fr::load(l_last + 2 * i) + Fr::load(l_blind_inout + 2 * i))).store(l_blind_inout + 2 * i);
}

// ---- NVRTC, bound at run time (libnvrtc.so.12 ships with the CUDA toolkit of this image) --------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* lib = nullptr;
  int (*create)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*compile)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*cubin_size)(nvrtcProgram, size_t*) = nullptr;
  int (*cubin)(nvrtcProgram, char*) = nullptr;
  int (*log_size)(nvrtcProgram, size_t*) = nullptr;
  int (*log)(nvrtcProgram, char*) = nullptr;
  int (*destroy)(nvrtcProgram*) = nullptr;
};
Nvrtc* nvrtc() {
  static Nvrtc N;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      N.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (N.lib) break;
    }
    if (N.lib) {
      N.create = (decltype(N.create))dlsym(N.lib, "nvrtcCreateProgram");
      N.compile = (decltype(N.compile))dlsym(N.lib, "nvrtcCompileProgram");
      N.cubin_size = (decltype(N.cubin_size))dlsym(N.lib, "nvrtcGetCUBINSize");
      N.cubin = (decltype(N.cubin))dlsym(N.lib, "nvrtcGetCUBIN");
      N.log_size = (decltype(N.log_size))dlsym(N.lib, "nvrtcGetProgramLogSize");
      N.log = (decltype(N.log))dlsym(N.lib, "nvrtcGetProgramLog");
      N.destroy = (decltype(N.destroy))dlsym(N.lib, "nvrtcDestroyProgram");
      if (!N.create || !N.compile || !N.cubin_size || !N.cubin || !N.destroy) N.lib = nullptr;
    }
  }
  return N.lib ? &N : nullptr;
}
// source -> sm_100a cubin; empty vector + `log` on failure
std::vector<char> nvrtc_compile(const std::string& src, std::string& log) {
  std::vector<char> out;
  Nvrtc* N = nvrtc();
  if (!N) { log = "libnvrtc.so.12 not found"; return out; }
  nvrtcProgram prog = nullptr;
  if (N->create(&prog, src.c_str(), "zk_evalh_gen.cu", 0, nullptr, nullptr) != 0) { log = "nvrtcCreateProgram failed"; return out; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
  int rc = N->compile(prog, 3, opts);          // -default-device is not needed: every function is annotated
  if (rc != 0) {
    size_t ls = 0;
    if (N->log_size && N->log && N->log_size(prog, &ls) == 0 && ls > 1) { log.resize(ls); N->log(prog, &log[0]); }
    else log = "nvrtcCompileProgram failed";
    N->destroy(&prog);
    return out;
  }
  size_t cs = 0;
  if (N->cubin_size(prog, &cs) == 0 && cs) { out.resize(cs); if (N->cubin(prog, out.data()) != 0) out.clear(); }
  if (out.empty()) log = "nvrtcGetCUBIN failed";
  N->destroy(&prog);
  return out;
}

Fr fr_canon(const uint32_t c[8]) { Fr v; memcpy(v.l, c, 32); return v.to_mont(); }
Fr fr_host(const zk_fr* p) { Fr v; memcpy(v.l, p, 32); return v; }

}  // namespace

// from ntt.cu
extern "C" int zk_coeff_to_extended_dev(zk_domain* d, const zk_fr* d_a, zk_fr* d_out_ext, size_t count);
const void* zk_domain_twiddles(const zk_domain* d);
size_t zk_domain_ext_n(const zk_domain* d);

extern "C" void zk_pk_destroy(zk_pk* pk) {
  if (!pk) return;
  cudaSetDevice(pk->ctx->device);
  cudaStreamSynchronize(pk->ctx->stream);
  cudaFree(pk->d_code); cudaFree(pk->d_consts); cudaFree(pk->d_resident); cudaFree(pk->d_witness);
  cudaFree(pk->d_stage); cudaFree((void*)pk->d_cols); cudaFree(pk->d_challenges);
  if (pk->d_gen_sc) cudaFree(pk->d_gen_sc);
  if (pk->gen_lib) cudaLibraryUnload(pk->gen_lib);
  if (pk->dom) zk_domain_destroy(pk->dom);
  delete pk;
}
extern "C" zk_domain* zk_pk_domain(zk_pk* pk) { return pk ? pk->dom : nullptr; }

// The CUDA C++ text of a circuit's evaluate_h (what zk_pk_upload compiles with NVRTC); `host_flavour` != 0 gives the same
// statements as a plain C++ function over a column pointer table (CPU tests).  Host only, no device needed.
extern "C" int zk_evalh_generate_source(const zk_evaluator_desc* desc, int host_flavour, char* buf, size_t cap, size_t* len) {
  if (!desc || !len) return ZK_ERR_ARG;
  CompiledEval C;
  if (!compile_program(*desc, C, true)) return ZK_ERR_ARG;
  const std::string src = generate_evalh_source(C, *desc, host_flavour != 0, host_flavour ? std::string() : std::string(kFfSource));
  *len = src.size();
  if (buf) {
    if (cap < src.size() + 1) return ZK_ERR_ARG;
    memcpy(buf, src.c_str(), src.size() + 1);
  }
  return ZK_OK;
}
// 1 = the generated (NVRTC-compiled) kernel evaluates h(X) for this key, 0 = the interpreter
extern "C" int zk_pk_evaluator_kind(const zk_pk* pk) { return pk && pk->gen_kernel ? 1 : 0; }
// Why the generated kernel is not in use ("" when it is), and how long NVRTC took
extern "C" const char* zk_pk_evaluator_log(const zk_pk* pk, double* compile_seconds) {
  if (compile_seconds) *compile_seconds = pk ? pk->gen_compile_s : 0;
  return pk ? pk->gen_log.c_str() : "null key";
}

namespace {
// evaluate_h as generated code: emit, compile for sm_100a, load.  Any failure leaves the interpreter in charge.
void build_generated_kernel(zk_pk* pk, const zk_evaluator_desc& desc) {
  if (const char* v = getenv("ZK_EVALH")) if (!strcmp(v, "interp")) { pk->gen_log = "disabled by ZK_EVALH=interp"; return; }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  const std::string src = generate_evalh_source(pk->C, desc, false, std::string(kFfSource));
  std::string log;
  std::vector<char> cubin = nvrtc_compile(src, log);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pk->gen_compile_s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  if (cubin.empty()) { pk->gen_log = "NVRTC: " + log; return; }
  if (cudaLibraryLoadData(&pk->gen_lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
    pk->gen_log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(cudaGetLastError()); pk->gen_lib = nullptr; return;
  }
  if (cudaLibraryGetKernel(&pk->gen_kernel, pk->gen_lib, "zk_evalh_gen") != cudaSuccess) {
    pk->gen_log = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(cudaGetLastError());
    cudaLibraryUnload(pk->gen_lib); pk->gen_lib = nullptr; pk->gen_kernel = nullptr; return;
  }
  if (cudaMalloc(&pk->d_gen_sc, 7 * 32) != cudaSuccess) {
    cudaGetLastError(); cudaLibraryUnload(pk->gen_lib); pk->gen_lib = nullptr; pk->gen_kernel = nullptr; pk->gen_log = "scalar buffer"; return;
  }
}
}  // namespace
extern "C" int zk_pk_program_stats(const zk_pk* pk, uint64_t* muls_per_row, uint64_t* columns_read) {
  if (!pk) return ZK_ERR_ARG;
  if (muls_per_row) *muls_per_row = pk->C.muls_per_row;
  if (columns_read) *columns_read = pk->C.columns_read;
  return ZK_OK;
}

extern "C" int zk_pk_upload(zk_ctx* ctx, const zk_evaluator_desc* desc, const zk_fr* const* fixed,
                            const zk_fr* const* perm, const zk_fr* l0, const zk_fr* l_last,
                            const zk_fr* l_active_row, int coeff_form, zk_pk** out) {
  if (!ctx) return ZK_ERR_ARG;
  if (!desc || !out || !l0 || !l_last || !l_active_row || (desc->n_fixed && !fixed) || (desc->n_perm_columns && !perm))
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_pk_upload: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  zk_pk* pk = new zk_pk();
  pk->ctx = ctx;
  if (!compile_program(*desc, pk->C, getenv("ZK_EVALH_NOOPT") == nullptr)) {
    std::string e = pk->C.error;
    delete pk;
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_pk_upload: %s", e.c_str());
  }
  if (pk->C.max_intermediates > 512) {
    delete pk;
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_pk_upload: graph needs more than 512 intermediates");
  }
  pk->shape = *desc;
  int r = zk_domain_create(ctx, desc->degree, desc->k, &pk->dom);
  if (r != ZK_OK) { delete pk; return r; }
  const size_t n = (size_t)1 << desc->k, E = zk_domain_ext_n(pk->dom);
  const EvalProgram& P = pk->C.prog;
  pk->n_resident = (size_t)desc->n_fixed + desc->n_perm_columns + 3;
  pk->n_witness = (size_t)desc->n_advice + desc->n_instance + P.n_sets + 3 * (size_t)desc->n_lookups;
#define PK_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ctx->err = std::string("zk_pk_upload: ") + #expr + ": " + cudaGetErrorString(_e);       \
      zk_pk_destroy(pk);                                                                      \
      return _e == cudaErrorMemoryAllocation ? ZK_ERR_OOM : ZK_ERR_CUDA;                      \
    }                                                                                         \
  } while (0)
  PK_CUDA(cudaMalloc(&pk->d_code, pk->C.code.size() * 4));
  PK_CUDA(cudaMalloc(&pk->d_consts, pk->C.consts.size() * 32));
  PK_CUDA(cudaMalloc(&pk->d_resident, pk->n_resident * E * 32));
  PK_CUDA(cudaMalloc(&pk->d_witness, (pk->n_witness ? pk->n_witness : 1) * E * 32));
  PK_CUDA(cudaMalloc(&pk->d_stage, (pk->n_witness ? pk->n_witness : 1) * n * 32));
  PK_CUDA(cudaMalloc((void**)&pk->d_cols, pk->C.n_slots * sizeof(void*)));
  PK_CUDA(cudaMalloc(&pk->d_challenges, (desc->n_challenges ? desc->n_challenges : 1) * 32));
  PK_CUDA(cudaMemcpyAsync(pk->d_code, pk->C.code.data(), pk->C.code.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  PK_CUDA(cudaMemcpyAsync(pk->d_consts, pk->C.consts.data(), pk->C.consts.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
  // resident columns, in slot order: fixed, sigma, l0, l_last, l_active
  std::vector<const zk_fr*> src;
  for (uint32_t i = 0; i < desc->n_fixed; i++) src.push_back(fixed[i]);
  for (uint32_t i = 0; i < desc->n_perm_columns; i++) src.push_back(perm[i]);
  src.push_back(l0); src.push_back(l_last); src.push_back(l_active_row);
  char* res = (char*)pk->d_resident;
  if (!coeff_form) {
    for (size_t i = 0; i < src.size(); i++) {
      if (!src[i]) { zk_pk_destroy(pk); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_pk_upload: null column pointer"); }
      PK_CUDA(cudaMemcpyAsync(res + i * E * 32, src[i], E * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
  } else {
    void* stage = nullptr;
    PK_CUDA(cudaMalloc(&stage, n * 32));
    for (size_t i = 0; i < src.size(); i++) {
      if (!src[i]) { cudaFree(stage); zk_pk_destroy(pk); ZK_FAIL(ctx, ZK_ERR_ARG, "zk_pk_upload: null column pointer"); }
      PK_CUDA(cudaMemcpyAsync(stage, src[i], n * 32, cudaMemcpyDefault, ctx->stream));   // host or device source
      int rr = zk_coeff_to_extended_dev(pk->dom, (const zk_fr*)stage, (zk_fr*)(res + i * E * 32), 1);
      if (rr != ZK_OK) { cudaFree(stage); zk_pk_destroy(pk); return rr; }
      PK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(stage);
    // l_active_row = 1 - (l_last + l_blind)
    size_t i_last = src.size() - 2, i_act = src.size() - 1;
    zk_timing_begin(ctx, "l_active");
    k_l_active<<<zk_div_up(E, 256), 256, 0, ctx->stream>>>((const uint4*)(res + i_last * E * 32), (uint4*)(res + i_act * E * 32), E);
    zk_timing_end(ctx);
  }
  PK_CUDA(cudaStreamSynchronize(ctx->stream));
#undef PK_CUDA
  build_generated_kernel(pk, *desc);
  *out = pk;
  return ZK_OK;
}

namespace {

// Fills the column pointer table and launches the fused kernel.  `wit` holds the witness
// cosets in slot order (advice, instance, z sets, lookup triples), E elements each.
int launch_evalh(zk_pk* pk, const zk_fr* challenges, const zk_fr* y, const zk_fr* beta, const zk_fr* gamma,
                 const zk_fr* theta, zk_fr* d_out_ext, size_t row0 = 0, size_t row1 = ~(size_t)0) {
  zk_ctx* ctx = pk->ctx;
  const EvalProgram& P = pk->C.prog;
  const zk_evaluator_desc& D = pk->shape;
  const size_t E = zk_domain_ext_n(pk->dom);
  std::vector<const void*> cols(pk->C.n_slots);
  const char* res = (const char*)pk->d_resident;
  const char* wit = (const char*)pk->d_witness;
  size_t ri = 0, wi = 0;
  for (uint32_t i = 0; i < D.n_fixed; i++) cols[P.off_fixed + i] = res + (ri++) * E * 32;
  for (uint32_t i = 0; i < D.n_perm_columns; i++) cols[P.off_sigma + i] = res + (ri++) * E * 32;
  cols[P.off_l0] = res + (ri++) * E * 32;
  cols[P.off_l_last] = res + (ri++) * E * 32;
  cols[P.off_l_active] = res + (ri++) * E * 32;
  for (uint32_t i = 0; i < D.n_advice; i++) cols[P.off_advice + i] = wit + (wi++) * E * 32;
  for (uint32_t i = 0; i < D.n_instance; i++) cols[P.off_instance + i] = wit + (wi++) * E * 32;
  for (uint32_t i = 0; i < P.n_sets; i++) cols[P.off_z + i] = wit + (wi++) * E * 32;
  for (uint32_t i = 0; i < 3 * P.n_lookups; i++) cols[P.off_lookup + i] = wit + (wi++) * E * 32;
  ZK_CUDA(ctx, cudaMemcpyAsync((void*)pk->d_cols, cols.data(), cols.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  if (D.n_challenges)
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->d_challenges, challenges, D.n_challenges * 32, cudaMemcpyHostToDevice, ctx->stream));
  // the pageable-memory copies above are staged by the runtime before returning, so `cols` may go out of scope
  EvalScalars S;
  S.beta = fr_host(beta); S.gamma = fr_host(gamma); S.theta = fr_host(theta); S.y = fr_host(y);
  const uint32_t zeta_c[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
  const uint32_t delta_c[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
  S.delta_start = S.beta * fr_canon(zeta_c);
  S.delta = fr_canon(delta_c);
  S.one = Fr::one();
  const uint4* tw = (const uint4*)zk_domain_twiddles(pk->dom);
  if (row1 > E) row1 = E;
  if (row0 >= row1) return ZK_OK;
  unsigned grid = zk_div_up(row1 - row0, 128);
  if (pk->gen_kernel) {
    // the circuit's own compiled kernel: the seven scalars travel in a small buffer (staged copy of a stack array)
    Fr scv[7] = {S.beta, S.gamma, S.theta, S.y, S.delta_start, S.delta, S.one};
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->d_gen_sc, scv, sizeof scv, cudaMemcpyHostToDevice, ctx->stream));
    const uint4* a_res = (const uint4*)pk->d_resident; const uint4* a_wit = (const uint4*)pk->d_witness;
    const uint4* a_ch = (const uint4*)pk->d_challenges; const uint4* a_sc = (const uint4*)pk->d_gen_sc;
    uint4* a_out = (uint4*)d_out_ext;
    unsigned long long a_r0 = row0, a_r1 = row1;
    void* args[] = {&a_res, &a_wit, &a_ch, &a_sc, &tw, &a_out, &a_r0, &a_r1};
    zk_timing_begin(ctx, "evaluate_h");
    cudaError_t le = cudaLaunchKernel((const void*)pk->gen_kernel, dim3(grid), dim3(128), args, 0, ctx->stream);
    zk_timing_end(ctx);
    if (le != cudaSuccess) ZK_FAIL(ctx, ZK_ERR_CUDA, "launch generated evaluate_h: %s", cudaGetErrorString(le));
    return ZK_OK;
  }
  uint32_t ni = pk->C.max_intermediates;
  // measured on B200: the select-based register file (NI = 4) is slower than 16 L1-resident slots (21.5 vs 19.2 ms)
  uint32_t ni_min = 16;
  if (const char* v = getenv("ZK_EVALH_NI_MIN")) ni_min = (uint32_t)atoi(v);
  if (ni < ni_min) ni = ni_min;
#define EVALH_ARGS P, S, (const uint32_t*)pk->d_code, (const uint4* const*)pk->d_cols, (const uint4*)pk->d_consts, \
                   (const uint4*)pk->d_challenges, tw, (uint4*)d_out_ext, row0, row1
  if (ni <= 4) { ZK_LAUNCH(ctx, "evaluate_h", k_evalh<4>, grid, 128, 0, EVALH_ARGS); }
  else if (ni <= 16) { ZK_LAUNCH(ctx, "evaluate_h", k_evalh<16>, grid, 128, 0, EVALH_ARGS); }
  else if (ni <= 32) { ZK_LAUNCH(ctx, "evaluate_h", k_evalh<32>, grid, 128, 0, EVALH_ARGS); }
  else if (ni <= 128) { ZK_LAUNCH(ctx, "evaluate_h", k_evalh<128>, grid, 128, 0, EVALH_ARGS); }
  else { ZK_LAUNCH(ctx, "evaluate_h", k_evalh<512>, grid, 128, 0, EVALH_ARGS); }
#undef EVALH_ARGS
  return ZK_OK;
}

}  // namespace

// ---- internal entry points for prover.cu: the witness cosets are produced by the caller (on another stream) -------
// slot order: advice | instance | permutation z sets | lookup (Z, A', S') triples; 2^extended_k elements each
size_t zk_pk_witness_bytes(zk_pk* pk) { return (pk->n_witness ? pk->n_witness : 1) * zk_domain_ext_n(pk->dom) * 32; }
char* zk_pk_witness_slot(zk_pk* pk, size_t slot) { return (char*)pk->d_witness + slot * zk_domain_ext_n(pk->dom) * 32; }
// rows [row0, row1) of the extended domain only (a group of GPUs splits the rows)
int zk_evaluate_h_launch(zk_pk* pk, const zk_fr* challenges, const zk_fr* y, const zk_fr* beta, const zk_fr* gamma,
                         const zk_fr* theta, zk_fr* d_out_ext, size_t row0, size_t row1) {
  return launch_evalh(pk, challenges, y, beta, gamma, theta, d_out_ext, row0, row1);
}

extern "C" int zk_evaluate_h_dev(zk_pk* pk, const zk_fr* const* d_advice, const zk_fr* const* d_instance,
                                 const zk_fr* challenges, const zk_fr* y, const zk_fr* beta, const zk_fr* gamma,
                                 const zk_fr* theta, const zk_fr* const* d_lookup, const zk_fr* const* d_perm_product,
                                 zk_fr* d_out_ext) {
  if (!pk) return ZK_ERR_ARG;
  zk_ctx* ctx = pk->ctx;
  const zk_evaluator_desc& D = pk->shape;
  const EvalProgram& P = pk->C.prog;
  if (!y || !beta || !gamma || !theta || !d_out_ext || (D.n_advice && !d_advice) || (D.n_instance && !d_instance) ||
      (D.n_challenges && !challenges) || (D.n_lookups && !d_lookup) || (P.n_sets && !d_perm_product))
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_evaluate_h: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t E = zk_domain_ext_n(pk->dom);
  // coeff_to_extended of every witness polynomial, in slot order
  std::vector<const zk_fr*> src;
  for (uint32_t i = 0; i < D.n_advice; i++) src.push_back(d_advice[i]);
  for (uint32_t i = 0; i < D.n_instance; i++) src.push_back(d_instance[i]);
  for (uint32_t i = 0; i < P.n_sets; i++) src.push_back(d_perm_product[i]);
  for (uint32_t i = 0; i < 3 * P.n_lookups; i++) src.push_back(d_lookup[i]);
  char* wit = (char*)pk->d_witness;
  for (size_t i = 0; i < src.size(); i++) {
    if (!src[i]) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_evaluate_h: null polynomial pointer");
    ZK_TRY(zk_coeff_to_extended_dev(pk->dom, src[i], (zk_fr*)(wit + i * E * 32), 1));
  }
  return launch_evalh(pk, challenges, y, beta, gamma, theta, d_out_ext);
}

extern "C" int zk_evaluate_h(zk_pk* pk, const zk_fr* const* advice, const zk_fr* const* instance,
                             const zk_fr* challenges, const zk_fr* y, const zk_fr* beta, const zk_fr* gamma,
                             const zk_fr* theta, const zk_fr* const* lookup, const zk_fr* const* perm_product,
                             zk_fr* out_ext) {
  if (!pk) return ZK_ERR_ARG;
  zk_ctx* ctx = pk->ctx;
  const zk_evaluator_desc& D = pk->shape;
  const EvalProgram& P = pk->C.prog;
  if (!y || !beta || !gamma || !theta || !out_ext || (D.n_advice && !advice) || (D.n_instance && !instance) ||
      (D.n_challenges && !challenges) || (D.n_lookups && !lookup) || (P.n_sets && !perm_product))
    ZK_FAIL(ctx, ZK_ERR_ARG, "zk_evaluate_h: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << D.k, E = zk_domain_ext_n(pk->dom);
  std::vector<const zk_fr*> src;
  for (uint32_t i = 0; i < D.n_advice; i++) src.push_back(advice[i]);
  for (uint32_t i = 0; i < D.n_instance; i++) src.push_back(instance[i]);
  for (uint32_t i = 0; i < P.n_sets; i++) src.push_back(perm_product[i]);
  for (uint32_t i = 0; i < 3 * P.n_lookups; i++) src.push_back(lookup[i]);
  // upload polynomial i+1 on the copy stream while polynomial i is being extended on the compute stream
  char* stage = (char*)pk->d_stage;
  char* wit = (char*)pk->d_witness;
  ZK_TRY(zk_copy_stream(ctx));
  for (size_t i = 0; i < src.size(); i++) {
    if (!src[i]) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_evaluate_h: null polynomial pointer");
    ZK_CUDA(ctx, cudaMemcpyAsync(stage + i * n * 32, src[i], n * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
    ZK_CUDA(ctx, cudaEventRecord(ctx->copy_event, ctx->copy_stream));
    ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_event, 0));
    ZK_TRY(zk_coeff_to_extended_dev(pk->dom, (const zk_fr*)(stage + i * n * 32), (zk_fr*)(wit + i * E * 32), 1));
  }
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, E * 32));
  ZK_TRY(launch_evalh(pk, challenges, y, beta, gamma, theta, (zk_fr*)ctx->misc.p));
  ZK_CUDA(ctx, cudaMemcpyAsync(out_ext, ctx->misc.p, E * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}
