// params.cu -- ParamsKZG on-disk format and G1 point (de)compression on the device.
//
// Restates halo2_proofs poly/kzg/commitment.rs `ParamsKZG::{write, read}` ([UPSTREAM] @ c254c75f, Cargo.lock:1314-1316)
// as used by halo2-base `gen_srs` (reference call sites circuits/src/sgx_dcap_verifier.rs:799,
// crates/p256-ecdsa/src/base.rs:118,134, bin/src/main.rs:227-231):
//     k (u32 LE) | n compressed g | n compressed g_lagrange | g2 (64 B) | s_g2 (64 B)
// and halo2curves 0.3.1 G1Affine::{to_bytes, from_bytes}: x little-endian, parity of y in bit 7 of byte 31, identity =
// all zeros [UPSTREAM?: SURVEY A.10 (1)].  Reading a 2^19 SRS decompresses 2^20 points, i.e. 2^20 square roots in Fq
// (y = (x^3 + 3)^((p+1)/4)): one thread per point here, seconds of multi-threaded CPU work upstream.
// The G2 elements are not used by the prover and pass through as opaque bytes.
#include "common.cuh"
#include "ec.cuh"

using namespace zk;

namespace {

// (p + 1) / 4, little-endian u32 limbs
__device__ __constant__ uint32_t kSqrtExp[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};

__global__ void __launch_bounds__(128) k_g1_compress(const uint4* pts, size_t n, uint4* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine a = Affine::load(pts + 4 * i);
  Fq x = a.x.from_mont(), y = a.y.from_mont();
  if (a.is_identity()) { x = Fq::zero(); y = Fq::zero(); }
  x.l[7] |= (y.l[0] & 1u) << 31;
  x.store_raw(out + 2 * i);
}
// ok[0] |= 1 when some encoding is invalid (x >= p, or x^3 + 3 is not a square)
__global__ void __launch_bounds__(128) k_g1_decompress(const uint4* in, size_t n, uint4* pts, uint32_t* bad) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = Fq::load(in + 2 * i);
  uint32_t sign = x.l[7] >> 31;
  x.l[7] &= 0x7fffffffu;
  if (x.is_zero_raw() && !sign) { Affine::identity().store(pts + 4 * i); return; }
  // canonical check: x < p
  Fq t, m = Fq::kp<1>();
  if (!sub8(t.l, x.l, m.l)) { atomicOr(bad, 1u); Affine::identity().store(pts + 4 * i); return; }
  Fq xm = x.to_mont();
  Fq three = Fq::one() + Fq::one() + Fq::one();
  Fq rhs = xm.sqr() * xm + three;
  uint32_t e[8];
#pragma unroll
  for (int k = 0; k < 8; k++) e[k] = kSqrtExp[k];
  Fq y = Fq::pow(rhs, e);
  if (y.sqr() != rhs) { atomicOr(bad, 1u); Affine::identity().store(pts + 4 * i); return; }
  if ((y.from_mont().l[0] & 1u) != sign) y = y.neg();
  Affine a; a.x = xm; a.y = y;
  a.store(pts + 4 * i);
}

}  // namespace

extern "C" int zk_srs_upload_dev(zk_ctx* ctx, uint32_t k, const zk_g1_affine* d_g, const zk_g1_affine* d_g_lagrange, int precompute_windows);

extern "C" int zk_g1_compress(zk_ctx* ctx, const zk_g1_affine* points, size_t n, uint8_t* out) {
  if (!ctx) return ZK_ERR_ARG;
  if (n && (!points || !out)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_compress: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!n) return ZK_OK;
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, n * 96));
  char* d = (char*)ctx->misc.p;
  ZK_CUDA(ctx, cudaMemcpyAsync(d, points, n * 64, cudaMemcpyDefault, ctx->stream));
  ZK_LAUNCH(ctx, "g1_compress", k_g1_compress, zk_div_up(n, 128), 128, 0, (const uint4*)d, n, (uint4*)(d + n * 64));
  ZK_CUDA(ctx, cudaMemcpyAsync(out, d + n * 64, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}

namespace {
// bytes (host) -> affine points in device memory at d_out
int decompress_to(zk_ctx* ctx, const uint8_t* in, size_t n, char* d_in, char* d_out, uint32_t* d_flag) {
  ZK_CUDA(ctx, cudaMemcpyAsync(d_in, in, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  ZK_LAUNCH(ctx, "g1_decompress", k_g1_decompress, zk_div_up(n, 128), 128, 0, (const uint4*)d_in, n, (uint4*)d_out, d_flag);
  return ZK_OK;
}
}  // namespace

extern "C" int zk_g1_decompress(zk_ctx* ctx, const uint8_t* in, size_t n, zk_g1_affine* out) {
  if (!ctx) return ZK_ERR_ARG;
  if (n && (!in || !out)) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_decompress: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!n) return ZK_OK;
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, n * 96 + 256));
  char* d = (char*)ctx->misc.p;
  uint32_t* flag = (uint32_t*)(d + n * 96);
  ZK_CUDA(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
  ZK_TRY(decompress_to(ctx, in, n, d, d + n * 32, flag));
  uint32_t bad = 0;
  ZK_CUDA(ctx, cudaMemcpyAsync(out, d + n * 32, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_decompress: invalid point encoding");
  return ZK_OK;
}

extern "C" int zk_params_read(zk_ctx* ctx, const uint8_t* file, size_t len, int precompute_windows, uint32_t* k_out,
                              uint8_t g2[64], uint8_t s_g2[64]) {
  if (!ctx) return ZK_ERR_ARG;
  if (!file || len < 4) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: bad argument");
  uint32_t k;
  memcpy(&k, file, 4);
  if (k > 27) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: k = %u out of range", k);
  const size_t n = (size_t)1 << k;
  if (len != 4 + 2 * n * 32 + 128) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: %zu bytes, expected %zu for k = %u", len, 4 + 2 * n * 32 + 128, k);
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, 2 * n * 96 + 256));
  char* d = (char*)ctx->misc.p;
  uint32_t* flag = (uint32_t*)(d + 2 * n * 96);
  ZK_CUDA(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
  ZK_TRY(decompress_to(ctx, file + 4, 2 * n, d, d + 2 * n * 32, flag));
  uint32_t bad = 0;
  ZK_CUDA(ctx, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: invalid point encoding in the SRS");
  const zk_g1_affine* pts = (const zk_g1_affine*)(d + 2 * n * 32);
  ZK_TRY(zk_srs_upload_dev(ctx, k, pts, pts + n, precompute_windows));
  if (g2) memcpy(g2, file + 4 + 2 * n * 32, 64);
  if (s_g2) memcpy(s_g2, file + 4 + 2 * n * 32 + 64, 64);
  if (k_out) *k_out = k;
  return ZK_OK;
}

extern "C" int zk_params_write(zk_ctx* ctx, const uint8_t g2[64], const uint8_t s_g2[64], uint8_t* file, size_t cap, size_t* len) {
  if (!ctx) return ZK_ERR_ARG;
  if (!file || !len) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_write: bad argument");
  if (!ctx->srs_g.bases || !ctx->srs_lagrange.bases) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_params_write: no SRS resident");
  const size_t n = ctx->srs_n, need = 4 + 2 * n * 32 + 128;
  if (cap < need) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_write: needs %zu bytes", need);
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, 2 * n * 32));
  char* d = (char*)ctx->misc.p;
  ZK_LAUNCH(ctx, "g1_compress", k_g1_compress, zk_div_up(n, 128), 128, 0, (const uint4*)ctx->srs_g.bases, n, (uint4*)d);
  ZK_LAUNCH(ctx, "g1_compress", k_g1_compress, zk_div_up(n, 128), 128, 0, (const uint4*)ctx->srs_lagrange.bases, n, (uint4*)(d + n * 32));
  uint32_t k = ctx->srs_k;
  memcpy(file, &k, 4);
  ZK_CUDA(ctx, cudaMemcpyAsync(file + 4, d, 2 * n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (g2) memcpy(file + 4 + 2 * n * 32, g2, 64); else memset(file + 4 + 2 * n * 32, 0, 64);
  if (s_g2) memcpy(file + 4 + 2 * n * 32 + 64, s_g2, 64); else memset(file + 4 + 2 * n * 32 + 64, 0, 64);
  *len = need;
  return ZK_OK;
}
