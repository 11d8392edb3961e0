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
//
// Formats (zk_serde_format): the halo2-axiom stack the reference's CLI runs on (halo2curves-axiom / halo2curves >= 0.4,
// Cargo.lock:1181-1199) moved the parity of y to bit 6 and flags the identity in bit 7 [UPSTREAM?]; the six flagged
// commitments of the reference's own golden proof (bin/assets/proof.bin, bin/src/main.rs:270-279) pin the sign position
// (tests/golden/ref_proof_points.json).  SerdeFormat::RawBytes is x | y as Montgomery limbs = the in-memory layout.
#include "common.cuh"
#include "ec.cuh"

using namespace zk;

namespace {

// (p + 1) / 4, little-endian u32 limbs
__device__ __constant__ uint32_t kSqrtExp[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};

__global__ void __launch_bounds__(128) k_g1_compress(const uint4* pts, size_t n, uint4* out, int v04) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine a = Affine::load(pts + 4 * i);
  Fq x = a.x.from_mont(), y = a.y.from_mont();
  if (a.is_identity()) {
    x = Fq::zero();
    if (v04) x.l[7] = 0x80000000u;
  } else {
    x.l[7] |= (y.l[0] & 1u) << (v04 ? 30 : 31);
  }
  x.store_raw(out + 2 * i);
}
// ok[0] |= 1 when some encoding is invalid (x >= p, or x^3 + 3 is not a square)
__global__ void __launch_bounds__(128) k_g1_decompress(const uint4* in, size_t n, uint4* pts, uint32_t* bad, int v04) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = Fq::load(in + 2 * i);
  uint32_t sign;
  if (v04) {
    uint32_t inf = x.l[7] >> 31;
    sign = (x.l[7] >> 30) & 1u;
    x.l[7] &= 0x3fffffffu;
    if (inf) {
      if (!x.is_zero_raw() || sign) atomicOr(bad, 1u);
      Affine::identity().store(pts + 4 * i);
      return;
    }
  } else {
    sign = x.l[7] >> 31;
    x.l[7] &= 0x7fffffffu;
    if (x.is_zero_raw() && !sign) { Affine::identity().store(pts + 4 * i); return; }
  }
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

// SerdeFormat::RawBytes reads check that the point is on the curve (RawBytesUnchecked does not; both accepted here, checked)
__global__ void __launch_bounds__(128) k_g1_check(const uint4* pts, size_t n, uint32_t* bad) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine a = Affine::load(pts + 4 * i);
  if (a.is_identity()) return;
  Fq three = Fq::one() + Fq::one() + Fq::one();
  if (a.y.sqr() != a.x.sqr() * a.x + three) atomicOr(bad, 1u);
}

}  // namespace

extern "C" int zk_srs_upload_dev(zk_ctx* ctx, uint32_t k, const zk_g1_affine* d_g, const zk_g1_affine* d_g_lagrange, int precompute_windows);

extern "C" int zk_g1_encode(zk_ctx* ctx, int format, const zk_g1_affine* points, size_t n, uint8_t* out) {
  if (!ctx) return ZK_ERR_ARG;
  if ((n && (!points || !out)) || format < 0 || format > ZK_SERDE_RAW_BYTES) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_encode: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!n) return ZK_OK;
  if (format == ZK_SERDE_RAW_BYTES) {
    ZK_CUDA(ctx, cudaMemcpyAsync(out, points, n * 64, cudaMemcpyDefault, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
  }
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, n * 96));
  char* d = (char*)ctx->misc.p;
  ZK_CUDA(ctx, cudaMemcpyAsync(d, points, n * 64, cudaMemcpyDefault, ctx->stream));
  ZK_LAUNCH(ctx, "g1_compress", k_g1_compress, zk_div_up(n, 128), 128, 0, (const uint4*)d, n, (uint4*)(d + n * 64), format == ZK_SERDE_PROCESSED_V04);
  ZK_CUDA(ctx, cudaMemcpyAsync(out, d + n * 64, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZK_OK;
}
extern "C" int zk_g1_compress(zk_ctx* ctx, const zk_g1_affine* points, size_t n, uint8_t* out) {
  return zk_g1_encode(ctx, ZK_SERDE_PROCESSED, points, n, out);
}

namespace {
// bytes (host) -> affine points in device memory at d_out
int decompress_to(zk_ctx* ctx, int format, const uint8_t* in, size_t n, char* d_in, char* d_out, uint32_t* d_flag) {
  if (format == ZK_SERDE_RAW_BYTES) {
    ZK_CUDA(ctx, cudaMemcpyAsync(d_out, in, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    ZK_LAUNCH(ctx, "g1_check", k_g1_check, zk_div_up(n, 128), 128, 0, (const uint4*)d_out, n, d_flag);
    return ZK_OK;
  }
  ZK_CUDA(ctx, cudaMemcpyAsync(d_in, in, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  ZK_LAUNCH(ctx, "g1_decompress", k_g1_decompress, zk_div_up(n, 128), 128, 0, (const uint4*)d_in, n, (uint4*)d_out, d_flag, format == ZK_SERDE_PROCESSED_V04);
  return ZK_OK;
}
inline size_t g1_bytes(int format) { return format == ZK_SERDE_RAW_BYTES ? 64 : 32; }
inline size_t g2_bytes(int format) { return format == ZK_SERDE_RAW_BYTES ? 128 : 64; }
}  // namespace

extern "C" int zk_g1_decompress(zk_ctx* ctx, const uint8_t* in, size_t n, zk_g1_affine* out) {
  return zk_g1_decode(ctx, ZK_SERDE_PROCESSED, in, n, out);
}
extern "C" int zk_g1_decode(zk_ctx* ctx, int format, const uint8_t* in, size_t n, zk_g1_affine* out) {
  if (!ctx) return ZK_ERR_ARG;
  if ((n && (!in || !out)) || format < 0 || format > ZK_SERDE_RAW_BYTES) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_decode: bad argument");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!n) return ZK_OK;
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, n * 96 + 256));
  char* d = (char*)ctx->misc.p;
  uint32_t* flag = (uint32_t*)(d + n * 96);
  ZK_CUDA(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
  ZK_TRY(decompress_to(ctx, format, in, n, d, d + n * 32, flag));
  uint32_t bad = 0;
  ZK_CUDA(ctx, cudaMemcpyAsync(out, d + n * 32, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_g1_decode: invalid point encoding");
  return ZK_OK;
}

extern "C" int zk_params_read(zk_ctx* ctx, const uint8_t* file, size_t len, int precompute_windows, uint32_t* k_out,
                              uint8_t g2[64], uint8_t s_g2[64]) {
  return zk_params_read_format(ctx, ZK_SERDE_PROCESSED, file, len, precompute_windows, k_out, g2, s_g2);
}
extern "C" int zk_params_read_format(zk_ctx* ctx, int format, const uint8_t* file, size_t len, int precompute_windows,
                                     uint32_t* k_out, uint8_t* g2, uint8_t* s_g2) {
  if (!ctx) return ZK_ERR_ARG;
  if (!file || len < 4 || format < 0 || format > ZK_SERDE_RAW_BYTES) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: bad argument");
  const size_t pb = g1_bytes(format), qb = g2_bytes(format);
  uint32_t k;
  memcpy(&k, file, 4);
  if (k > 27) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: k = %u out of range", k);
  const size_t n = (size_t)1 << k;
  if (len != 4 + 2 * n * pb + 2 * qb) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: %zu bytes, expected %zu for k = %u", len, 4 + 2 * n * pb + 2 * qb, k);
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, 2 * n * 96 + 256));
  char* d = (char*)ctx->misc.p;
  uint32_t* flag = (uint32_t*)(d + 2 * n * 96);
  ZK_CUDA(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
  ZK_TRY(decompress_to(ctx, format, file + 4, 2 * n, d, d + 2 * n * 32, flag));
  uint32_t bad = 0;
  ZK_CUDA(ctx, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_read: invalid point encoding in the SRS");
  const zk_g1_affine* pts = (const zk_g1_affine*)(d + 2 * n * 32);
  ZK_TRY(zk_srs_upload_dev(ctx, k, pts, pts + n, precompute_windows));
  if (g2) memcpy(g2, file + 4 + 2 * n * pb, qb);
  if (s_g2) memcpy(s_g2, file + 4 + 2 * n * pb + qb, qb);
  if (k_out) *k_out = k;
  return ZK_OK;
}

extern "C" int zk_params_write(zk_ctx* ctx, const uint8_t g2[64], const uint8_t s_g2[64], uint8_t* file, size_t cap, size_t* len) {
  return zk_params_write_format(ctx, ZK_SERDE_PROCESSED, g2, s_g2, file, cap, len);
}
extern "C" int zk_params_write_format(zk_ctx* ctx, int format, const uint8_t* g2, const uint8_t* s_g2, uint8_t* file, size_t cap,
                                      size_t* len) {
  if (!ctx) return ZK_ERR_ARG;
  if (!file || !len || format < 0 || format > ZK_SERDE_RAW_BYTES) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_write: bad argument");
  if (!ctx->srs_g.bases || !ctx->srs_lagrange.bases) ZK_FAIL(ctx, ZK_ERR_STATE, "zk_params_write: no SRS resident");
  const size_t pb = g1_bytes(format), qb = g2_bytes(format);
  const size_t n = ctx->srs_n, need = 4 + 2 * n * pb + 2 * qb;
  if (cap < need) ZK_FAIL(ctx, ZK_ERR_ARG, "zk_params_write: needs %zu bytes", need);
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t k = ctx->srs_k;
  memcpy(file, &k, 4);
  if (format == ZK_SERDE_RAW_BYTES) {
    ZK_CUDA(ctx, cudaMemcpyAsync(file + 4, ctx->srs_g.bases, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(file + 4 + n * 64, ctx->srs_lagrange.bases, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    const int v04 = format == ZK_SERDE_PROCESSED_V04;
    ZK_TRY(zk_scratch_reserve(ctx, ctx->misc, 2 * n * 32));
    char* d = (char*)ctx->misc.p;
    ZK_LAUNCH(ctx, "g1_compress", k_g1_compress, zk_div_up(n, 128), 128, 0, (const uint4*)ctx->srs_g.bases, n, (uint4*)d, v04);
    ZK_LAUNCH(ctx, "g1_compress", k_g1_compress, zk_div_up(n, 128), 128, 0, (const uint4*)ctx->srs_lagrange.bases, n, (uint4*)(d + n * 32), v04);
    ZK_CUDA(ctx, cudaMemcpyAsync(file + 4, d, 2 * n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  }
  ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint8_t* tail = file + 4 + 2 * n * pb;
  if (g2) memcpy(tail, g2, qb); else memset(tail, 0, qb);
  if (s_g2) memcpy(tail + qb, s_g2, qb); else memset(tail + qb, 0, qb);
  *len = need;
  return ZK_OK;
}
