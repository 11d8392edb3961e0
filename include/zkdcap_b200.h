/*
 * zkdcap_b200.h -- C ABI of the B200-native halo2 prover backend.
 *
 * This is the drop-in boundary for the hot path of zk-dcap-verifier's proof
 * generation.  The reference has no FFI of its own: the
 * boundary is the three halo2_proofs functions that `create_proof`
 * (reference call site circuits/src/sgx_dcap_verifier.rs:814-822), `keygen_vk`
 * (:803), `keygen_pk` (:807) and snark-verifier-sdk `gen_proof`
 * (crates/p256-ecdsa/src/base.rs:193-212) reach:
 *
 *   halo2_proofs::arithmetic::best_multiexp          -> zk_msm / zk_commit*
 *   halo2_proofs::arithmetic::best_fft               -> zk_ntt
 *   halo2_proofs::poly::EvaluationDomain::*          -> zk_domain_* / zk_lagrange_to_coeff ...
 *   halo2_proofs::plonk::evaluation::Evaluator::evaluate_h -> zk_evaluate_h
 *
 * (halo2_proofs 0.2.0 @ zkwebauthn/halo2 c254c75f, Cargo.lock:1314-1316;
 *  halo2curves 0.3.1 @ bdb2e661, Cargo.lock:1329-1331 -- sources not vendored.)
 *
 * Data layout is exactly what halo2curves keeps in memory:
 *   Fr / Fq   : 4 x u64 little-endian limbs, MONTGOMERY form (R = 2^256)
 *   G1Affine  : {x, y}, 64 bytes, identity encoded as (0, 0)
 *   G1        : Jacobian {x, y, z}, 96 bytes, identity has z = 0
 * No conversion happens at the boundary.
 *
 * Every entry point returns 0 on success or a negative zk_status; the text of
 * the last failure of a context is available from zk_last_error().  The
 * library has NO CPU fallback: without a CUDA device every compute entry point
 * fails with ZK_ERR_CUDA.
 *
 * Threading: one zk_ctx = one device + one stream; calls on a context are
 * synchronous at the boundary unless the name ends in _async/_dev.  Different
 * contexts may be driven from different host threads concurrently.
 */
#ifndef ZKDCAP_B200_H
#define ZKDCAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } zk_fr;        /* bn256::Fr, Montgomery */
typedef struct { uint64_t l[4]; } zk_fq;        /* bn256::Fq, Montgomery */
typedef struct { zk_fq x, y; } zk_g1_affine;    /* bn256::G1Affine */
typedef struct { zk_fq x, y, z; } zk_g1;        /* bn256::G1 (Jacobian) */

typedef enum {
  ZK_OK = 0,
  ZK_ERR_ARG = -1,     /* bad argument (null pointer, size not a power of two ...) */
  ZK_ERR_CUDA = -2,    /* CUDA runtime failure, or no device */
  ZK_ERR_OOM = -3,     /* device allocation failed */
  ZK_ERR_STATE = -4    /* object used before the data it needs was uploaded */
} zk_status;

typedef struct zk_ctx zk_ctx;
typedef struct zk_domain zk_domain;
typedef struct zk_pk zk_pk;

/* ---- context ----------------------------------------------------------- */

/* Create a context on `device`.  `cuda_stream` may be NULL (the context then
 * owns a private non-blocking stream) or an existing cudaStream_t, e.g.
 * torch.cuda.current_stream().cuda_stream, on which all work is enqueued. */
int zk_ctx_create(int device, void* cuda_stream, zk_ctx** out);
void zk_ctx_destroy(zk_ctx* ctx);
const char* zk_last_error(const zk_ctx* ctx);
int zk_ctx_sync(zk_ctx* ctx);
/* Library/ABI version and the sm architecture the kernels were built for. */
const char* zk_version(void);

/* Per-kernel timing (CUDA events on the context's stream).  When enabled every
 * kernel family records its device time and launch count. */
typedef struct {
  char name[40];
  uint64_t launches;
  double total_ms;
} zk_kernel_stat;
int zk_ctx_timing(zk_ctx* ctx, int enable);
int zk_ctx_timing_reset(zk_ctx* ctx);
/* Returns the number of rows written (<= cap). Resolves pending events. */
int zk_ctx_timing_read(zk_ctx* ctx, zk_kernel_stat* rows, int cap);
/* Number of kernels this context has launched since creation. */
uint64_t zk_ctx_launch_count(const zk_ctx* ctx);

/* Device memory helpers (so a host language without a CUDA binding can keep
 * operands resident).  Pointers are plain device addresses. */
int zk_dev_alloc(zk_ctx* ctx, size_t bytes, void** dptr);
int zk_dev_free(zk_ctx* ctx, void* dptr);
int zk_h2d(zk_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int zk_d2h(zk_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- ParamsKZG: SRS resident in HBM ------------------------------------- */

/* Mirrors ParamsKZG { k, n, g, g_lagrange } (halo2_proofs poly/kzg/commitment.rs).
 * Uploads both base sets once; with `precompute_windows` != 0 the library also
 * builds, on the device, the table 2^(c*w) * base[i] for every Pippenger window
 * so that all windows share one bucket set (costs W x the base memory). */
int zk_srs_upload(zk_ctx* ctx, uint32_t k, const zk_g1_affine* g,
                  const zk_g1_affine* g_lagrange, int precompute_windows);

/* Same with both base sets already in device memory (they are copied). */
int zk_srs_upload_dev(zk_ctx* ctx, uint32_t k, const zk_g1_affine* d_g,
                      const zk_g1_affine* d_g_lagrange, int precompute_windows);
/* d_out[i] = d_scalars[i] * base, affine (the fixed-base products of
 * ParamsKZG::setup: g[i] = s^i * G).  Device pointers; asynchronous. */
int zk_g1_fixed_base_mul_dev(zk_ctx* ctx, const zk_g1_affine* base, const zk_fr* d_scalars,
                             size_t n, zk_g1_affine* d_out);

/* Lets `dst` (a second context = a second stream on the same device) commit
 * against the SRS that `src` keeps resident; `src` must outlive `dst`'s use. */
int zk_ctx_share_srs(zk_ctx* dst, const zk_ctx* src);

/* Window width c and window count W of the resident table of base set `which` (ZK_BASES_G / ZK_BASES_G_LAGRANGE);
 * both 0 when the SRS was uploaded without precompute_windows.  Lets callers and tests see which Pippenger
 * configuration a commitment of this SRS runs. */
int zk_srs_window(zk_ctx* ctx, int which, uint32_t* c_out, uint32_t* w_out);

/* Work counters kept on the device: out[0] = bucket entries (non-zero signed digits, i.e. the group additions the
 * bucket accumulation performs), out[1] = MSMs, out[2..3] reserved; since context creation or the last reset.
 * Synchronises the context stream.  Measurement only (bench.py's integer-pipe roofline of msm_accumulate). */
int zk_ctx_counters(zk_ctx* ctx, uint64_t out[4], int reset);

/* ParamsKZG::read / write (poly/kzg/commitment.rs; halo2-base gen_srs keeps params/kzg_bn254_{k}.srs):
 *   k (u32 LE) | n compressed g | n compressed g_lagrange | g2 (64 B) | s_g2 (64 B)
 * zk_params_read decompresses the 2n points on the device (one square root in Fq each) and keeps them resident
 * exactly like zk_srs_upload; g2 / s_g2 are returned as the opaque 64-byte encodings (the prover never uses them).
 * zk_params_write serialises the resident SRS.  Point encoding = halo2curves 0.3.1 G1Affine::to_bytes. */
int zk_params_read(zk_ctx* ctx, const uint8_t* file, size_t len, int precompute_windows, uint32_t* k_out,
                   uint8_t g2[64], uint8_t s_g2[64]);
int zk_params_write(zk_ctx* ctx, const uint8_t g2[64], const uint8_t s_g2[64], uint8_t* file, size_t cap, size_t* len);

/* SerdeFormat of the two halo2 stacks the reference links (A = halo2_proofs 0.2.0 / halo2curves 0.3.1,
 * C = halo2-axiom 0.4.2 / halo2curves-axiom, Cargo.lock:1181-1199):
 *   ZK_SERDE_PROCESSED      compressed, parity of y in bit 7 of byte 31, identity = zeros      (halo2curves 0.3.1)
 *   ZK_SERDE_PROCESSED_V04  compressed, parity of y in bit 6, identity flag in bit 7           (halo2curves >= 0.4; the
 *                           encoding of the commitments in bin/assets/proof.bin, bin/src/main.rs:270-279)
 *   ZK_SERDE_RAW_BYTES      x | y as Montgomery limbs (64 B) = zk_g1_affine; G2 elements take 128 B
 * zk_g1_encode / zk_g1_decode take host or device point arrays and host byte buffers; *_format are zk_params_read /
 * zk_params_write for a given format (g2 / s_g2: 64 B compressed, 128 B for RAW_BYTES). */
enum zk_serde_format { ZK_SERDE_PROCESSED = 0, ZK_SERDE_PROCESSED_V04 = 1, ZK_SERDE_RAW_BYTES = 2 };
int zk_g1_encode(zk_ctx* ctx, int format, const zk_g1_affine* points, size_t n, uint8_t* out);
int zk_g1_decode(zk_ctx* ctx, int format, const uint8_t* in, size_t n, zk_g1_affine* out);
int zk_params_read_format(zk_ctx* ctx, int format, const uint8_t* file, size_t len, int precompute_windows,
                          uint32_t* k_out, uint8_t* g2, uint8_t* s_g2);
int zk_params_write_format(zk_ctx* ctx, int format, const uint8_t* g2, const uint8_t* s_g2, uint8_t* file, size_t cap,
                           size_t* len);
/* G1Affine::to_bytes / from_bytes for n points (host or device pointers in, host out); decompression fails with
 * ZK_ERR_ARG on a non-canonical x or a point that is not on the curve. */
int zk_g1_compress(zk_ctx* ctx, const zk_g1_affine* points, size_t n, uint8_t* out);
int zk_g1_decompress(zk_ctx* ctx, const uint8_t* in, size_t n, zk_g1_affine* out);

/* ---- best_multiexp ------------------------------------------------------ */

typedef enum { ZK_BASES_HOST = 0, ZK_BASES_G = 1, ZK_BASES_G_LAGRANGE = 2 } zk_bases;

/* out = sum_i scalars[i] * bases[i]  (arithmetic.rs best_multiexp).
 * which == ZK_BASES_HOST : `bases` is a host array of n points;
 * otherwise `bases` is ignored and the first n resident SRS points are used
 * (ParamsKZG::commit / commit_lagrange).  `out` is normalised: z = 1, or the
 * identity (0, 1, 0). */
int zk_msm(zk_ctx* ctx, const zk_fr* scalars, const zk_g1_affine* bases,
           int which, size_t n, zk_g1* out);
/* Same with scalars (and, for ZK_BASES_HOST, bases) already in device memory
 * and the result left in device memory; asynchronous on the context stream. */
int zk_msm_dev(zk_ctx* ctx, const zk_fr* d_scalars, const zk_g1_affine* d_bases,
               int which, size_t n, zk_g1* d_out);
/* Batch of `count` commitments over the same resident base set: scalars is
 * count x n (row-major, device); d_out receives count points. */
int zk_msm_batch_dev(zk_ctx* ctx, const zk_fr* d_scalars, int which, size_t n,
                     size_t count, zk_g1* d_out);
/* out = sum of `count` Jacobian points (used to fold per-GPU partial MSMs). */
int zk_g1_sum(zk_ctx* ctx, const zk_g1* points, size_t count, zk_g1* out);
int zk_g1_sum_dev(zk_ctx* ctx, const zk_g1* d_points, size_t count, zk_g1* d_out);

/* ---- best_fft ------------------------------------------------------------ */

/* In-place radix-2 FFT over Fr, natural order in and out, a.len() = 2^log_n
 * (arithmetic.rs best_fft).  `omega` must be a primitive 2^log_n-th root. */
int zk_ntt(zk_ctx* ctx, zk_fr* a, const zk_fr* omega, uint32_t log_n);
int zk_ntt_dev(zk_ctx* ctx, zk_fr* d_a, const zk_fr* omega_host, uint32_t log_n);

/* ---- EvaluationDomain ---------------------------------------------------- */

/* EvaluationDomain::new(j, k) (poly/domain.rs): j = cs.degree(). */
int zk_domain_create(zk_ctx* ctx, uint32_t j, uint32_t k, zk_domain** out);
void zk_domain_destroy(zk_domain* d);
uint32_t zk_domain_k(const zk_domain* d);
uint32_t zk_domain_extended_k(const zk_domain* d);
/* Writes omega, omega_inv, extended_omega, extended_omega_inv, g_coset,
 * g_coset_inv, ifft_divisor, extended_ifft_divisor (8 elements). */
int zk_domain_constants(const zk_domain* d, zk_fr out[8]);

/* lagrange_to_coeff: a (n values) -> coefficients, in place. */
int zk_lagrange_to_coeff(zk_domain* d, zk_fr* a);
/* coeff_to_lagrange (fft with omega), in place -- halo2 `coeff_to_lagrange`. */
int zk_coeff_to_lagrange(zk_domain* d, zk_fr* a);
/* coeff_to_extended: n coefficients -> 2^extended_k coset evaluations. */
int zk_coeff_to_extended(zk_domain* d, const zk_fr* a, zk_fr* out_ext);
/* extended_to_coeff: 2^extended_k coset evaluations -> n*(j-1) coefficients. */
int zk_extended_to_coeff(zk_domain* d, const zk_fr* a_ext, zk_fr* out);
/* divide_by_vanishing_poly, in place on 2^extended_k values. */
int zk_divide_by_vanishing_poly(zk_domain* d, zk_fr* a_ext);
/* Device-pointer variants (asynchronous on the context stream).  `count`
 * independent polynomials stored back to back are transformed per call. */
int zk_lagrange_to_coeff_dev(zk_domain* d, zk_fr* d_a, size_t count);
int zk_coeff_to_lagrange_dev(zk_domain* d, zk_fr* d_a, size_t count);
int zk_coeff_to_extended_dev(zk_domain* d, const zk_fr* d_a, zk_fr* d_out_ext, size_t count);
int zk_extended_to_coeff_dev(zk_domain* d, zk_fr* d_a_ext /* clobbered */, zk_fr* d_out);
int zk_divide_by_vanishing_poly_dev(zk_domain* d, zk_fr* d_a_ext);

/* ---- Evaluator (plonk/evaluation.rs) -------------------------------------- */

/* ValueSource */
typedef enum {
  ZK_VS_CONSTANT = 0,      /* a = index into constants */
  ZK_VS_INTERMEDIATE = 1,  /* a = index into intermediates */
  ZK_VS_FIXED = 2,         /* a = column index, b = index into rotations */
  ZK_VS_ADVICE = 3,
  ZK_VS_INSTANCE = 4,
  ZK_VS_CHALLENGE = 5,     /* a = challenge index */
  ZK_VS_BETA = 6,
  ZK_VS_GAMMA = 7,
  ZK_VS_THETA = 8,
  ZK_VS_Y = 9,
  ZK_VS_PREVIOUS_VALUE = 10
} zk_vs_kind;
typedef struct { uint32_t kind, a, b; } zk_value_source;

/* Calculation */
typedef enum {
  ZK_CALC_ADD = 0, ZK_CALC_SUB = 1, ZK_CALC_MUL = 2, ZK_CALC_SQUARE = 3,
  ZK_CALC_DOUBLE = 4, ZK_CALC_NEGATE = 5,
  ZK_CALC_HORNER = 6,      /* a = start value, b = factor, parts = horner_parts[parts_off .. +parts_len] */
  ZK_CALC_STORE = 7
} zk_calc_kind;
/* CalculationInfo { calculation, target } */
typedef struct {
  uint32_t kind;
  zk_value_source a, b;
  uint32_t parts_off, parts_len;
  uint32_t target;           /* intermediate written by this calculation */
} zk_calculation;

/* GraphEvaluator { constants, rotations, calculations, num_intermediates } */
typedef struct {
  const zk_fr* constants;      uint32_t n_constants;
  const int32_t* rotations;    uint32_t n_rotations;
  const zk_calculation* calcs; uint32_t n_calcs;
  const zk_value_source* horner_parts; uint32_t n_horner_parts;
  uint32_t num_intermediates;
} zk_graph;

typedef enum { ZK_COL_ADVICE = 0, ZK_COL_FIXED = 1, ZK_COL_INSTANCE = 2 } zk_col_type;
typedef struct { uint32_t column_type, index; } zk_column;

/* What Evaluator::evaluate_h reads from ProvingKey / ConstraintSystem. */
typedef struct {
  uint32_t k;                  /* domain size 2^k */
  uint32_t degree;             /* cs.degree() */
  uint32_t blinding_factors;   /* cs.blinding_factors() */
  uint32_t n_fixed, n_advice, n_instance, n_challenges;
  zk_graph custom_gates;       /* Evaluator.custom_gates */
  const zk_graph* lookups;     /* Evaluator.lookups */
  uint32_t n_lookups;
  const zk_column* perm_columns; /* cs.permutation.columns, in order */
  uint32_t n_perm_columns;
} zk_evaluator_desc;

/* Upload the proving-key side of evaluate_h and keep it resident in HBM:
 * pk.fixed_cosets (n_fixed x ext), pk.permutation.cosets (n_perm_columns x
 * ext), pk.l0 / l_last / l_active_row (ext each), all in extended Lagrange
 * form exactly as ProvingKey stores them (host pointers; copied).
 * Alternatively pass the coefficient forms (n each) with coeff_form != 0 and
 * the cosets are produced on the device (l_active_row is then derived from
 * l_last and l_blind: pass l_blind's coefficients in its place). */
int zk_pk_upload(zk_ctx* ctx, const zk_evaluator_desc* desc,
                 const zk_fr* const* fixed, const zk_fr* const* perm,
                 const zk_fr* l0, const zk_fr* l_last, const zk_fr* l_active_row,
                 int coeff_form, zk_pk** out);
void zk_pk_destroy(zk_pk* pk);
zk_domain* zk_pk_domain(zk_pk* pk);
/* Number of field multiplications the compiled per-row program performs and
 * the number of distinct extended columns it reads (roofline bookkeeping). */
int zk_pk_program_stats(const zk_pk* pk, uint64_t* muls_per_row, uint64_t* columns_read);

/* evaluate_h is COMPILED from the circuit's expression graphs: zk_pk_upload (and zk_keygen through it) emits straight-line
 * CUDA C++ for the optimised program and compiles it with NVRTC for sm_100a; the word-stream interpreter remains as the
 * checked reference (ZK_EVALH=interp selects it; it also takes over when NVRTC is unavailable).
 *   zk_evalh_generate_source  the generated text (host_flavour != 0: the same statements as a plain C++ function, for CPU
 *                             tests); buf may be NULL to query the length.  Host only.
 *   zk_pk_evaluator_kind      1 = generated kernel in use, 0 = interpreter
 *   zk_pk_evaluator_log       why not ("" when it is); *compile_seconds = NVRTC time at key generation */
int zk_evalh_generate_source(const zk_evaluator_desc* desc, int host_flavour, char* buf, size_t cap, size_t* len);
int zk_pk_evaluator_kind(const zk_pk* pk);
const char* zk_pk_evaluator_log(const zk_pk* pk, double* compile_seconds);

/* Evaluator::evaluate_h for one circuit instance.  Polynomials are in
 * coefficient form (n each), host pointers:
 *   advice_polys[n_advice], instance_polys[n_instance], challenges[n_challenges],
 *   lookup_polys[3*n_lookups] = {product, permuted_input, permuted_table} per lookup,
 *   perm_product_polys[n_sets], n_sets = ceil(n_perm_columns / (degree-2)).
 * out_ext receives 2^extended_k values of h's numerator in extended Lagrange
 * form (before divide_by_vanishing_poly, like the Rust function). */
int zk_evaluate_h(zk_pk* pk,
                  const zk_fr* const* advice_polys, const zk_fr* const* instance_polys,
                  const zk_fr* challenges,
                  const zk_fr* y, const zk_fr* beta, const zk_fr* gamma, const zk_fr* theta,
                  const zk_fr* const* lookup_polys, const zk_fr* const* perm_product_polys,
                  zk_fr* out_ext);
/* Same with every polynomial already resident: the pointer tables are host
 * arrays of DEVICE pointers; d_out_ext is a device buffer. Asynchronous. */
int zk_evaluate_h_dev(zk_pk* pk,
                      const zk_fr* const* d_advice_polys, const zk_fr* const* d_instance_polys,
                      const zk_fr* challenges,
                      const zk_fr* y, const zk_fr* beta, const zk_fr* gamma, const zk_fr* theta,
                      const zk_fr* const* d_lookup_polys, const zk_fr* const* d_perm_product_polys,
                      zk_fr* d_out_ext);

/* ---- keygen_pk / create_proof (plonk/keygen.rs, plonk/prover.rs) -------------- */
/* the prover's n-sized loops between the MSM / FFT /
 * evaluate_h calls (lookup compression + permutation, grand products, polynomial
 * evaluation, SHPLONK quotients) run on the device too, so that one proof needs
 * one upload of the advice columns and one download of the proof bytes.
 * Replaces halo2_proofs::plonk::{keygen_vk, keygen_pk, create_proof} as called at
 * circuits/src/sgx_dcap_verifier.rs:803, :807, :814-822 for
 * KZGCommitmentScheme<Bn256> + ProverSHPLONK + Blake2bWrite/Challenge255, one
 * circuit instance, first-phase advice columns. */

typedef struct { uint32_t column; int32_t rotation; } zk_query;

typedef struct {
  zk_evaluator_desc ev;              /* Evaluator image + column counts + permutation columns */
  const zk_graph* lookup_inputs;     /* ev.n_lookups graphs: theta-compressed input expressions, */
  const zk_graph* lookup_tables;     /*   resp. table expressions, evaluated on the n Lagrange rows */
  const zk_query* advice_queries; uint32_t n_advice_queries;   /* cs.advice_queries, in order */
  const zk_query* fixed_queries;  uint32_t n_fixed_queries;    /* cs.fixed_queries, in order */
  zk_fr transcript_repr;             /* vk.transcript_repr: upstream hashes the Debug string of the pinned
                                        vk, which only Rust can produce -- supplied by the caller */
  const zk_query* instance_queries; uint32_t n_instance_queries;   /* cs.instance_queries (read by zk_verify_proof only) */
} zk_circuit_desc;

typedef struct zk_pkey zk_pkey;

/* keygen_vk + keygen_pk.  fixed_values: ev.n_fixed host columns of n Lagrange values (after selector
 * compression); perm_mapping: ev.n_perm_columns x n pairs (column', row') -- cell (i, j) of the i-th
 * permutation column is mapped to cell (column', row') by the copy-constraint cycles (identity when
 * unconstrained), as permutation::keygen::Assembly holds it.  The SRS must have been uploaded to ctx.
 * Everything create_proof needs stays resident in HBM. */
int zk_keygen(zk_ctx* ctx, const zk_circuit_desc* circuit, const zk_fr* const* fixed_values,
              const uint32_t* perm_mapping, zk_pkey** out);
void zk_pkey_destroy(zk_pkey* pk);
/* vk.fixed_commitments (n_fixed) and vk.permutation.commitments (n_perm_columns), affine. */
int zk_pkey_vk_commitments(const zk_pkey* pk, zk_g1_affine* fixed_commitments, zk_g1_affine* perm_commitments);
/* The evaluate_h side of the key (for zk_evaluate_h* / zk_pk_program_stats). */
zk_pk* zk_pkey_pk(zk_pkey* pk);
/* The key's second context (= its second stream: the witness polynomials' transforms overlap the commitments
 * there), for zk_ctx_timing* / zk_ctx_launch_count; NULL when ZK_PROVER_STREAMS=1.  Owned by the key. */
zk_ctx* zk_pkey_aux_ctx(zk_pkey* pk);
/* Transcript of the key's proofs: halo2's Blake2bWrite / Challenge255 (default; circuits/src/sgx_dcap_verifier.rs:813)
 * or snark-verifier's EvmTranscript (Keccak-256, uncompressed big-endian points; gen_evm_proof_shplonk at
 * crates/p256-ecdsa/src/base.rs:193-199), or snark-verifier's PoseidonTranscript over the native loader (T = 5, RATE = 4,
 * R_F = 8, R_P = 60; gen_proof at base.rs:200-211, the CLI's verifier at bin/src/main.rs:239-252): points absorbed as
 * (x mod r, y mod r) and written compressed in the halo2curves >= 0.4 encoding, challenges = the sponge output. */
typedef enum { ZK_TRANSCRIPT_BLAKE2B = 0, ZK_TRANSCRIPT_EVM = 1, ZK_TRANSCRIPT_POSEIDON = 2 } zk_transcript_kind;
int zk_pkey_set_transcript(zk_pkey* pk, int kind);
/* Number of Fr::random draws one create_proof makes for this circuit. */
size_t zk_proof_randoms(const zk_pkey* pk);
/* Upper bound of the proof size in bytes. */
size_t zk_proof_max_bytes(const zk_pkey* pk);

/* create_proof.  advice_values: ev.n_advice host columns of n Lagrange values as `synthesize` left
 * them (the last blinding_factors+1 rows are overwritten with randoms here, as upstream);
 * instance_values / instance_lens: ev.n_instance public-input columns; randoms: the caller's RNG
 * pre-drawn -- element i is the i-th Fr::random(&mut rng) upstream would make (zk_proof_randoms of them).
 * proof receives the transcript bytes (Blake2bWrite::finalize).  The advice columns and `randoms` may also be
 * DEVICE pointers (operands already resident in HBM); instance values must be host memory. */
int zk_create_proof(zk_pkey* pk, const zk_fr* const* advice_values,
                    const zk_fr* const* instance_values, const uint32_t* instance_lens,
                    const zk_fr* randoms, size_t n_randoms,
                    uint8_t* proof, size_t proof_cap, size_t* proof_len);

/* ---- verify_proof (plonk/verifier.rs, VerifierSHPLONK, SingleStrategy) -------- */
/* the check the reference makes at circuits/src/sgx_dcap_verifier.rs:827-843 and
 * bin/src/main.rs:244-251, on the HOST (a few hundred field operations per column, ~150 scalar multiplications and
 * two BN254 optimal-ate pairings; no device needed).  g2 / s_g2: ParamsKZG.g2 and .s_g2 as 128 uncompressed bytes
 * (x.c0, x.c1, y.c0, y.c1 canonical little-endian); g1 = params.g[0]. */
typedef struct {
  const zk_g1_affine* fixed_commitments;        /* vk.fixed_commitments, ev.n_fixed */
  const zk_g1_affine* permutation_commitments;  /* vk.permutation.commitments, ev.n_perm_columns */
  zk_g1_affine g1;
  const uint8_t* g2;
  const uint8_t* s_g2;
  int transcript_kind;                          /* zk_transcript_kind the proof was written with */
} zk_verifier_params;
/* ParamsKZG.g2 / s_g2 between the SRS file encodings (zk_params_read* returns them as the file holds them: 64 compressed
 * bytes for the PROCESSED formats, 128 bytes of Montgomery limbs for RAW_BYTES) and the 128 uncompressed canonical bytes
 * zk_verify_proof consumes.  Host code (one Fq2 square root).  Errors: encodings off the curve / not canonical. */
/* The Poseidon permutation behind ZK_TRANSCRIPT_POSEIDON (host code; parameters by the Grain LFSR of the Poseidon paper for
 * width t <= 16, r_f full and r_p partial rounds over bn256::Fr), in place on `state[t]`; and the transcript's sponge (snark-verifier
 * `Poseidon::{update, squeeze}`, T = 5, RATE = 4): absorbs `count` elements and writes `squeezes` (>= 1) successive outputs. */
int zk_poseidon_permute(zk_fr* state, uint32_t t, uint32_t r_f, uint32_t r_p);
int zk_poseidon_hash(const zk_fr* elements, size_t count, zk_fr* out, uint32_t squeezes);
int zk_g2_decode(int format, const uint8_t* in, uint8_t out_uncompressed[128]);
int zk_g2_encode(int format, const uint8_t in_uncompressed[128], uint8_t* out);

/* *accepted = 1 when the proof verifies, 0 otherwise (malformed proofs are rejected, not errors). */
int zk_verify_proof(const zk_circuit_desc* circuit, const zk_verifier_params* vp,
                    const zk_fr* const* instance_values, const uint32_t* instance_lens,
                    const uint8_t* proof, size_t proof_len, int* accepted);

/* ---- one proof over several GPUs ------------------------- */
/* One process per GPU, as torch.distributed launches them.  Every rank creates its own context, SRS and key from the
 * SAME inputs and calls zk_create_proof with the SAME witness and randoms; inside, the independent column commitments
 * and the witness polynomials' transforms are split over the ranks (a single commitment: by base range), h(X) by row
 * range, and the results are all-gathered
 * over NVLink (NCCL, bound at run time), so that every rank ends with the same proof bytes.  The reference has no
 * counterpart: halo2 scales only with rayon threads on one host. */
typedef struct zk_group zk_group;
/* rank 0 obtains the rendezvous id (ncclGetUniqueId) and distributes it to the other ranks out of band */
int zk_group_unique_id(uint8_t id[256]);
int zk_group_create(zk_ctx* ctx, uint32_t rank, uint32_t world, const uint8_t id[256], zk_group** out);
void zk_group_destroy(zk_group* g);
uint32_t zk_group_rank(const zk_group* g);
uint32_t zk_group_world(const zk_group* g);
/* 1 when the small exchanges of this group (commitment points, status words) go through peer-mapped windows: stores over
 * NVLink issued from the producing kernel's epilogue plus a flag, instead of NCCL all-gathers (ZK_GROUP_NO_P2P=1 or a
 * failed CUDA IPC mapping leave NCCL in charge) */
int zk_group_p2p_enabled(const zk_group* g);
/* Attach (or detach with NULL) a group: the key's proofs are then computed jointly.  The group must outlive its use. */
int zk_pkey_set_group(zk_pkey* pk, zk_group* g);
/* ONE MSM over the group's GPUs, sharded by contiguous base range (`zk_msm_sharded` of 8b): every rank has
 * uploaded ITS slice of the bases as the context's resident SRS (zk_srs_upload*, window table per slice) and passes its slice
 * of the scalars (device memory).  The partial points are exchanged by the MSM's own last kernel (stores into the peers'
 * windows over NVLink + a flag; NCCL all-gather when the windows are unavailable) and summed on every rank:
 * d_out = one normalised G1 (device), identical everywhere. */
int zk_msm_sharded_dev(zk_group* g, const zk_fr* d_scalars, int which, size_t n_local, zk_g1* d_out);
/* How `count` commitments over n bases are split: *mode 0 = rank commits columns [*lo, *hi) (every rank's share padded
 * to *per columns), *mode 1 = every rank commits all of them over the bases [*lo, *hi).  Pure host function. */
/* Cost-aware plan of a batch of `count` >= world column commitments: owner[i] = rank of column i (longest first onto the
 * least loaded rank, at most ceil(count / world) columns per rank; host logic). */
void zk_group_balance_plan(const uint32_t* cost, size_t count, uint32_t world, uint32_t* owner);
void zk_group_shard_plan(uint32_t rank, uint32_t world, size_t count, size_t n, int* mode, size_t* per, size_t* lo, size_t* hi);

/* ---- element-wise field arithmetic (parity self-test of the device field) --- */
/* op: 0 add, 1 sub, 2 mul, 3 neg, 4 invert, 5 square, 6 to Montgomery form,
 * 7 from Montgomery form (to_repr), 8 double.  field: 0 = Fr, 1 = Fq.
 * a, b (b may be NULL for unary ops) and out are host arrays of n elements. */
int zk_field_op(zk_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n);

/* ---- micro-benchmarks (roofline denominators measured on the box) -------- */

/* Runs a register-resident chain of `iters` Montgomery multiplications per
 * thread over the whole chip and returns 1e9 field multiplications / second. */
int zk_bench_modmul(zk_ctx* ctx, int field /*0 Fr, 1 Fq*/, uint32_t iters, double* gmul_per_s);
/* Raw 32-bit integer multiply-add throughput, in 1e12 IMAD / s; kind 0 =
 * mad.lo.u32 chains, 1 = mad.wide.u32 chains. */
int zk_bench_imad(zk_ctx* ctx, int kind, uint32_t iters, double* timad_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ZKDCAP_B200_H */
