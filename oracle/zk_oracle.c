/*
 * zk_oracle.c -- TEST INFRASTRUCTURE: CPU oracle for the halo2 hot path.
 *
 * PARITY UNPINNED: the algorithms restated here live in third-party crates
 * that are not vendored in /root/reference and cannot be built in this image
 * (no cargo/rustc, no network):
 *     halo2_proofs 0.2.0  git zkwebauthn/halo2      @ c254c75f (Cargo.lock:1314-1316)
 *     halo2curves  0.3.1  git zkwebauthn/halo2curves @ bdb2e661 (Cargo.lock:1329-1331)
 * The reference tree holds no golden vector for MSM / FFT / evaluate_h
 * (SURVEY.md section 4, 8c).  This file is pinned instead by (1) the published
 * BN254 constants, (2) the EIP-196 known point 2*(1,2), (3) algebraic
 * invariants, and (4) an independent Python big-integer model
 * (oracle/pymodel.py); see tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may
 * load this library.  Nothing in the product path links or calls it.
 *
 * Each function cites the upstream file it follows ([UPSTREAM] = restated from
 * the pinned revision's published source, not from a file on disk) and the
 * reference call site that reaches it.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "../include/zkdcap_b200.h"   /* plain-data struct definitions only */

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define FNAME fr
#define FMOD {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}
#include "field_tmpl.h"
#define FNAME fq
#define FMOD {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}
#include "field_tmpl.h"

/* ------------------------------------------------------------------------ */
/* constants (halo2curves bn256/fr.rs: GENERATOR = 7, S = 28)                */

static fr FR_ROOT_OF_UNITY, FR_ROOT_OF_UNITY_INV, FR_ZETA, FR_DELTA;
static fq FQ_B3;   /* curve constant b = 3 */
static int g_inited = 0;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void do_init(void) {
  fr_init(); fq_init();
  fr g = fr_from_u64(7);
  /* ROOT_OF_UNITY = 7^((r-1)/2^28) */
  u64 e[4]; memcpy(e, fr_MOD, 32); e[0] -= 1;
  /* shift right by 28 */
  for (int i = 0; i < 4; i++) e[i] = (e[i] >> 28) | (i < 3 ? (e[i + 1] << 36) : 0);
  FR_ROOT_OF_UNITY = fr_pow(g, e);
  FR_ROOT_OF_UNITY_INV = fr_inv(FR_ROOT_OF_UNITY);
  /* DELTA = 7^(2^28) */
  FR_DELTA = fr_pow_u64(g, 1ULL << 28);
  /* ZETA: the constant published in bn256/fr.rs */
  static const u64 zeta_raw[4] = {0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL};
  FR_ZETA = fr_from_raw(zeta_raw);
  FQ_B3 = fq_from_u64(3);
  g_inited = 1;
}
static inline void ensure_init(void) { pthread_once(&g_once, do_init); }

void orc_init(void) { ensure_init(); }

/* out[0..5] = ROOT_OF_UNITY, ROOT_OF_UNITY_INV, ZETA, DELTA, R (one), R2 -- all as stored (Montgomery) */
void orc_fr_constants(zk_fr out[6]) {
  ensure_init();
  memcpy(&out[0], &FR_ROOT_OF_UNITY, 32); memcpy(&out[1], &FR_ROOT_OF_UNITY_INV, 32);
  memcpy(&out[2], &FR_ZETA, 32); memcpy(&out[3], &FR_DELTA, 32);
  memcpy(&out[4], &fr_R1, 32); memcpy(&out[5], &fr_R2, 32);
}
void orc_fq_constants(zk_fq out[3]) {
  ensure_init();
  memcpy(&out[0], &fq_R1, 32); memcpy(&out[1], &fq_R2, 32); memcpy(&out[2], &fq_R3, 32);
}
u64 orc_fr_inv64(void) { ensure_init(); return fr_INV; }
u64 orc_fq_inv64(void) { ensure_init(); return fq_INV; }

/* ------------------------------------------------------------------------ */
/* element-wise field API for tests (field: 0 = Fr, 1 = Fq)                  */
/* op: 0 add 1 sub 2 mul 3 neg 4 inv 5 sqr 6 from_raw 7 to_raw 8 double      */
void orc_field_op(int field, int op, const u64* a, const u64* b, u64* out, size_t n) {
  ensure_init();
  for (size_t i = 0; i < n; i++) {
    if (field == 0) {
      fr x, y = fr_ZERO_, z; memcpy(&x, a + 4 * i, 32); if (b) memcpy(&y, b + 4 * i, 32);
      switch (op) {
        case 0: z = fr_add(x, y); break; case 1: z = fr_sub(x, y); break;
        case 2: z = fr_mul(x, y); break; case 3: z = fr_neg(x); break;
        case 4: z = fr_inv(x); break; case 5: z = fr_sqr(x); break;
        case 6: z = fr_from_raw(x.l); break; case 7: fr_to_raw(x, z.l); break;
        default: z = fr_dbl(x); break;
      }
      memcpy(out + 4 * i, &z, 32);
    } else {
      fq x, y = fq_ZERO_, z; memcpy(&x, a + 4 * i, 32); if (b) memcpy(&y, b + 4 * i, 32);
      switch (op) {
        case 0: z = fq_add(x, y); break; case 1: z = fq_sub(x, y); break;
        case 2: z = fq_mul(x, y); break; case 3: z = fq_neg(x); break;
        case 4: z = fq_inv(x); break; case 5: z = fq_sqr(x); break;
        case 6: z = fq_from_raw(x.l); break; case 7: fq_to_raw(x, z.l); break;
        default: z = fq_dbl(x); break;
      }
      memcpy(out + 4 * i, &z, 32);
    }
  }
}
/* Fr::from_u512 / from_bytes_wide over n 64-byte inputs */
void orc_fr_from_u512(const u64* wide, u64* out, size_t n) {
  ensure_init();
  for (size_t i = 0; i < n; i++) { fr z = fr_from_u512(wide + 8 * i); memcpy(out + 4 * i, &z, 32); }
}

/* ------------------------------------------------------------------------ */
/* G1: y^2 = x^3 + 3 over Fq (halo2curves bn256/curve.rs, derive/curve.rs)    */

typedef struct { fq x, y; } g1a;
typedef struct { fq x, y, z; } g1j;

static inline int g1a_is_identity(const g1a* p) { return fq_is_zero(p->x) && fq_is_zero(p->y); }
static inline int g1j_is_identity(const g1j* p) { return fq_is_zero(p->z); }
static inline g1j g1j_identity(void) { g1j r; r.x = fq_ZERO_; r.y = fq_R1; r.z = fq_ZERO_; return r; }
static inline g1j g1a_to_j(const g1a* p) {
  if (g1a_is_identity(p)) return g1j_identity();
  g1j r; r.x = p->x; r.y = p->y; r.z = fq_R1; return r;
}
static g1j g1j_double(const g1j* p) {
  if (g1j_is_identity(p)) return g1j_identity();
  fq a = fq_sqr(p->x), b = fq_sqr(p->y), c = fq_sqr(b);
  fq d = fq_sub(fq_sub(fq_sqr(fq_add(p->x, b)), a), c); d = fq_dbl(d);
  fq e = fq_add(fq_dbl(a), a), f = fq_sqr(e);
  g1j r;
  r.z = fq_dbl(fq_mul(p->y, p->z));
  r.x = fq_sub(f, fq_dbl(d));
  fq c8 = fq_dbl(fq_dbl(fq_dbl(c)));
  r.y = fq_sub(fq_mul(e, fq_sub(d, r.x)), c8);
  return r;
}
static g1j g1j_add(const g1j* p, const g1j* q) {
  if (g1j_is_identity(p)) return *q;
  if (g1j_is_identity(q)) return *p;
  fq z1z1 = fq_sqr(p->z), z2z2 = fq_sqr(q->z);
  fq u1 = fq_mul(p->x, z2z2), u2 = fq_mul(q->x, z1z1);
  fq s1 = fq_mul(fq_mul(p->y, z2z2), q->z), s2 = fq_mul(fq_mul(q->y, z1z1), p->z);
  if (fq_eq(u1, u2)) {
    if (fq_eq(s1, s2)) return g1j_double(p);
    return g1j_identity();
  }
  fq h = fq_sub(u2, u1), i = fq_sqr(fq_dbl(h)), j = fq_mul(h, i);
  fq rr = fq_dbl(fq_sub(s2, s1)), v = fq_mul(u1, i);
  g1j r;
  r.x = fq_sub(fq_sub(fq_sqr(rr), j), fq_dbl(v));
  r.y = fq_sub(fq_mul(rr, fq_sub(v, r.x)), fq_dbl(fq_mul(s1, j)));
  r.z = fq_mul(fq_sub(fq_sub(fq_sqr(fq_add(p->z, q->z)), z1z1), z2z2), h);
  return r;
}
static g1j g1j_add_mixed(const g1j* p, const g1a* q) {
  if (g1a_is_identity(q)) return *p;
  if (g1j_is_identity(p)) return g1a_to_j(q);
  fq z1z1 = fq_sqr(p->z);
  fq u2 = fq_mul(q->x, z1z1), s2 = fq_mul(fq_mul(q->y, z1z1), p->z);
  if (fq_eq(p->x, u2)) {
    if (fq_eq(p->y, s2)) return g1j_double(p);
    return g1j_identity();
  }
  fq h = fq_sub(u2, p->x), hh = fq_sqr(h), i = fq_dbl(fq_dbl(hh)), j = fq_mul(h, i);
  fq rr = fq_dbl(fq_sub(s2, p->y)), v = fq_mul(p->x, i);
  g1j r;
  r.x = fq_sub(fq_sub(fq_sqr(rr), j), fq_dbl(v));
  r.y = fq_sub(fq_mul(rr, fq_sub(v, r.x)), fq_dbl(fq_mul(p->y, j)));
  r.z = fq_sub(fq_sub(fq_sqr(fq_add(p->z, h)), z1z1), hh);
  return r;
}
static g1a g1j_to_affine(const g1j* p) {
  g1a r;
  if (g1j_is_identity(p)) { r.x = fq_ZERO_; r.y = fq_ZERO_; return r; }
  fq zi = fq_inv(p->z), zi2 = fq_sqr(zi);
  r.x = fq_mul(p->x, zi2);
  r.y = fq_mul(p->y, fq_mul(zi2, zi));
  return r;
}
static int g1a_on_curve(const g1a* p) {
  if (g1a_is_identity(p)) return 1;
  fq lhs = fq_sqr(p->y), rhs = fq_add(fq_mul(fq_sqr(p->x), p->x), FQ_B3);
  return fq_eq(lhs, rhs);
}
/* plain double-and-add, scalar given in Montgomery form */
static g1j g1_scalar_mul(const g1a* base, fr scalar) {
  u64 e[4]; fr_to_raw(scalar, e);
  g1j acc = g1j_identity();
  for (int i = 255; i >= 0; i--) {
    acc = g1j_double(&acc);
    if ((e[i / 64] >> (i % 64)) & 1) acc = g1j_add_mixed(&acc, base);
  }
  return acc;
}

/* exported curve helpers */
void orc_g1_add(const zk_g1* a, const zk_g1* b, zk_g1* out) {
  ensure_init(); g1j r = g1j_add((const g1j*)a, (const g1j*)b); memcpy(out, &r, 96);
}
void orc_g1_double(const zk_g1* a, zk_g1* out) { ensure_init(); g1j r = g1j_double((const g1j*)a); memcpy(out, &r, 96); }
void orc_g1_add_mixed(const zk_g1* a, const zk_g1_affine* b, zk_g1* out) {
  ensure_init(); g1j r = g1j_add_mixed((const g1j*)a, (const g1a*)b); memcpy(out, &r, 96);
}
void orc_g1_to_affine(const zk_g1* a, zk_g1_affine* out, size_t n) {
  ensure_init();
  for (size_t i = 0; i < n; i++) { g1a r = g1j_to_affine((const g1j*)&a[i]); memcpy(&out[i], &r, 64); }
}
int orc_g1_on_curve(const zk_g1_affine* a, size_t n) {
  ensure_init();
  for (size_t i = 0; i < n; i++) if (!g1a_on_curve((const g1a*)&a[i])) return 0;
  return 1;
}
void orc_g1_mul(const zk_g1_affine* base, const zk_fr* scalar, zk_g1* out) {
  ensure_init(); fr s; memcpy(&s, scalar, 32);
  g1j r = g1_scalar_mul((const g1a*)base, s); memcpy(out, &r, 96);
}
/* naive MSM: sum of independent double-and-add products (the definition) */
void orc_msm_naive(const zk_fr* scalars, const zk_g1_affine* bases, size_t n, zk_g1* out) {
  ensure_init();
  g1j acc = g1j_identity();
  for (size_t i = 0; i < n; i++) {
    fr s; memcpy(&s, &scalars[i], 32);
    g1j t = g1_scalar_mul((const g1a*)&bases[i], s);
    acc = g1j_add(&acc, &t);
  }
  memcpy(out, &acc, 96);
}

/* ------------------------------------------------------------------------ */
/* thread helper                                                             */

typedef void (*range_fn)(size_t lo, size_t hi, int tid, void* arg);
typedef struct { range_fn fn; size_t lo, hi; int tid; void* arg; } range_job;
static void* range_tramp(void* p) { range_job* j = (range_job*)p; j->fn(j->lo, j->hi, j->tid, j->arg); return NULL; }
/* halo2 `parallelize`: split [0,n) in `threads` contiguous chunks */
static void parallel_for(size_t n, int threads, range_fn fn, void* arg) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  if (threads == 1) { fn(0, n, 0, arg); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  range_job* jobs = (range_job*)malloc(sizeof(range_job) * threads);
  size_t chunk = (n + threads - 1) / threads;
  int used = 0;
  for (int t = 0; t < threads; t++) {
    size_t lo = (size_t)t * chunk, hi = lo + chunk > n ? n : lo + chunk;
    if (lo >= hi) break;
    jobs[t] = (range_job){fn, lo, hi, t, arg};
    pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
    used++;
  }
  for (int t = 0; t < used; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

/* ------------------------------------------------------------------------ */
/* best_multiexp  [UPSTREAM halo2_proofs/src/arithmetic.rs: multiexp_serial,  */
/* best_multiexp; reached from create_proof at                               */
/* circuits/src/sgx_dcap_verifier.rs:814 via ParamsKZG::commit(_lagrange)]    */

static size_t get_at(size_t segment, size_t c, const unsigned char bytes[32]) {
  size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
  if (skip_bytes >= 32) return 0;
  unsigned char v[8] = {0};
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
  u64 tmp; memcpy(&tmp, v, 8);           /* little-endian host */
  tmp >>= skip_bits - skip_bytes * 8;
  tmp = tmp % (1ULL << c);
  return (size_t)tmp;
}

static void multiexp_serial(const fr* coeffs, const g1a* bases, size_t n, g1j* acc) {
  unsigned char (*reprs)[32] = malloc(n * 32);
  for (size_t i = 0; i < n; i++) { u64 t[4]; fr_to_raw(coeffs[i], t); memcpy(reprs[i], t, 32); }
  size_t c;
  if (n < 4) c = 1; else if (n < 32) c = 3; else c = (size_t)ceil(log((double)n));
  size_t segments = 256 / c + 1;
  size_t nb = ((size_t)1 << c) - 1;
  /* Bucket::{None, Affine, Projective} */
  unsigned char* state = malloc(nb);
  g1j* bk = malloc(nb * sizeof(g1j));
  for (size_t seg = segments; seg-- > 0;) {
    for (size_t i = 0; i < c; i++) *acc = g1j_double(acc);
    memset(state, 0, nb);
    for (size_t i = 0; i < n; i++) {
      size_t d = get_at(seg, c, reprs[i]);
      if (!d) continue;
      size_t b = d - 1;
      if (state[b] == 0) { bk[b].x = bases[i].x; bk[b].y = bases[i].y; state[b] = 1; }
      else if (state[b] == 1) {
        g1a prev = {bk[b].x, bk[b].y};
        g1j pj = g1a_to_j(&prev);
        bk[b] = g1j_add_mixed(&pj, &bases[i]);    /* affine + affine -> projective */
        state[b] = 2;
      } else bk[b] = g1j_add_mixed(&bk[b], &bases[i]);
    }
    /* summation by parts */
    g1j running = g1j_identity();
    for (size_t b = nb; b-- > 0;) {
      if (state[b] == 1) { g1a a = {bk[b].x, bk[b].y}; running = g1j_add_mixed(&running, &a); }
      else if (state[b] == 2) running = g1j_add(&running, &bk[b]);
      *acc = g1j_add(acc, &running);
    }
  }
  free(state); free(bk); free(reprs);
}

typedef struct { const fr* coeffs; const g1a* bases; size_t n, chunk; g1j* results; } msm_job;
static void msm_chunk_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; msm_job* j = (msm_job*)arg;
  for (size_t ci = lo; ci < hi; ci++) {
    size_t a = ci * j->chunk, b = a + j->chunk > j->n ? j->n : a + j->chunk;
    j->results[ci] = g1j_identity();
    multiexp_serial(j->coeffs + a, j->bases + a, b - a, &j->results[ci]);
  }
}
/* `threads` plays the role of rayon's current_num_threads() */
void orc_best_multiexp(const zk_fr* coeffs, const zk_g1_affine* bases, size_t n, int threads, zk_g1* out) {
  ensure_init();
  if (threads < 1) threads = 1;
  g1j acc = g1j_identity();
  if (n > (size_t)threads) {
    size_t chunk = n / threads;
    size_t num_chunks = (n + chunk - 1) / chunk;
    g1j* results = malloc(num_chunks * sizeof(g1j));
    msm_job job = {(const fr*)coeffs, (const g1a*)bases, n, chunk, results};
    parallel_for(num_chunks, (int)num_chunks, msm_chunk_fn, &job);
    for (size_t i = 0; i < num_chunks; i++) acc = g1j_add(&acc, &results[i]);
    free(results);
  } else {
    multiexp_serial((const fr*)coeffs, (const g1a*)bases, n, &acc);
  }
  memcpy(out, &acc, 96);
}

/* ------------------------------------------------------------------------ */
/* best_fft  [UPSTREAM arithmetic.rs: best_fft, recursive_butterfly_arithmetic,*/
/* bitreverse; reached through EvaluationDomain from keygen_pk (:807) and     */
/* create_proof (:814)]                                                      */

static inline u64 bitreverse(u64 n, unsigned l) {
  u64 r = 0;
  for (unsigned i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
  return r;
}
static int log2_floor(int x) { int l = 0; while ((1 << (l + 1)) <= x) l++; return l; }

typedef struct { fr* a; size_t n; size_t twiddle_chunk; const fr* tw; int depth; } rb_job;
static void recursive_butterfly(fr* a, size_t n, size_t twiddle_chunk, const fr* tw, int par_depth);
static void* rb_tramp(void* p) { rb_job* j = (rb_job*)p; recursive_butterfly(j->a, j->n, j->twiddle_chunk, j->tw, j->depth); return NULL; }
static void recursive_butterfly(fr* a, size_t n, size_t twiddle_chunk, const fr* tw, int par_depth) {
  if (n == 2) {
    fr t = a[1]; a[1] = a[0]; a[0] = fr_add(a[0], t); a[1] = fr_sub(a[1], t);
    return;
  }
  fr* left = a; fr* right = a + n / 2;
  if (par_depth > 0 && n >= 1024) {       /* rayon::join */
    pthread_t th; rb_job j = {left, n / 2, twiddle_chunk * 2, tw, par_depth - 1};
    pthread_create(&th, NULL, rb_tramp, &j);
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, par_depth - 1);
    pthread_join(th, NULL);
  } else {
    recursive_butterfly(left, n / 2, twiddle_chunk * 2, tw, 0);
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, 0);
  }
  /* case when twiddle factor is one */
  fr t = right[0]; right[0] = left[0]; left[0] = fr_add(left[0], t); right[0] = fr_sub(right[0], t);
  for (size_t i = 1; i < n / 2; i++) {
    fr tt = fr_mul(right[i], tw[i * twiddle_chunk]);
    right[i] = left[i]; left[i] = fr_add(left[i], tt); right[i] = fr_sub(right[i], tt);
  }
}
typedef struct { fr* tw; fr omega; } tw_job;
static void tw_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; tw_job* j = (tw_job*)arg;
  fr w = fr_pow_u64(j->omega, (u64)lo);
  for (size_t i = lo; i < hi; i++) { j->tw[i] = w; w = fr_mul(w, j->omega); }
}
static void best_fft(fr* a, fr omega, unsigned log_n, int threads) {
  size_t n = (size_t)1 << log_n;
  if (threads < 1) threads = 1;
  int log_threads = log2_floor(threads);
  for (size_t k = 0; k < n; k++) {
    size_t rk = bitreverse(k, log_n);
    if (k < rk) { fr t = a[rk]; a[rk] = a[k]; a[k] = t; }
  }
  if (n == 1) return;
  /* precompute twiddle factors */
  fr* tw = malloc((n / 2 ? n / 2 : 1) * sizeof(fr));
  tw_job tj = {tw, omega};
  parallel_for(n / 2, threads, tw_fn, &tj);
  if (log_n <= (unsigned)log_threads) {
    size_t chunk = 2, twiddle_chunk = n / 2;
    for (unsigned l = 0; l < log_n; l++) {
      for (size_t base = 0; base < n; base += chunk) {
        fr* left = a + base; fr* right = a + base + chunk / 2;
        fr t = right[0]; right[0] = left[0]; left[0] = fr_add(left[0], t); right[0] = fr_sub(right[0], t);
        for (size_t i = 1; i < chunk / 2; i++) {
          fr tt = fr_mul(right[i], tw[i * twiddle_chunk]);
          right[i] = left[i]; left[i] = fr_add(left[i], tt); right[i] = fr_sub(right[i], tt);
        }
      }
      chunk *= 2; twiddle_chunk /= 2;
    }
  } else {
    recursive_butterfly(a, n, 1, tw, log_threads);
  }
  free(tw);
}
void orc_best_fft(zk_fr* a, const zk_fr* omega, uint32_t log_n, int threads) {
  ensure_init(); fr w; memcpy(&w, omega, 32);
  best_fft((fr*)a, w, log_n, threads);
}
/* O(n^2) DFT straight from the definition: out[i] = sum_j a[j] * omega^(i*j) */
void orc_dft_naive(const zk_fr* a, const zk_fr* omega, uint32_t log_n, zk_fr* out) {
  ensure_init();
  size_t n = (size_t)1 << log_n; fr w; memcpy(&w, omega, 32);
  const fr* in = (const fr*)a;
  fr wi = fr_R1;
  for (size_t i = 0; i < n; i++) {
    fr acc = fr_ZERO_, x = fr_R1;
    for (size_t j = 0; j < n; j++) { acc = fr_add(acc, fr_mul(in[j], x)); x = fr_mul(x, wi); }
    memcpy(&out[i], &acc, 32);
    wi = fr_mul(wi, w);
  }
}
/* eval_polynomial (Horner), arithmetic.rs */
void orc_eval_polynomial(const zk_fr* poly, size_t n, const zk_fr* point, zk_fr* out) {
  ensure_init(); fr x; memcpy(&x, point, 32);
  fr acc = fr_ZERO_;
  for (size_t i = n; i-- > 0;) { fr c; memcpy(&c, &poly[i], 32); acc = fr_add(fr_mul(acc, x), c); }
  memcpy(out, &acc, 32);
}

/* ------------------------------------------------------------------------ */
/* EvaluationDomain  [UPSTREAM halo2_proofs/src/poly/domain.rs]               */

typedef struct {
  uint32_t k, extended_k, quotient_poly_degree;
  size_t n, ext_n;
  fr omega, omega_inv, extended_omega, extended_omega_inv;
  fr g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor, barycentric_weight;
  fr* t_evaluations; size_t n_t;
  int threads;
} orc_domain;

orc_domain* orc_domain_new(uint32_t j, uint32_t k, int threads) {
  ensure_init();
  orc_domain* d = calloc(1, sizeof(orc_domain));
  d->threads = threads < 1 ? 1 : threads;
  d->quotient_poly_degree = j - 1;
  d->k = k; d->n = (size_t)1 << k;
  uint32_t ek = k;
  while (((size_t)1 << ek) < d->n * d->quotient_poly_degree) ek++;
  d->extended_k = ek; d->ext_n = (size_t)1 << ek;
  fr eo = FR_ROOT_OF_UNITY;
  for (uint32_t i = ek; i < 28; i++) eo = fr_sqr(eo);
  d->extended_omega = eo;
  d->extended_omega_inv = fr_inv(eo);
  fr o = eo, oi = d->extended_omega_inv;
  for (uint32_t i = k; i < ek; i++) { o = fr_sqr(o); oi = fr_sqr(oi); }
  d->omega = o; d->omega_inv = oi;
  d->g_coset = FR_ZETA; d->g_coset_inv = fr_sqr(FR_ZETA);
  d->n_t = (size_t)1 << (ek - k);
  d->t_evaluations = malloc(d->n_t * sizeof(fr));
  fr orig = fr_pow_u64(FR_ZETA, (u64)d->n), step = fr_pow_u64(eo, (u64)d->n), cur = orig;
  size_t cnt = 0;
  do { d->t_evaluations[cnt++] = cur; cur = fr_mul(cur, step); } while (!fr_eq(cur, orig) && cnt < d->n_t);
  for (size_t i = 0; i < d->n_t; i++) d->t_evaluations[i] = fr_sub(d->t_evaluations[i], fr_R1);
  fr_batch_invert(d->t_evaluations, d->n_t);
  d->ifft_divisor = fr_inv(fr_from_u64((u64)1 << k));
  d->extended_ifft_divisor = fr_inv(fr_from_u64((u64)1 << ek));
  d->barycentric_weight = d->ifft_divisor;
  return d;
}
void orc_domain_free(orc_domain* d) { if (d) { free(d->t_evaluations); free(d); } }
uint32_t orc_domain_extended_k(const orc_domain* d) { return d->extended_k; }
void orc_domain_constants(const orc_domain* d, zk_fr out[8]) {
  const fr* src[8] = {&d->omega, &d->omega_inv, &d->extended_omega, &d->extended_omega_inv,
                      &d->g_coset, &d->g_coset_inv, &d->ifft_divisor, &d->extended_ifft_divisor};
  for (int i = 0; i < 8; i++) memcpy(&out[i], src[i], 32);
}
void orc_domain_t_evaluations(const orc_domain* d, zk_fr* out) { memcpy(out, d->t_evaluations, d->n_t * 32); }

typedef struct { fr* a; fr s; } scale_job;
static void scale_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; scale_job* j = (scale_job*)arg;
  for (size_t i = lo; i < hi; i++) j->a[i] = fr_mul(j->a[i], j->s);
}
static void domain_ifft(const orc_domain* d, fr* a, fr omega_inv, unsigned log_n, fr divisor) {
  best_fft(a, omega_inv, log_n, d->threads);
  scale_job sj = {a, divisor};
  parallel_for((size_t)1 << log_n, d->threads, scale_fn, &sj);
}
static void distribute_powers_zeta(const orc_domain* d, fr* a, size_t len, int into_coset) {
  fr cp[2];
  if (into_coset) { cp[0] = d->g_coset; cp[1] = d->g_coset_inv; } else { cp[0] = d->g_coset_inv; cp[1] = d->g_coset; }
  for (size_t i = 0; i < len; i++) { size_t j = i % 3; if (j) a[i] = fr_mul(a[i], cp[j - 1]); }
}
void orc_lagrange_to_coeff(const orc_domain* d, zk_fr* a) { domain_ifft(d, (fr*)a, d->omega_inv, d->k, d->ifft_divisor); }
void orc_coeff_to_lagrange(const orc_domain* d, zk_fr* a) { best_fft((fr*)a, d->omega, d->k, d->threads); }
void orc_coeff_to_extended(const orc_domain* d, const zk_fr* a, zk_fr* out_ext) {
  fr* o = (fr*)out_ext;
  memcpy(o, a, d->n * 32);
  distribute_powers_zeta(d, o, d->n, 1);
  memset(o + d->n, 0, (d->ext_n - d->n) * 32);
  best_fft(o, d->extended_omega, d->extended_k, d->threads);
}
/* a_ext is clobbered; out gets n * quotient_poly_degree coefficients */
void orc_extended_to_coeff(const orc_domain* d, zk_fr* a_ext, zk_fr* out) {
  fr* a = (fr*)a_ext;
  domain_ifft(d, a, d->extended_omega_inv, d->extended_k, d->extended_ifft_divisor);
  distribute_powers_zeta(d, a, d->ext_n, 0);
  memcpy(out, a, d->n * d->quotient_poly_degree * 32);
}
void orc_divide_by_vanishing_poly(const orc_domain* d, zk_fr* a_ext) {
  fr* a = (fr*)a_ext;
  for (size_t i = 0; i < d->ext_n; i++) a[i] = fr_mul(a[i], d->t_evaluations[i % d->n_t]);
}

/* ------------------------------------------------------------------------ */
/* Evaluator::evaluate_h  [UPSTREAM halo2_proofs/src/plonk/evaluation.rs:      */
/* ValueSource::get, Calculation::evaluate, GraphEvaluator::evaluate,         */
/* Evaluator::evaluate_h; run inside create_proof, sgx_dcap_verifier.rs:814]  */

typedef struct {
  const fr* const* fixed; const fr* const* advice; const fr* const* instance;
  const fr* challenges; fr beta, gamma, theta, y;
} eval_env;

static inline size_t get_rotation_idx(size_t idx, int32_t rot, int32_t rot_scale, int64_t isize) {
  int64_t v = ((int64_t)idx + (int64_t)rot * rot_scale) % isize;
  if (v < 0) v += isize;
  return (size_t)v;
}
static inline fr vs_get(const zk_value_source* s, const zk_graph* g, const size_t* rot, const fr* inter,
                        const eval_env* e, fr prev) {
  switch (s->kind) {
    case ZK_VS_CONSTANT: return ((const fr*)g->constants)[s->a];
    case ZK_VS_INTERMEDIATE: return inter[s->a];
    case ZK_VS_FIXED: return e->fixed[s->a][rot[s->b]];
    case ZK_VS_ADVICE: return e->advice[s->a][rot[s->b]];
    case ZK_VS_INSTANCE: return e->instance[s->a][rot[s->b]];
    case ZK_VS_CHALLENGE: return e->challenges[s->a];
    case ZK_VS_BETA: return e->beta;
    case ZK_VS_GAMMA: return e->gamma;
    case ZK_VS_THETA: return e->theta;
    case ZK_VS_Y: return e->y;
    default: return prev;
  }
}
static fr graph_evaluate(const zk_graph* g, size_t* rot, fr* inter, const eval_env* e, fr prev,
                         size_t idx, int32_t rot_scale, int64_t isize) {
  for (uint32_t i = 0; i < g->n_rotations; i++) rot[i] = get_rotation_idx(idx, g->rotations[i], rot_scale, isize);
  for (uint32_t i = 0; i < g->n_calcs; i++) {
    const zk_calculation* c = &g->calcs[i];
    fr a = vs_get(&c->a, g, rot, inter, e, prev), r;
    switch (c->kind) {
      case ZK_CALC_ADD: r = fr_add(a, vs_get(&c->b, g, rot, inter, e, prev)); break;
      case ZK_CALC_SUB: r = fr_sub(a, vs_get(&c->b, g, rot, inter, e, prev)); break;
      case ZK_CALC_MUL: r = fr_mul(a, vs_get(&c->b, g, rot, inter, e, prev)); break;
      case ZK_CALC_SQUARE: r = fr_sqr(a); break;
      case ZK_CALC_DOUBLE: r = fr_dbl(a); break;
      case ZK_CALC_NEGATE: r = fr_neg(a); break;
      case ZK_CALC_HORNER: {
        fr factor = vs_get(&c->b, g, rot, inter, e, prev);
        r = a;
        for (uint32_t p = 0; p < c->parts_len; p++)
          r = fr_add(fr_mul(r, factor), vs_get(&g->horner_parts[c->parts_off + p], g, rot, inter, e, prev));
        break;
      }
      default: r = a; break;   /* Store */
    }
    inter[c->target] = r;
  }
  if (g->n_calcs) return inter[g->calcs[g->n_calcs - 1].target];
  return fr_ZERO_;
}

typedef struct {
  const zk_evaluator_desc* desc; const orc_domain* dom; eval_env env;
  const fr* l0; const fr* l_last; const fr* l_active;
  const fr* const* perm_cosets;          /* pk.permutation.cosets */
  fr** perm_product_cosets; uint32_t n_sets;
  const fr* product_coset; const fr* pin_coset; const fr* ptab_coset; uint32_t lookup_idx;
  fr* values; int phase;
} evalh_job;

static void evalh_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid;
  evalh_job* J = (evalh_job*)arg;
  const zk_evaluator_desc* D = J->desc; const orc_domain* dom = J->dom;
  int32_t rot_scale = 1 << (dom->extended_k - dom->k);
  int64_t isize = (int64_t)dom->ext_n;
  const eval_env* e = &J->env;
  fr one = fr_R1, y = e->y, beta = e->beta, gamma = e->gamma;
  if (J->phase == 0) {                       /* custom gates */
    const zk_graph* g = &D->custom_gates;
    size_t* rot = malloc((g->n_rotations + 1) * sizeof(size_t));
    fr* inter = malloc((g->num_intermediates + 1) * sizeof(fr));
    for (size_t idx = lo; idx < hi; idx++)
      J->values[idx] = graph_evaluate(g, rot, inter, e, J->values[idx], idx, rot_scale, isize);
    free(rot); free(inter);
  } else if (J->phase == 1) {                /* permutation argument */
    uint32_t bf = D->blinding_factors, chunk_len = D->degree - 2;
    int32_t last_rotation = -(int32_t)(bf + 1);
    fr delta_start = fr_mul(beta, FR_ZETA);
    fr beta_term = fr_pow_u64(dom->extended_omega, (u64)lo);
    fr** z = J->perm_product_cosets; uint32_t ns = J->n_sets;
    for (size_t idx = lo; idx < hi; idx++) {
      size_t r_next = get_rotation_idx(idx, 1, rot_scale, isize);
      size_t r_last = get_rotation_idx(idx, last_rotation, rot_scale, isize);
      fr v = J->values[idx];
      v = fr_add(fr_mul(v, y), fr_mul(fr_sub(one, z[0][idx]), J->l0[idx]));
      fr zl = z[ns - 1][idx];
      v = fr_add(fr_mul(v, y), fr_mul(fr_sub(fr_mul(zl, zl), zl), J->l_last[idx]));
      for (uint32_t s = 1; s < ns; s++)
        v = fr_add(fr_mul(v, y), fr_mul(fr_sub(z[s][idx], z[s - 1][r_last]), J->l0[idx]));
      fr current_delta = fr_mul(delta_start, beta_term);
      for (uint32_t s = 0; s < ns; s++) {
        uint32_t c0 = s * chunk_len, c1 = c0 + chunk_len > D->n_perm_columns ? D->n_perm_columns : c0 + chunk_len;
        fr left = z[s][r_next];
        for (uint32_t c = c0; c < c1; c++) {
          const zk_column* col = &D->perm_columns[c];
          const fr* vals = col->column_type == ZK_COL_ADVICE ? e->advice[col->index]
                         : col->column_type == ZK_COL_FIXED ? e->fixed[col->index] : e->instance[col->index];
          left = fr_mul(left, fr_add(fr_add(vals[idx], fr_mul(beta, J->perm_cosets[c][idx])), gamma));
        }
        fr right = z[s][idx];
        for (uint32_t c = c0; c < c1; c++) {
          const zk_column* col = &D->perm_columns[c];
          const fr* vals = col->column_type == ZK_COL_ADVICE ? e->advice[col->index]
                         : col->column_type == ZK_COL_FIXED ? e->fixed[col->index] : e->instance[col->index];
          right = fr_mul(right, fr_add(fr_add(vals[idx], current_delta), gamma));
          current_delta = fr_mul(current_delta, FR_DELTA);
        }
        v = fr_add(fr_mul(v, y), fr_mul(fr_sub(left, right), J->l_active[idx]));
      }
      beta_term = fr_mul(beta_term, dom->extended_omega);
      J->values[idx] = v;
    }
  } else {                                   /* one lookup argument */
    const zk_graph* g = &D->lookups[J->lookup_idx];
    size_t* rot = malloc((g->n_rotations + 1) * sizeof(size_t));
    fr* inter = malloc((g->num_intermediates + 1) * sizeof(fr));
    const fr* Z = J->product_coset; const fr* A = J->pin_coset; const fr* S = J->ptab_coset;
    for (size_t idx = lo; idx < hi; idx++) {
      fr table_value = graph_evaluate(g, rot, inter, e, fr_ZERO_, idx, rot_scale, isize);
      size_t r_next = get_rotation_idx(idx, 1, rot_scale, isize);
      size_t r_prev = get_rotation_idx(idx, -1, rot_scale, isize);
      fr a_minus_s = fr_sub(A[idx], S[idx]);
      fr v = J->values[idx];
      v = fr_add(fr_mul(v, y), fr_mul(fr_sub(one, Z[idx]), J->l0[idx]));
      v = fr_add(fr_mul(v, y), fr_mul(fr_sub(fr_mul(Z[idx], Z[idx]), Z[idx]), J->l_last[idx]));
      fr t1 = fr_mul(fr_mul(Z[r_next], fr_add(A[idx], beta)), fr_add(S[idx], gamma));
      fr t2 = fr_mul(Z[idx], table_value);
      v = fr_add(fr_mul(v, y), fr_mul(fr_sub(t1, t2), J->l_active[idx]));
      v = fr_add(fr_mul(v, y), fr_mul(a_minus_s, J->l0[idx]));
      v = fr_add(fr_mul(v, y), fr_mul(fr_mul(a_minus_s, fr_sub(A[idx], A[r_prev])), J->l_active[idx]));
      J->values[idx] = v;
    }
    free(rot); free(inter);
  }
}

/* fixed / perm / l0 / l_last / l_active: extended cosets (ext_n each) as held
 * by ProvingKey.  advice / instance / lookup / perm_product: coefficient form. */
void orc_evaluate_h(const zk_evaluator_desc* desc, int threads,
                    const zk_fr* const* fixed_cosets, const zk_fr* const* perm_cosets,
                    const zk_fr* l0, const zk_fr* l_last, const zk_fr* l_active_row,
                    const zk_fr* const* advice_polys, const zk_fr* const* instance_polys,
                    const zk_fr* challenges,
                    const zk_fr* y, const zk_fr* beta, const zk_fr* gamma, const zk_fr* theta,
                    const zk_fr* const* lookup_polys, const zk_fr* const* perm_product_polys,
                    zk_fr* out_ext) {
  ensure_init();
  orc_domain* dom = orc_domain_new(desc->degree, desc->k, threads);
  size_t E = dom->ext_n;
  fr** advice = malloc((desc->n_advice + 1) * sizeof(fr*));
  fr** instance = malloc((desc->n_instance + 1) * sizeof(fr*));
  for (uint32_t i = 0; i < desc->n_advice; i++) { advice[i] = malloc(E * 32); orc_coeff_to_extended(dom, advice_polys[i], (zk_fr*)advice[i]); }
  for (uint32_t i = 0; i < desc->n_instance; i++) { instance[i] = malloc(E * 32); orc_coeff_to_extended(dom, instance_polys[i], (zk_fr*)instance[i]); }
  evalh_job J; memset(&J, 0, sizeof J);
  J.desc = desc; J.dom = dom;
  J.env.fixed = (const fr* const*)fixed_cosets; J.env.advice = (const fr* const*)advice; J.env.instance = (const fr* const*)instance;
  J.env.challenges = (const fr*)challenges;
  memcpy(&J.env.y, y, 32); memcpy(&J.env.beta, beta, 32); memcpy(&J.env.gamma, gamma, 32); memcpy(&J.env.theta, theta, 32);
  J.l0 = (const fr*)l0; J.l_last = (const fr*)l_last; J.l_active = (const fr*)l_active_row;
  J.perm_cosets = (const fr* const*)perm_cosets;
  J.values = (fr*)out_ext;
  memset(out_ext, 0, E * 32);
  J.phase = 0; parallel_for(E, threads, evalh_fn, &J);
  if (desc->n_perm_columns) {
    uint32_t chunk_len = desc->degree - 2;
    uint32_t ns = (desc->n_perm_columns + chunk_len - 1) / chunk_len;
    fr** z = malloc(ns * sizeof(fr*));
    for (uint32_t s = 0; s < ns; s++) { z[s] = malloc(E * 32); orc_coeff_to_extended(dom, perm_product_polys[s], (zk_fr*)z[s]); }
    J.perm_product_cosets = z; J.n_sets = ns;
    J.phase = 1; parallel_for(E, threads, evalh_fn, &J);
    for (uint32_t s = 0; s < ns; s++) free(z[s]);
    free(z);
  }
  for (uint32_t l = 0; l < desc->n_lookups; l++) {
    fr* Z = malloc(E * 32); fr* A = malloc(E * 32); fr* S = malloc(E * 32);
    orc_coeff_to_extended(dom, lookup_polys[3 * l + 0], (zk_fr*)Z);
    orc_coeff_to_extended(dom, lookup_polys[3 * l + 1], (zk_fr*)A);
    orc_coeff_to_extended(dom, lookup_polys[3 * l + 2], (zk_fr*)S);
    J.product_coset = Z; J.pin_coset = A; J.ptab_coset = S; J.lookup_idx = l;
    J.phase = 2; parallel_for(E, threads, evalh_fn, &J);
    free(Z); free(A); free(S);
  }
  for (uint32_t i = 0; i < desc->n_advice; i++) free(advice[i]);
  for (uint32_t i = 0; i < desc->n_instance; i++) free(instance[i]);
  free(advice); free(instance);
  orc_domain_free(dom);
}

/* ------------------------------------------------------------------------ */
/* deterministic test-input generators (oracle side only)                    */

static inline u64 splitmix64(u64* s) {
  u64 z = (*s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
/* uniform Fr via from_u512 of 8 SplitMix64 words */
void orc_gen_scalars(u64 seed, size_t n, zk_fr* out) {
  ensure_init();
  u64 s = seed;
  for (size_t i = 0; i < n; i++) {
    u64 w[8]; for (int j = 0; j < 8; j++) w[j] = splitmix64(&s);
    fr z = fr_from_u512(w); memcpy(&out[i], &z, 32);
  }
}
typedef struct { u64 seed; zk_g1_affine* out; } genp_job;
static void genp_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; genp_job* j = (genp_job*)arg;
  /* sqrt exponent (p+1)/4, p = 3 mod 4 */
  u64 e[4]; memcpy(e, fq_MOD, 32);
  e[0] += 1;    /* no carry: low limb ends in ...47 */
  for (int i = 0; i < 4; i++) e[i] = (e[i] >> 2) | (i < 3 ? (e[i + 1] << 62) : 0);
  for (size_t i = lo; i < hi; i++) {
    u64 s = j->seed ^ (0xD1B54A32D192ED03ULL * (u64)(i + 1));
    u64 w[8]; for (int t = 0; t < 8; t++) w[t] = splitmix64(&s);
    fq x = fq_from_u512(w);
    for (;;) {
      fq rhs = fq_add(fq_mul(fq_sqr(x), x), FQ_B3);
      fq yv = fq_pow(rhs, e);
      if (fq_eq(fq_sqr(yv), rhs)) {
        u64 raw[4]; fq_to_raw(yv, raw);
        if (raw[0] & 1) yv = fq_neg(yv);   /* even canonical y */
        memcpy(&j->out[i].x, &x, 32); memcpy(&j->out[i].y, &yv, 32);
        break;
      }
      x = fq_add(x, fq_R1);
    }
  }
}
/* try-and-increment curve points (cofactor 1: every point is in G1) */
void orc_gen_points(u64 seed, size_t n, int threads, zk_g1_affine* out) {
  ensure_init();
  genp_job j = {seed, out};
  parallel_for(n, threads, genp_fn, &j);
}


/* ------------------------------------------------------------------------ */
/* ParamsKZG::setup with a KNOWN trapdoor (test / benchmark SRS; [UPSTREAM]   */
/* poly/kzg/commitment.rs: g[i] = s^i G, g_lagrange[i] = L_i(s) G with       */
/* L_i(s) = (s^n - 1)/n * w^i / (s - w^i)).  Fixed-base comb:  */
/* table T[j][d] = d * 2^(8j) * G, 32 mixed additions per point.             */
typedef struct { const fr* sc; const g1a* table; zk_g1_affine* out; } fbm_job;
static void fbm_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; fbm_job* j = (fbm_job*)arg;
  enum { CH = 256 };
  g1j acc[CH]; fq zs[CH];
  for (size_t b = lo; b < hi; b += CH) {
    size_t m = hi - b < CH ? hi - b : CH;
    for (size_t i = 0; i < m; i++) {
      u64 e[4]; fr_to_raw(j->sc[b + i], e);
      g1j a = g1j_identity();
      for (int w = 0; w < 32; w++) {
        unsigned d = (unsigned)((e[w / 8] >> (8 * (w % 8))) & 0xff);
        if (d) a = g1j_add_mixed(&a, &j->table[w * 256 + d]);
      }
      acc[i] = a; zs[i] = g1j_is_identity(&a) ? fq_R1 : a.z;
    }
    fq_batch_invert(zs, m);
    for (size_t i = 0; i < m; i++) {
      g1a r;
      if (g1j_is_identity(&acc[i])) { r.x = fq_ZERO_; r.y = fq_ZERO_; }
      else { fq zi2 = fq_sqr(zs[i]); r.x = fq_mul(acc[i].x, zi2); r.y = fq_mul(acc[i].y, fq_mul(zi2, zs[i])); }
      memcpy(&j->out[b + i], &r, 64);
    }
  }
}
static void fixed_base_batch(const g1a* base, const fr* sc, size_t n, int threads, zk_g1_affine* out) {
  g1a* table = (g1a*)malloc(sizeof(g1a) * 32 * 256);
  g1j row = g1a_to_j(base);
  for (int w = 0; w < 32; w++) {
    g1j cur = g1j_identity();
    for (int d = 1; d < 256; d++) { cur = g1j_add(&cur, &row); table[w * 256 + d] = g1j_to_affine(&cur); }
    table[w * 256].x = fq_ZERO_; table[w * 256].y = fq_ZERO_;
    for (int t = 0; t < 8; t++) row = g1j_double(&row);
  }
  fbm_job j = {sc, table, out};
  parallel_for(n, threads, fbm_fn, &j);
  free(table);
}
/* s_raw: the trapdoor as a canonical little-endian integer */
void orc_setup_params(uint32_t k, const u64 s_raw[4], int threads, zk_g1_affine* g, zk_g1_affine* g_lagrange) {
  ensure_init();
  size_t n = (size_t)1 << k;
  fr s = fr_from_raw(s_raw);
  fr w = FR_ROOT_OF_UNITY;
  for (uint32_t i = k; i < 28; i++) w = fr_sqr(w);
  fr* sp = (fr*)malloc(n * 32); fr* li = (fr*)malloc(n * 32); fr* den = (fr*)malloc(n * 32);
  fr cur = fr_R1, wp = fr_R1;
  for (size_t i = 0; i < n; i++) { sp[i] = cur; li[i] = wp; den[i] = fr_sub(s, wp); cur = fr_mul(cur, s); wp = fr_mul(wp, w); }
  /* cur = s^n now */
  fr num = fr_mul(fr_sub(cur, fr_R1), fr_inv(fr_from_u64((u64)n)));
  fr_batch_invert(den, n);
  for (size_t i = 0; i < n; i++) li[i] = fr_mul(fr_mul(num, li[i]), den[i]);
  g1a gen; gen.x = fq_from_u64(1); gen.y = fq_from_u64(2);
  fixed_base_batch(&gen, sp, n, threads, g);
  fixed_base_batch(&gen, li, n, threads, g_lagrange);
  free(sp); free(li); free(den);
}

/* prover tail: lookup / permutation products, kate division (test infrastructure) */
#include "prover_tail.inc.h"
