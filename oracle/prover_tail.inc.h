/*
 * prover_tail.inc.h -- TEST INFRASTRUCTURE (included at the end of zk_oracle.c).
 *
 * CPU restatement of the n-sized loops of halo2's prover that sit between the MSM / FFT /
 * evaluate_h calls of create_proof (section 8f rank 1).  [UPSTREAM] =
 * halo2_proofs 0.2.0 @ zkwebauthn/halo2 c254c75f (Cargo.lock:1314-1316), sources not vendored:
 *   plonk/lookup/prover.rs        compress_expressions, permute_expression_pair, commit_product
 *   plonk/permutation/prover.rs   Argument::commit (grand product of one column set)
 *   poly/kzg/multiopen/shplonk/prover.rs + arithmetic.rs   kate_division, linear combinations
 * Reference call site of all of them: create_proof, circuits/src/sgx_dcap_verifier.rs:814-822.
 */

/* ---- lookup: compress_expressions over the n Lagrange rows ------------------------------- */
/* Evaluates graph `g` (inputs folded with theta, built by the caller) on rows [0,n) of the
 * Lagrange columns with rot_scale = 1, isize = n  [UPSTREAM plonk/evaluation.rs `evaluate`]. */
typedef struct { const zk_graph* g; const eval_env* e; size_t n; fr* out; } rows_job;
static void rows_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; rows_job* J = (rows_job*)arg;
  size_t* rot = malloc((J->g->n_rotations + 1) * sizeof(size_t));
  fr* inter = malloc((J->g->num_intermediates + 1) * sizeof(fr));
  for (size_t i = lo; i < hi; i++) J->out[i] = graph_evaluate(J->g, rot, inter, J->e, fr_ZERO_, i, 1, (int64_t)J->n);
  free(rot); free(inter);
}
void orc_eval_graph_rows(const zk_graph* g, int threads, uint32_t k,
                         const zk_fr* const* fixed, const zk_fr* const* advice, const zk_fr* const* instance,
                         const zk_fr* challenges, const zk_fr* theta, zk_fr* out) {
  ensure_init();
  eval_env e; memset(&e, 0, sizeof e);
  e.fixed = (const fr* const*)fixed; e.advice = (const fr* const*)advice; e.instance = (const fr* const*)instance;
  e.challenges = (const fr*)challenges;
  memcpy(&e.theta, theta, 32);
  rows_job J = {g, &e, (size_t)1 << k, (fr*)out};
  parallel_for(J.n, threads, rows_fn, &J);
}

/* ---- lookup: permute_expression_pair -------------------------------------------------------- */
static int fr_qsort_cmp(const void* a, const void* b) {
  /* canonical (to_raw) little-endian limbs, numeric order */
  const u64* x = (const u64*)a; const u64* y = (const u64*)b;
  for (int i = 3; i >= 0; i--) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
  return 0;
}
/* input/table: first `usable` rows of the compressed expressions (Montgomery).  Writes the
 * first `usable` rows of A' and S' (Montgomery); the caller appends the blinding rows.
 * Returns 0, or -1 when an input value is missing from the table (ConstraintSystemFailure). */
int orc_permute_expression_pair(const zk_fr* input, const zk_fr* table, size_t usable,
                                zk_fr* permuted_input, zk_fr* permuted_table) {
  ensure_init();
  u64 (*a)[4] = malloc(usable * 32); u64 (*t)[4] = malloc(usable * 32);
  for (size_t i = 0; i < usable; i++) { fr x, y; memcpy(&x, &input[i], 32); memcpy(&y, &table[i], 32); fr_to_raw(x, a[i]); fr_to_raw(y, t[i]); }
  qsort(a, usable, 32, fr_qsort_cmp);      /* permuted_input_expression.sort() */
  qsort(t, usable, 32, fr_qsort_cmp);      /* BTreeMap<value, count> == sorted multiset */
  /* leftover_table_map as (distinct value, count) runs of the sorted table */
  size_t nd = 0; size_t* start = malloc((usable + 1) * sizeof(size_t)); u64* count = malloc((usable + 1) * sizeof(u64));
  for (size_t i = 0; i < usable; i++) {
    if (i == 0 || fr_qsort_cmp(t[i], t[i - 1]) != 0) { start[nd] = i; count[nd] = 0; nd++; }
    count[nd - 1]++;
  }
  u64 (*s)[4] = calloc(usable ? usable : 1, 32);
  size_t* repeated = malloc((usable + 1) * sizeof(size_t)); size_t nrep = 0;
  size_t d = 0; int rc = 0;
  for (size_t row = 0; row < usable && rc == 0; row++) {
    if (row == 0 || fr_qsort_cmp(a[row], a[row - 1]) != 0) {
      memcpy(s[row], a[row], 32);
      while (d < nd && fr_qsort_cmp(t[start[d]], a[row]) < 0) d++;      /* both ascending */
      if (d >= nd || fr_qsort_cmp(t[start[d]], a[row]) != 0 || count[d] == 0) rc = -1;
      else count[d]--;
    } else repeated[nrep++] = row;
  }
  if (rc == 0) {
    for (size_t i = 0; i < nd; i++)                                   /* ascending key order */
      for (u64 c = 0; c < count[i]; c++) {
        if (!nrep) { rc = -1; break; }
        memcpy(s[repeated[--nrep]], t[start[i]], 32);                 /* repeated_input_rows.pop() */
      }
    if (nrep) rc = -1;
  }
  for (size_t i = 0; i < usable; i++) {
    fr x = fr_from_raw(a[i]), y = fr_from_raw(s[i]);
    memcpy(&permuted_input[i], &x, 32); memcpy(&permuted_table[i], &y, 32);
  }
  free(a); free(t); free(s); free(start); free(count); free(repeated);
  return rc;
}

/* ---- permutation::Argument::commit: grand product of one column set --------------------------- */
typedef struct {
  size_t n; uint32_t ncols; const fr* const* values; const fr* const* sigma;
  fr beta, gamma, omega, deltaomega0; fr* mv; int phase;
} perm_job;
static void perm_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; perm_job* J = (perm_job*)arg;
  if (J->phase == 0) {
    for (size_t i = lo; i < hi; i++) {
      fr m = fr_R1;
      for (uint32_t c = 0; c < J->ncols; c++)
        m = fr_mul(m, fr_add(fr_add(fr_mul(J->beta, J->sigma[c][i]), J->gamma), J->values[c][i]));
      J->mv[i] = m;
    }
  } else {
    fr dstart = J->deltaomega0;
    for (uint32_t c = 0; c < J->ncols; c++) {
      fr dw = fr_mul(dstart, fr_pow_u64(J->omega, (u64)lo));
      for (size_t i = lo; i < hi; i++) {
        J->mv[i] = fr_mul(J->mv[i], fr_add(fr_add(fr_mul(dw, J->beta), J->gamma), J->values[c][i]));
        dw = fr_mul(dw, J->omega);
      }
      dstart = fr_mul(dstart, FR_DELTA);
    }
  }
}
/* z_out[0] = last_z, z_out[row] = z_out[row-1] * num[row-1] / den[row-1]  (n values, before blinding) */
void orc_permutation_product(uint32_t k, int threads, uint32_t ncols, const zk_fr* const* values,
                             const zk_fr* const* sigma, const zk_fr* beta, const zk_fr* gamma,
                             const zk_fr* omega, const zk_fr* deltaomega0, const zk_fr* last_z, zk_fr* z_out) {
  ensure_init();
  perm_job J; memset(&J, 0, sizeof J);
  J.n = (size_t)1 << k; J.ncols = ncols; J.values = (const fr* const*)values; J.sigma = (const fr* const*)sigma;
  memcpy(&J.beta, beta, 32); memcpy(&J.gamma, gamma, 32); memcpy(&J.omega, omega, 32); memcpy(&J.deltaomega0, deltaomega0, 32);
  J.mv = malloc(J.n * sizeof(fr));
  J.phase = 0; parallel_for(J.n, threads, perm_fn, &J);
  fr_batch_invert(J.mv, J.n);
  J.phase = 1; parallel_for(J.n, threads, perm_fn, &J);
  fr* z = (fr*)z_out; memcpy(&z[0], last_z, 32);
  for (size_t row = 1; row < J.n; row++) z[row] = fr_mul(z[row - 1], J.mv[row - 1]);
  free(J.mv);
}

/* ---- lookup::Permuted::commit_product ----------------------------------------------------------- */
/* z_out[0] = 1, z_out[i+1] = z_out[i] * (a_i + beta)(s_i + gamma) / ((A'_i + beta)(S'_i + gamma)); n values */
void orc_lookup_product(uint32_t k, const zk_fr* cin, const zk_fr* ctab, const zk_fr* pin, const zk_fr* ptab,
                        const zk_fr* beta_, const zk_fr* gamma_, zk_fr* z_out) {
  ensure_init();
  size_t n = (size_t)1 << k;
  fr beta, gamma; memcpy(&beta, beta_, 32); memcpy(&gamma, gamma_, 32);
  const fr* a = (const fr*)cin; const fr* s = (const fr*)ctab; const fr* A = (const fr*)pin; const fr* S = (const fr*)ptab;
  fr* lp = malloc(n * sizeof(fr));
  for (size_t i = 0; i < n; i++) lp[i] = fr_mul(fr_add(beta, A[i]), fr_add(gamma, S[i]));
  fr_batch_invert(lp, n);
  for (size_t i = 0; i < n; i++) { lp[i] = fr_mul(lp[i], fr_add(a[i], beta)); lp[i] = fr_mul(lp[i], fr_add(s[i], gamma)); }
  fr* z = (fr*)z_out; fr state = fr_R1;
  z[0] = state;
  for (size_t i = 0; i + 1 < n; i++) { state = fr_mul(state, lp[i]); z[i + 1] = state; }
  free(lp);
}

/* ---- arithmetic.rs kate_division: a(X) / (X - z), remainder dropped; out has n-1 coefficients ---- */
void orc_kate_division(const zk_fr* a_, size_t n, const zk_fr* z_, zk_fr* out) {
  ensure_init();
  const fr* a = (const fr*)a_; fr* q = (fr*)out; fr z; memcpy(&z, z_, 32);
  fr b = fr_neg(z);                     /* divisor X + b */
  fr tmp = fr_ZERO_;
  for (size_t i = n - 1; i >= 1; i--) {
    fr lead = fr_sub(a[i], tmp);
    q[i - 1] = lead;
    tmp = fr_mul(lead, b);
  }
}

/* out = sum_j scalars[j] * polys[j]  (n coefficients each) */
typedef struct { size_t m; const fr* const* p; const fr* s; fr* out; } lc_job;
static void lc_fn(size_t lo, size_t hi, int tid, void* arg) {
  (void)tid; lc_job* J = (lc_job*)arg;
  for (size_t i = lo; i < hi; i++) {
    fr acc = fr_ZERO_;
    for (size_t j = 0; j < J->m; j++) acc = fr_add(acc, fr_mul(J->p[j][i], J->s[j]));
    J->out[i] = acc;
  }
}
void orc_poly_lincomb(size_t n, int threads, size_t m, const zk_fr* const* polys, const zk_fr* scalars, zk_fr* out) {
  ensure_init();
  lc_job J = {m, (const fr* const*)polys, (const fr*)scalars, (fr*)out};
  parallel_for(n, threads, lc_fn, &J);
}

/* sigma_i[j] = delta^{i'} * omega^{j'} for the cell (i', j') that (i, j) maps to
 * [UPSTREAM plonk/permutation/keygen.rs Assembly::build_pk]; mapping = ncols x n x 2 u32 */
void orc_permutation_sigma(uint32_t k, uint32_t ncols, const uint32_t* mapping, const zk_fr* omega_, zk_fr* out) {
  ensure_init();
  size_t n = (size_t)1 << k; fr omega; memcpy(&omega, omega_, 32);
  fr* wp = malloc(n * sizeof(fr)); fr* dp = malloc((ncols + 1) * sizeof(fr));
  wp[0] = fr_R1; for (size_t i = 1; i < n; i++) wp[i] = fr_mul(wp[i - 1], omega);
  dp[0] = fr_R1; for (uint32_t i = 1; i < ncols; i++) dp[i] = fr_mul(dp[i - 1], FR_DELTA);
  fr* o = (fr*)out;
  for (size_t c = 0; c < ncols; c++)
    for (size_t j = 0; j < n; j++) {
      const uint32_t* m = &mapping[(c * n + j) * 2];
      o[c * n + j] = fr_mul(dp[m[0]], wp[m[1]]);
    }
  free(wp); free(dp);
}
