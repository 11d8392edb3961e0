// evalh.cuh -- per-row evaluator of the quotient numerator h(X) (device + host-testable).
//
// Restates the per-row semantics of halo2_proofs plonk::evaluation::{ValueSource::get,
// Calculation::evaluate, GraphEvaluator::evaluate, Evaluator::evaluate_h} ([UPSTREAM]
// src/plonk/evaluation.rs @ c254c75f, Cargo.lock:1314-1316; run inside create_proof,
// circuits/src/sgx_dcap_verifier.rs:814) as ONE fused pass: for a row of the extended
// domain the custom-gate graph, the permutation argument terms and every lookup argument's
// terms are folded into `value = value * y + term` in exactly the reference order, reading
// every extended column once.  The reference makes 2 + #lookups sweeps over `values`.
//
// The circuit's expression graphs arrive as zk_graph (the C image of GraphEvaluator) and are
// flattened by the host (evalh.cu: compile_program) into a compact word stream that this
// interpreter walks; all threads of a warp execute the same word, so control flow is uniform.
#pragma once
#include "ff.cuh"

namespace zk {

// packed value source: kind << 28 | b << 20 | a
ZK_HD uint32_t vs_pack(uint32_t kind, uint32_t a, uint32_t b) { return (kind << 28) | (b << 20) | a; }

enum : uint32_t {
  VS_CONSTANT = 0, VS_INTERMEDIATE = 1, VS_FIXED = 2, VS_ADVICE = 3, VS_INSTANCE = 4, VS_CHALLENGE = 5,
  VS_BETA = 6, VS_GAMMA = 7, VS_THETA = 8, VS_Y = 9, VS_PREVIOUS = 10
};
enum : uint32_t { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_SQUARE = 3, OP_DOUBLE = 4, OP_NEGATE = 5, OP_HORNER = 6, OP_STORE = 7 };

struct EvalGraph {
  uint32_t code_off;     // first word of the calculation stream
  uint32_t n_calcs;
  uint32_t rot_off;      // n_rot int32 rotations stored in the word stream
  uint32_t n_rot;
  uint32_t const_base;   // first constant of this graph in the constant pool
  uint32_t last_target;
};

constexpr int kMaxLookups = 64;
constexpr int kMaxRot = 16;

struct EvalProgram {
  uint32_t k, ext_k;
  // slots of the column pointer table
  uint32_t off_fixed, off_advice, off_instance, off_sigma, off_z, off_lookup, off_l0, off_l_last, off_l_active;
  uint32_t n_sets, chunk_len, n_perm_cols, n_lookups;
  uint32_t perm_cols_off;          // n_perm_cols words: slot of each permutation column's values
  int32_t last_rotation;           // -(blinding_factors + 1)
  EvalGraph gates;
  EvalGraph lookups[kMaxLookups];
};

struct EvalScalars {
  Fr beta, gamma, theta, y;
  Fr delta_start;    // beta * ZETA
  Fr delta;          // DELTA = 7^(2^28)
  Fr one;
};

template <int NI>
struct EvalState {
  Fr inter[NI];
  uint32_t rot[kMaxRot];
  ZK_HD Fr get(uint32_t i) const { return inter[i]; }
  ZK_HD void set(uint32_t i, const Fr& v) { inter[i] = v; }
};
// After optimisation most programs need <= 4 live intermediates: keep those in REGISTERS (a dynamically
// indexed array would live in local memory) and pick by index with selects.
template <>
struct EvalState<4> {
  Fr r0, r1, r2, r3;
  uint32_t rot[kMaxRot];
  ZK_HD Fr get(uint32_t i) const {
    Fr v;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint32_t lo = (i & 1) ? r1.l[k] : r0.l[k], hi = (i & 1) ? r3.l[k] : r2.l[k];
      v.l[k] = (i & 2) ? hi : lo;
    }
    return v;
  }
  ZK_HD void set(uint32_t i, const Fr& v) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      r0.l[k] = (i == 0) ? v.l[k] : r0.l[k]; r1.l[k] = (i == 1) ? v.l[k] : r1.l[k];
      r2.l[k] = (i == 2) ? v.l[k] : r2.l[k]; r3.l[k] = (i == 3) ? v.l[k] : r3.l[k];
    }
  }
};

// The evaluator multiplies at ~50 call sites; inlining the 165-instruction multiplier at each of them made the
// kernel ~130 KB of code and instruction-fetch stalls its top stall reason (ncu: stalled_no_instruction 5.8 per
// issue).  Field operations are therefore CALLED here, which keeps the whole kernel inside the instruction cache.
#ifdef __CUDA_ARCH__
#define ZK_EV_FN __device__ __noinline__
#else
#define ZK_EV_FN inline
#endif
ZK_EV_FN Fr ev_mul(Fr a, Fr b) { return a * b; }
ZK_EV_FN Fr ev_add(Fr a, Fr b) { return a + b; }
ZK_EV_FN Fr ev_sub(Fr a, Fr b) { return a - b; }
ZK_EV_FN Fr ev_fma(Fr a, Fr b, Fr c) { return a * b + c; }      // a * b + c

ZK_HD uint32_t rot_idx(size_t idx, int32_t rot, uint32_t rot_shift, size_t mask) {
  // (idx + rot * 2^rot_shift) mod 2^ext_k, rot may be negative
  int64_t v = (int64_t)idx + (int64_t)rot * ((int64_t)1 << rot_shift);
  return (uint32_t)((uint64_t)v & mask);
}

template <int NI>
ZK_HD Fr vs_fetch(uint32_t src, const EvalProgram& P, const EvalGraph& G, const uint4* const* cols,
                  const uint4* consts, const uint4* challenges, const EvalScalars& S,
                  const EvalState<NI>& st, const Fr& prev) {
  uint32_t kind = src >> 28, a = src & 0xFFFFFu, b = (src >> 20) & 0xFFu;
  switch (kind) {
    case VS_CONSTANT: return Fr::load(consts + 2 * (size_t)(G.const_base + a));
    case VS_INTERMEDIATE: return st.get(a);
    case VS_FIXED: return Fr::load(cols[P.off_fixed + a] + 2 * (size_t)st.rot[b]);
    case VS_ADVICE: return Fr::load(cols[P.off_advice + a] + 2 * (size_t)st.rot[b]);
    case VS_INSTANCE: return Fr::load(cols[P.off_instance + a] + 2 * (size_t)st.rot[b]);
    case VS_CHALLENGE: return Fr::load(challenges + 2 * (size_t)a);
    case VS_BETA: return S.beta;
    case VS_GAMMA: return S.gamma;
    case VS_THETA: return S.theta;
    case VS_Y: return S.y;
    default: return prev;
  }
}

// GraphEvaluator::evaluate for row idx
template <int NI>
ZK_HD Fr graph_eval(const EvalProgram& P, const EvalGraph& G, const uint32_t* code, const uint4* const* cols,
                    const uint4* consts, const uint4* challenges, const EvalScalars& S, EvalState<NI>& st,
                    const Fr& prev, size_t idx) {
  const uint32_t rot_shift = P.ext_k - P.k;
  const size_t mask = ((size_t)1 << P.ext_k) - 1;
  for (uint32_t r = 0; r < G.n_rot; r++) st.rot[r] = rot_idx(idx, (int32_t)code[G.rot_off + r], rot_shift, mask);
  uint32_t pc = G.code_off;
  for (uint32_t i = 0; i < G.n_calcs; i++) {
    uint32_t w0 = code[pc], target = code[pc + 1], sa = code[pc + 2], sb = code[pc + 3];
    uint32_t op = w0 & 0xFFu, nparts = w0 >> 8;
    pc += 4;
    // one inline copy of the operand fetch for `a`, one for `b` / the Horner parts (code size matters, see above)
    Fr a = vs_fetch<NI>(sa, P, G, cols, consts, challenges, S, st, prev), r = a;
    const bool binary = op <= OP_MUL || op == OP_HORNER;
    const uint32_t nfetch = op == OP_HORNER ? nparts + 1 : (binary ? 1u : 0u);
    Fr f = a;
    for (uint32_t p = 0; p < nfetch; p++) {
      // p == 0: operand b (the Horner factor); p >= 1: Horner part p-1
      Fr b = vs_fetch<NI>(p == 0 ? sb : code[pc + p - 1], P, G, cols, consts, challenges, S, st, prev);
      if (op == OP_HORNER) {
        if (p == 0) f = b; else r = ev_fma(r, f, b);
      } else if (op == OP_ADD) r = ev_add(a, b);
      else if (op == OP_SUB) r = ev_sub(a, b);
      else r = ev_mul(a, b);
    }
    if (op == OP_HORNER) pc += nparts;
    else if (op == OP_SQUARE) r = ev_mul(a, a);
    else if (op == OP_DOUBLE) r = ev_add(a, a);
    else if (op == OP_NEGATE) r = ev_sub(Fr::zero(), a);
    st.set(target, r);
  }
  return G.n_calcs ? st.get(G.last_target) : Fr::zero();
}

// One row of Evaluator::evaluate_h.  `beta_term` = extended_omega^idx.
template <int NI>
ZK_HD Fr evalh_row(const EvalProgram& P, const uint32_t* code, const uint4* const* cols, const uint4* consts,
                   const uint4* challenges, const EvalScalars& S, EvalState<NI>& st, size_t idx, const Fr& beta_term) {
  const uint32_t rot_shift = P.ext_k - P.k;
  const size_t mask = ((size_t)1 << P.ext_k) - 1;
  // custom gates
  Fr value = graph_eval<NI>(P, P.gates, code, cols, consts, challenges, S, st, Fr::zero(), idx);
  const uint4* l0c = cols[P.off_l0];
  const uint4* llc = cols[P.off_l_last];
  const uint4* lac = cols[P.off_l_active];
  Fr l0 = Fr::load(l0c + 2 * idx), l_last = Fr::load(llc + 2 * idx), l_active = Fr::load(lac + 2 * idx);
  // permutation argument
  if (P.n_sets) {
    size_t r_next = rot_idx(idx, 1, rot_shift, mask), r_last = rot_idx(idx, P.last_rotation, rot_shift, mask);
    Fr z0 = Fr::load(cols[P.off_z] + 2 * idx);
    value = ev_fma(value, S.y, ev_mul(ev_sub(S.one, z0), l0));
    Fr zl = Fr::load(cols[P.off_z + P.n_sets - 1] + 2 * idx);
    value = ev_fma(value, S.y, ev_mul(ev_sub(ev_mul(zl, zl), zl), l_last));
    for (uint32_t s = 1; s < P.n_sets; s++) {
      Fr zi = Fr::load(cols[P.off_z + s] + 2 * idx), zp = Fr::load(cols[P.off_z + s - 1] + 2 * r_last);
      value = ev_fma(value, S.y, ev_mul(ev_sub(zi, zp), l0));
    }
    Fr current_delta = ev_mul(S.delta_start, beta_term);
    for (uint32_t s = 0; s < P.n_sets; s++) {
      uint32_t c0 = s * P.chunk_len, c1 = c0 + P.chunk_len < P.n_perm_cols ? c0 + P.chunk_len : P.n_perm_cols;
      Fr left = Fr::load(cols[P.off_z + s] + 2 * r_next);
      Fr right = Fr::load(cols[P.off_z + s] + 2 * idx);
      for (uint32_t c = c0; c < c1; c++) {
        Fr v = Fr::load(cols[code[P.perm_cols_off + c]] + 2 * idx);
        Fr sg = Fr::load(cols[P.off_sigma + c] + 2 * idx);
        left = ev_mul(left, ev_add(ev_fma(S.beta, sg, v), S.gamma));
        right = ev_mul(right, ev_add(ev_add(v, current_delta), S.gamma));
        current_delta = ev_mul(current_delta, S.delta);
      }
      value = ev_fma(value, S.y, ev_mul(ev_sub(left, right), l_active));
    }
  }
  // lookup arguments
  for (uint32_t l = 0; l < P.n_lookups; l++) {
    Fr table_value = graph_eval<NI>(P, P.lookups[l], code, cols, consts, challenges, S, st, Fr::zero(), idx);
    size_t r_next = rot_idx(idx, 1, rot_shift, mask), r_prev = rot_idx(idx, -1, rot_shift, mask);
    const uint4* Zc = cols[P.off_lookup + 3 * l];
    const uint4* Ac = cols[P.off_lookup + 3 * l + 1];
    const uint4* Sc = cols[P.off_lookup + 3 * l + 2];
    Fr Z = Fr::load(Zc + 2 * idx), A = Fr::load(Ac + 2 * idx), Sv = Fr::load(Sc + 2 * idx);
    Fr a_minus_s = ev_sub(A, Sv);
    value = ev_fma(value, S.y, ev_mul(ev_sub(S.one, Z), l0));
    value = ev_fma(value, S.y, ev_mul(ev_sub(ev_mul(Z, Z), Z), l_last));
    Fr Zn = Fr::load(Zc + 2 * r_next);
    Fr t1 = ev_mul(ev_mul(Zn, ev_add(A, S.beta)), ev_add(Sv, S.gamma));
    value = ev_fma(value, S.y, ev_mul(ev_sub(t1, ev_mul(Z, table_value)), l_active));
    value = ev_fma(value, S.y, ev_mul(a_minus_s, l0));
    Fr Ap = Fr::load(Ac + 2 * r_prev);
    value = ev_fma(value, S.y, ev_mul(ev_mul(a_minus_s, ev_sub(A, Ap)), l_active));
  }
  return value;
}

}  // namespace zk
