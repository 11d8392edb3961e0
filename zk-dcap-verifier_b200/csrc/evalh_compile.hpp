// evalh_compile.hpp -- host side: flatten an Evaluator description (zk_evaluator_desc, the C
// image of halo2's Evaluator { custom_gates, lookups } + the permutation argument's column
// list) into the word stream that evalh.cuh interprets.  Pure C++, no CUDA.
#pragma once
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/zkdcap_b200.h"
#include "evalh.cuh"

namespace zk {

struct CompiledEval {
  EvalProgram prog;
  std::vector<uint32_t> code;
  std::vector<zk_fr> consts;
  uint32_t n_slots = 0;
  uint32_t max_intermediates = 0;
  uint64_t muls_per_row = 0;
  uint64_t columns_read = 0;
  std::string error;
};

namespace detail {

inline bool pack_source(const zk_value_source& s, const zk_graph& g, const zk_evaluator_desc& d, uint32_t* out,
                        std::set<std::pair<uint32_t, uint32_t>>& cols, std::string& err) {
  uint32_t limit_a = 0;
  switch (s.kind) {
    case ZK_VS_CONSTANT: limit_a = g.n_constants; break;
    case ZK_VS_INTERMEDIATE: limit_a = g.num_intermediates; break;
    case ZK_VS_FIXED: limit_a = d.n_fixed; break;
    case ZK_VS_ADVICE: limit_a = d.n_advice; break;
    case ZK_VS_INSTANCE: limit_a = d.n_instance; break;
    case ZK_VS_CHALLENGE: limit_a = d.n_challenges; break;
    case ZK_VS_BETA: case ZK_VS_GAMMA: case ZK_VS_THETA: case ZK_VS_Y: case ZK_VS_PREVIOUS_VALUE: limit_a = 1; break;
    default: err = "unknown value source kind"; return false;
  }
  bool column = s.kind == ZK_VS_FIXED || s.kind == ZK_VS_ADVICE || s.kind == ZK_VS_INSTANCE;
  uint32_t a = (s.kind >= ZK_VS_BETA) ? 0 : s.a;
  if (a >= limit_a || a >= (1u << 20)) { err = "value source index out of range"; return false; }
  uint32_t b = 0;
  if (column) {
    if (s.b >= g.n_rotations || s.b >= (uint32_t)kMaxRot) { err = "rotation index out of range"; return false; }
    b = s.b;
    cols.insert({s.kind, a});
  }
  *out = vs_pack(s.kind, a, b);
  return true;
}

inline bool compile_graph(const zk_graph& g, const zk_evaluator_desc& d, CompiledEval& C, EvalGraph* out,
                          std::set<std::pair<uint32_t, uint32_t>>& cols) {
  if (g.n_rotations > (uint32_t)kMaxRot) { C.error = "graph uses more than 16 distinct rotations"; return false; }
  out->const_base = (uint32_t)C.consts.size();
  for (uint32_t i = 0; i < g.n_constants; i++) C.consts.push_back(g.constants[i]);
  out->rot_off = (uint32_t)C.code.size();
  out->n_rot = g.n_rotations;
  for (uint32_t i = 0; i < g.n_rotations; i++) C.code.push_back((uint32_t)g.rotations[i]);
  out->code_off = (uint32_t)C.code.size();
  out->n_calcs = g.n_calcs;
  out->last_target = 0;
  if (g.num_intermediates > C.max_intermediates) C.max_intermediates = g.num_intermediates;
  for (uint32_t i = 0; i < g.n_calcs; i++) {
    const zk_calculation& c = g.calcs[i];
    if (c.kind > ZK_CALC_STORE) { C.error = "unknown calculation kind"; return false; }
    if (c.target >= g.num_intermediates) { C.error = "calculation target out of range"; return false; }
    uint32_t sa = 0, sb = 0;
    if (!pack_source(c.a, g, d, &sa, cols, C.error)) return false;
    bool binary = c.kind == ZK_CALC_ADD || c.kind == ZK_CALC_SUB || c.kind == ZK_CALC_MUL || c.kind == ZK_CALC_HORNER;
    if (binary && !pack_source(c.b, g, d, &sb, cols, C.error)) return false;
    uint32_t nparts = c.kind == ZK_CALC_HORNER ? c.parts_len : 0;
    if (nparts >= (1u << 24) || (nparts && (uint64_t)c.parts_off + nparts > g.n_horner_parts)) { C.error = "horner parts out of range"; return false; }
    C.code.push_back(c.kind | (nparts << 8));
    C.code.push_back(c.target);
    C.code.push_back(sa);
    C.code.push_back(sb);
    for (uint32_t p = 0; p < nparts; p++) {
      uint32_t sp = 0;
      if (!pack_source(g.horner_parts[c.parts_off + p], g, d, &sp, cols, C.error)) return false;
      C.code.push_back(sp);
    }
    if (c.kind == ZK_CALC_MUL || c.kind == ZK_CALC_SQUARE) C.muls_per_row += 1;
    if (c.kind == ZK_CALC_HORNER) C.muls_per_row += nparts;
    out->last_target = c.target;
  }
  return true;
}

}  // namespace detail

inline bool compile_program(const zk_evaluator_desc& d, CompiledEval& C) {
  C = CompiledEval();
  EvalProgram& P = C.prog;
  memset(&P, 0, sizeof P);
  if (d.degree < 3 && d.n_perm_columns) { C.error = "cs.degree() must be >= 3 with a permutation argument"; return false; }
  if (d.n_lookups > (uint32_t)kMaxLookups) { C.error = "too many lookup arguments"; return false; }
  P.k = d.k;
  uint32_t ek = d.k;
  while (((uint64_t)1 << ek) < ((uint64_t)1 << d.k) * (d.degree - 1)) ek++;
  P.ext_k = ek;
  P.chunk_len = d.degree - 2;
  P.n_perm_cols = d.n_perm_columns;
  P.n_sets = d.n_perm_columns ? (d.n_perm_columns + P.chunk_len - 1) / P.chunk_len : 0;
  P.n_lookups = d.n_lookups;
  P.last_rotation = -(int32_t)(d.blinding_factors + 1);
  uint32_t slot = 0;
  P.off_fixed = slot; slot += d.n_fixed;
  P.off_advice = slot; slot += d.n_advice;
  P.off_instance = slot; slot += d.n_instance;
  P.off_sigma = slot; slot += d.n_perm_columns;
  P.off_z = slot; slot += P.n_sets;
  P.off_lookup = slot; slot += 3 * d.n_lookups;
  P.off_l0 = slot++; P.off_l_last = slot++; P.off_l_active = slot++;
  C.n_slots = slot;
  std::set<std::pair<uint32_t, uint32_t>> cols;
  if (!detail::compile_graph(d.custom_gates, d, C, &P.gates, cols)) return false;
  for (uint32_t l = 0; l < d.n_lookups; l++)
    if (!detail::compile_graph(d.lookups[l], d, C, &P.lookups[l], cols)) return false;
  P.perm_cols_off = (uint32_t)C.code.size();
  for (uint32_t c = 0; c < d.n_perm_columns; c++) {
    const zk_column& col = d.perm_columns[c];
    uint32_t base, lim, kind;
    if (col.column_type == ZK_COL_ADVICE) { base = P.off_advice; lim = d.n_advice; kind = ZK_VS_ADVICE; }
    else if (col.column_type == ZK_COL_FIXED) { base = P.off_fixed; lim = d.n_fixed; kind = ZK_VS_FIXED; }
    else if (col.column_type == ZK_COL_INSTANCE) { base = P.off_instance; lim = d.n_instance; kind = ZK_VS_INSTANCE; }
    else { C.error = "unknown permutation column type"; return false; }
    if (col.index >= lim) { C.error = "permutation column index out of range"; return false; }
    C.code.push_back(base + col.index);
    cols.insert({kind, col.index});
  }
  // fixed structure of evaluate_h (see evalh_row)
  if (P.n_sets) {
    C.muls_per_row += 2 + 3 + 2 * (P.n_sets - 1) + 1 + 4 * (uint64_t)P.n_perm_cols + 2 * (uint64_t)P.n_sets;
  }
  C.muls_per_row += 15 * (uint64_t)P.n_lookups;
  C.columns_read = cols.size() + P.n_perm_cols + P.n_sets + 3 * (uint64_t)P.n_lookups + 3;
  if (C.consts.empty()) { zk_fr z; memset(&z, 0, sizeof z); C.consts.push_back(z); }
  if (C.code.empty()) C.code.push_back(0);
  return true;
}

}  // namespace zk
