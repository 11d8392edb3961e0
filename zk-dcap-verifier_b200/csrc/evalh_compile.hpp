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

// One calculation of the intermediate representation the optimiser works on.
struct IrOp {
  uint32_t kind, target, a, b;          // packed sources
  std::vector<uint32_t> parts;          // Horner parts (packed)
};

inline bool is_inter(uint32_t src) { return (src >> 28) == VS_INTERMEDIATE; }
inline uint32_t inter_idx(uint32_t src) { return src & 0xFFFFFu; }

// Value-preserving rewrites of a GraphEvaluator program (the reference's own add_expression rules are
// likewise only value preserving):
//   1. Store forwarding: `Store(src)` defines no new value -- its consumers read `src` directly.
//   2. Horner splitting: Horner(start, [p0..pm), f) becomes m steps acc = acc*f + p_i, each placed right after
//      p_i is available, so that the parts do not all have to stay live until the end.
//   3. Slot allocation: intermediates are renamed onto a minimal set of slots by live range.
// The result keeps per-thread state small enough to stay in L1 (the unoptimised graphs needed 4 KiB per row).
inline void optimise_graph(std::vector<IrOp>& ops, uint32_t& num_intermediates, uint32_t& last_target) {
  if (ops.empty()) return;
  // ---- 1. store forwarding -----------------------------------------------------------------------
  std::vector<uint32_t> alias(num_intermediates, 0xFFFFFFFFu);
  auto resolve = [&](uint32_t src) {
    while (is_inter(src) && alias[inter_idx(src)] != 0xFFFFFFFFu) src = alias[inter_idx(src)];
    return src;
  };
  std::vector<IrOp> fwd;
  for (size_t i = 0; i < ops.size(); i++) {
    IrOp o = ops[i];
    o.a = resolve(o.a);
    o.b = resolve(o.b);
    for (auto& p : o.parts) p = resolve(p);
    bool last = (i + 1 == ops.size());
    if (o.kind == OP_STORE && !last) { alias[o.target] = o.a; continue; }
    fwd.push_back(o);
  }
  // ---- 2. Horner splitting ------------------------------------------------------------------------
  // def_pos[t] = index in `fwd` of the op defining intermediate t
  std::vector<IrOp> sched;
  {
    std::vector<int> def_pos(num_intermediates, -1);
    for (size_t i = 0; i < fwd.size(); i++) def_pos[fwd[i].target] = (int)i;
    // steps[i] = list of Horner steps to emit right after fwd[i]
    std::vector<std::vector<IrOp>> after(fwd.size());
    std::vector<bool> drop(fwd.size(), false);
    for (size_t h = 0; h < fwd.size(); h++) {
      const IrOp& H = fwd[h];
      if (H.kind != OP_HORNER || H.parts.size() < 2) continue;
      // the start value must be available before the first step: place steps no earlier than its definition
      int pos = is_inter(H.a) ? def_pos[inter_idx(H.a)] : -1;
      if (is_inter(H.b) && def_pos[inter_idx(H.b)] > pos) pos = def_pos[inter_idx(H.b)];
      uint32_t acc = H.a;
      bool ok = true;
      std::vector<std::pair<int, IrOp>> steps;
      for (size_t i = 0; i < H.parts.size(); i++) {
        uint32_t pi = H.parts[i];
        int dp = is_inter(pi) ? def_pos[inter_idx(pi)] : -1;
        if (dp >= (int)h) { ok = false; break; }
        if (dp > pos) pos = dp;
        IrOp st;
        st.kind = OP_HORNER; st.a = acc; st.b = H.b; st.parts = {pi};
        bool final_step = (i + 1 == H.parts.size());
        st.target = final_step ? H.target : num_intermediates++;
        acc = vs_pack(VS_INTERMEDIATE, st.target, 0);
        steps.push_back({pos, st});
      }
      if (!ok) continue;
      drop[h] = true;
      for (auto& ps : steps) {
        // position -1 = before every other op: kept in after[0] with a marker bit
        (ps.first < 0 ? after[0] : after[ps.first]).push_back(ps.second);
        if (ps.first < 0) after[0].back().kind |= 0x100u;
      }
    }
    for (size_t i = 0; i < fwd.size(); i++) {
      // steps flagged 0x100 go BEFORE fwd[i] (only possible for i == 0)
      for (auto& st : after[i]) if (st.kind & 0x100u) { IrOp c = st; c.kind &= 0xFFu; sched.push_back(c); }
      if (!drop[i]) sched.push_back(fwd[i]);
      for (auto& st : after[i]) if (!(st.kind & 0x100u)) sched.push_back(st);
    }
  }
  // ---- 3. slot allocation by live range ----------------------------------------------------------------
  std::vector<int> last_use(num_intermediates, -1);
  auto use = [&](uint32_t src, int i) { if (is_inter(src)) last_use[inter_idx(src)] = i; };
  for (size_t i = 0; i < sched.size(); i++) {
    use(sched[i].a, (int)i); use(sched[i].b, (int)i);
    for (auto p : sched[i].parts) use(p, (int)i);
  }
  uint32_t final_t = sched.back().target;
  last_use[final_t] = (int)sched.size();               // the graph's result must survive
  std::vector<uint32_t> slot_of(num_intermediates, 0xFFFFFFFFu), free_slots;
  uint32_t n_slots = 0;
  auto remap = [&](uint32_t src) { return is_inter(src) ? vs_pack(VS_INTERMEDIATE, slot_of[inter_idx(src)], 0) : src; };
  for (size_t i = 0; i < sched.size(); i++) {
    IrOp& o = sched[i];
    uint32_t oa = o.a, ob = o.b;
    std::vector<uint32_t> oparts = o.parts;
    o.a = remap(o.a); o.b = remap(o.b);
    for (auto& p : o.parts) p = remap(p);
    // operands whose last use is this op release their slot before the target is allocated (in-place update)
    auto release = [&](uint32_t src) {
      if (is_inter(src) && last_use[inter_idx(src)] == (int)i && slot_of[inter_idx(src)] != 0xFFFFFFFFu) {
        free_slots.push_back(slot_of[inter_idx(src)]);
        last_use[inter_idx(src)] = -2;                  // released once
      }
    };
    release(oa); release(ob);
    for (auto p : oparts) release(p);
    uint32_t sl;
    if (!free_slots.empty()) { sl = free_slots.back(); free_slots.pop_back(); } else sl = n_slots++;
    slot_of[o.target] = sl;
    if (last_use[o.target] == -1) { /* dead value: its slot is reusable right away */ free_slots.push_back(sl); }
    o.target = sl;
  }
  ops.swap(sched);
  num_intermediates = n_slots ? n_slots : 1;
  last_target = ops.back().target;
}

inline bool compile_graph(const zk_graph& g, const zk_evaluator_desc& d, CompiledEval& C, EvalGraph* out,
                          std::set<std::pair<uint32_t, uint32_t>>& cols, bool optimise) {
  if (g.n_rotations > (uint32_t)kMaxRot) { C.error = "graph uses more than 16 distinct rotations"; return false; }
  out->const_base = (uint32_t)C.consts.size();
  for (uint32_t i = 0; i < g.n_constants; i++) C.consts.push_back(g.constants[i]);
  out->rot_off = (uint32_t)C.code.size();
  out->n_rot = g.n_rotations;
  for (uint32_t i = 0; i < g.n_rotations; i++) C.code.push_back((uint32_t)g.rotations[i]);
  std::vector<IrOp> ops;
  for (uint32_t i = 0; i < g.n_calcs; i++) {
    const zk_calculation& c = g.calcs[i];
    if (c.kind > ZK_CALC_STORE) { C.error = "unknown calculation kind"; return false; }
    if (c.target >= g.num_intermediates) { C.error = "calculation target out of range"; return false; }
    IrOp o; o.kind = c.kind; o.target = c.target; o.a = 0; o.b = 0;
    if (!pack_source(c.a, g, d, &o.a, cols, C.error)) return false;
    bool binary = c.kind == ZK_CALC_ADD || c.kind == ZK_CALC_SUB || c.kind == ZK_CALC_MUL || c.kind == ZK_CALC_HORNER;
    if (binary && !pack_source(c.b, g, d, &o.b, cols, C.error)) return false;
    uint32_t nparts = c.kind == ZK_CALC_HORNER ? c.parts_len : 0;
    if (nparts >= (1u << 24) || (nparts && (uint64_t)c.parts_off + nparts > g.n_horner_parts)) { C.error = "horner parts out of range"; return false; }
    for (uint32_t p = 0; p < nparts; p++) {
      uint32_t sp = 0;
      if (!pack_source(g.horner_parts[c.parts_off + p], g, d, &sp, cols, C.error)) return false;
      o.parts.push_back(sp);
    }
    ops.push_back(o);
  }
  uint32_t n_inter = g.num_intermediates, last_target = ops.empty() ? 0 : ops.back().target;
  if (optimise) optimise_graph(ops, n_inter, last_target);
  out->code_off = (uint32_t)C.code.size();
  out->n_calcs = (uint32_t)ops.size();
  out->last_target = last_target;
  if (n_inter > C.max_intermediates) C.max_intermediates = n_inter;
  for (const IrOp& o : ops) {
    C.code.push_back(o.kind | ((uint32_t)o.parts.size() << 8));
    C.code.push_back(o.target);
    C.code.push_back(o.a);
    C.code.push_back(o.b);
    for (auto p : o.parts) C.code.push_back(p);
    if (o.kind == OP_MUL || o.kind == OP_SQUARE) C.muls_per_row += 1;
    if (o.kind == OP_HORNER) C.muls_per_row += o.parts.size();
  }
  return true;
}

}  // namespace detail

inline bool compile_program(const zk_evaluator_desc& d, CompiledEval& C, bool optimise = true) {
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
  if (!detail::compile_graph(d.custom_gates, d, C, &P.gates, cols, optimise)) return false;
  for (uint32_t l = 0; l < d.n_lookups; l++)
    if (!detail::compile_graph(d.lookups[l], d, C, &P.lookups[l], cols, optimise)) return false;
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
