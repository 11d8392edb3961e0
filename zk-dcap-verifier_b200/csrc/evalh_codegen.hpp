// evalh_codegen.hpp -- host side: emit straight-line CUDA C++ for one circuit's evaluate_h.
//
// north_star / SURVEY K8: "a fused per-row gate evaluator COMPILED from the circuit's expression graph".  The optimised
// program that evalh_compile.hpp produces for the interpreter (store forwarding, Horner splitting, slot allocation) is
// decoded here into source text: every calculation becomes one statement on named locals (the slots live in registers,
// there is no operand decoding, no program counter and no local-memory slot file), every column access has its column
// offset and rotation as literals, and the fixed structure of Evaluator::evaluate_h (permutation argument, the five
// terms of each lookup argument; [UPSTREAM] plonk/evaluation.rs @ c254c75f) is unrolled with the circuit's own counts.
// evalh.cu compiles the text with NVRTC for sm_100a at key generation (zk_pk_upload / zk_keygen) and launches the
// resulting kernel; the interpreter (evalh.cuh) stays as the checked reference: tests compare the two bit for bit, on
// the device and -- through the host flavour of the same text -- on the CPU (tests/test_evalh_host.py).
//
// The statement order is exactly the interpreter's, so both produce identical field values.  Pure C++, no CUDA.
#pragma once
#include <stdarg.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "evalh_compile.hpp"

namespace zk {

namespace codegen {

struct Emitter {
  std::string s;
  void line(const std::string& t) { s += t; s += '\n'; }
  void fmt(const char* f, ...) __attribute__((format(printf, 2, 3))) {
    char b[1024];
    va_list ap; va_start(ap, f); vsnprintf(b, sizeof b, f, ap); va_end(ap);
    s += b;
  }
};

// where a column slot lives: base 0 = resident (key) columns, 1 = witness columns; index = column number inside it
struct SlotMap {
  std::vector<int> base, index;
};
inline SlotMap slot_map(const CompiledEval& C, const zk_evaluator_desc& D) {
  const EvalProgram& P = C.prog;
  SlotMap M; M.base.assign(C.n_slots, 0); M.index.assign(C.n_slots, 0);
  size_t ri = 0, wi = 0;
  for (uint32_t i = 0; i < D.n_fixed; i++) { M.base[P.off_fixed + i] = 0; M.index[P.off_fixed + i] = (int)ri++; }
  for (uint32_t i = 0; i < D.n_perm_columns; i++) { M.base[P.off_sigma + i] = 0; M.index[P.off_sigma + i] = (int)ri++; }
  M.base[P.off_l0] = 0; M.index[P.off_l0] = (int)ri++;
  M.base[P.off_l_last] = 0; M.index[P.off_l_last] = (int)ri++;
  M.base[P.off_l_active] = 0; M.index[P.off_l_active] = (int)ri++;
  for (uint32_t i = 0; i < D.n_advice; i++) { M.base[P.off_advice + i] = 1; M.index[P.off_advice + i] = (int)wi++; }
  for (uint32_t i = 0; i < D.n_instance; i++) { M.base[P.off_instance + i] = 1; M.index[P.off_instance + i] = (int)wi++; }
  for (uint32_t i = 0; i < P.n_sets; i++) { M.base[P.off_z + i] = 1; M.index[P.off_z + i] = (int)wi++; }
  for (uint32_t i = 0; i < 3 * P.n_lookups; i++) { M.base[P.off_lookup + i] = 1; M.index[P.off_lookup + i] = (int)wi++; }
  return M;
}

inline std::string col_expr(const SlotMap& M, uint32_t slot) {
  char b[96];
  snprintf(b, sizeof b, "COLP(%u, %d, %d)", slot, M.base[slot], M.index[slot]);
  return b;
}

inline std::string operand(uint32_t src, const EvalProgram& P, const EvalGraph& G, const SlotMap& M) {
  const uint32_t kind = src >> 28, a = src & 0xFFFFFu, b = (src >> 20) & 0xFFu;
  char t[160];
  switch (kind) {
    case VS_CONSTANT: snprintf(t, sizeof t, "KLOAD(%u)", G.const_base + a); return t;
    case VS_INTERMEDIATE: snprintf(t, sizeof t, "s%u", a); return t;
    case VS_FIXED: snprintf(t, sizeof t, "Fr::load(%s + 2 * (size_t)r%u)", col_expr(M, P.off_fixed + a).c_str(), b); return t;
    case VS_ADVICE: snprintf(t, sizeof t, "Fr::load(%s + 2 * (size_t)r%u)", col_expr(M, P.off_advice + a).c_str(), b); return t;
    case VS_INSTANCE: snprintf(t, sizeof t, "Fr::load(%s + 2 * (size_t)r%u)", col_expr(M, P.off_instance + a).c_str(), b); return t;
    case VS_CHALLENGE: snprintf(t, sizeof t, "Fr::load(challenges + 2 * %u)", a); return t;
    case VS_BETA: return "SC_BETA";
    case VS_GAMMA: return "SC_GAMMA";
    case VS_THETA: return "SC_THETA";
    case VS_Y: return "SC_Y";
    default: return "Fr::zero()";          // PreviousValue: evaluate_h starts every graph from zero
  }
}

// one GraphEvaluator program -> a block that leaves its value in `result`
inline void emit_graph(Emitter& E, const CompiledEval& C, const EvalGraph& G, const SlotMap& M, const char* result) {
  const EvalProgram& P = C.prog;
  const uint32_t* code = C.code.data();
  E.line("  {");
  for (uint32_t r = 0; r < G.n_rot; r++)
    E.fmt("    const uint32_t r%u = ROT(%d);\n", r, (int32_t)code[G.rot_off + r]);
  // slots used by this graph
  uint32_t max_slot = 0;
  {
    uint32_t pc = G.code_off;
    for (uint32_t i = 0; i < G.n_calcs; i++) {
      uint32_t w0 = code[pc], target = code[pc + 1];
      if (target > max_slot) max_slot = target;
      pc += 4 + ((w0 & 0xFFu) == OP_HORNER ? (w0 >> 8) : 0);
    }
  }
  if (G.n_calcs) {
    E.fmt("    Fr s0");
    for (uint32_t i = 1; i <= max_slot; i++) E.fmt(", s%u", i);
    E.line(";");
  }
  uint32_t pc = G.code_off;
  for (uint32_t i = 0; i < G.n_calcs; i++) {
    const uint32_t w0 = code[pc], target = code[pc + 1], sa = code[pc + 2], sb = code[pc + 3];
    const uint32_t op = w0 & 0xFFu, nparts = w0 >> 8;
    pc += 4;
    const std::string a = operand(sa, P, G, M);
    switch (op) {
      case OP_ADD: E.fmt("    s%u = ev_add(%s, %s);\n", target, a.c_str(), operand(sb, P, G, M).c_str()); break;
      case OP_SUB: E.fmt("    s%u = ev_sub(%s, %s);\n", target, a.c_str(), operand(sb, P, G, M).c_str()); break;
      case OP_MUL: E.fmt("    s%u = ev_mul(%s, %s);\n", target, a.c_str(), operand(sb, P, G, M).c_str()); break;
      case OP_SQUARE: E.fmt("    { const Fr t_ = %s; s%u = ev_mul(t_, t_); }\n", a.c_str(), target); break;
      case OP_DOUBLE: E.fmt("    { const Fr t_ = %s; s%u = ev_add(t_, t_); }\n", a.c_str(), target); break;
      case OP_NEGATE: E.fmt("    s%u = ev_sub(Fr::zero(), %s);\n", target, a.c_str()); break;
      case OP_HORNER: {
        E.fmt("    { Fr h_ = %s; const Fr f_ = %s;\n", a.c_str(), operand(sb, P, G, M).c_str());
        for (uint32_t p = 0; p < nparts; p++) E.fmt("      h_ = ev_fma(h_, f_, %s);\n", operand(code[pc + p], P, G, M).c_str());
        E.fmt("      s%u = h_; }\n", target);
        pc += nparts;
        break;
      }
      default: E.fmt("    s%u = %s;\n", target, a.c_str()); break;      // OP_STORE
    }
  }
  if (G.n_calcs) E.fmt("    %s = s%u;\n", result, G.last_target);
  else E.fmt("    %s = Fr::zero();\n", result);
  E.line("  }");
}

}  // namespace codegen

// The whole translation unit.  `host_flavour`: a plain C++ function `zk_evalh_gen_host(...)` over a column pointer table
// that includes the real ff.cuh (CPU test); otherwise a self-contained CUDA kernel `zk_evalh_gen` (`prelude` = the text of
// ff.cuh, embedded by the build).
inline std::string generate_evalh_source(const CompiledEval& C, const zk_evaluator_desc& D, bool host_flavour, const std::string& prelude) {
  using namespace codegen;
  const EvalProgram& P = C.prog;
  const SlotMap M = slot_map(C, D);
  Emitter E;
  E.line("// generated by csrc/evalh_codegen.hpp -- evaluate_h of one circuit as straight-line code");
  if (host_flavour) {
    E.line("#include <stddef.h>");
    E.line("#include <string.h>");
    E.line("struct uint4 { unsigned x, y, z, w; };");
    E.line("#include \"ff.cuh\"");
    E.line("#define ZK_GEN_FN static inline");
    E.line("#define ZK_GEN_CALL static");
    E.line("#define COLP(slot, base, index) ((const uint4*)cols[slot])");
  } else {
    E.line("typedef unsigned int uint32_t; typedef int int32_t; typedef unsigned long long uint64_t; typedef long long int64_t;");
    E.line("#define ZK_NVRTC 1");
    E.s += prelude;
    E.line("");
    E.line("#define ZK_GEN_FN __device__ __forceinline__");
    E.line("#define ZK_GEN_CALL __device__ __noinline__");
    E.fmt("#define COLP(slot, base, index) (((base) ? wit : res) + (size_t)(index) * %lluull)\n", (unsigned long long)(2ull << P.ext_k));
  }
  E.line("using namespace zk;");
  // constants of all graphs (Montgomery form, as stored by the interpreter's constant pool)
  E.fmt("%s uint32_t ZK_KC[%zu][8] = {\n", host_flavour ? "static const" : "__constant__", C.consts.size());
  for (size_t i = 0; i < C.consts.size(); i++) {
    uint32_t w[8];
    memcpy(w, &C.consts[i], 32);
    E.fmt("  {0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu},\n", w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
  }
  E.line("};");
  E.line("ZK_GEN_FN Fr KLOAD(uint32_t i) { Fr r; for (int j = 0; j < 8; j++) r.l[j] = ZK_KC[i][j]; return r; }");
  // field operations: called on the device (one copy of the multiplier in the instruction cache), as in the interpreter
  E.line("ZK_GEN_CALL Fr ev_mul(Fr a, Fr b) { return a * b; }");
  E.line("ZK_GEN_CALL Fr ev_fma(Fr a, Fr b, Fr c) { return a * b + c; }");
  E.line("ZK_GEN_CALL Fr ev_add(Fr a, Fr b) { return a + b; }");
  E.line("ZK_GEN_CALL Fr ev_sub(Fr a, Fr b) { return a - b; }");
  // scalars: beta, gamma, theta, y, delta_start, delta, one -- seven Fr in a small buffer
  E.line("#define SC_BETA Fr::load(sc + 0)");
  E.line("#define SC_GAMMA Fr::load(sc + 2)");
  E.line("#define SC_THETA Fr::load(sc + 4)");
  E.line("#define SC_Y Fr::load(sc + 6)");
  E.line("#define SC_DELTA_START Fr::load(sc + 8)");
  E.line("#define SC_DELTA Fr::load(sc + 10)");
  E.line("#define SC_ONE Fr::load(sc + 12)");
  E.fmt("#define ROT(r) ((uint32_t)(((long long)idx + (long long)(r) * %lldll) & %lldll))\n", 1ll << (P.ext_k - P.k), (1ll << P.ext_k) - 1);
  // the five terms of one lookup argument (the same code for every lookup: called, not inlined)
  if (P.n_lookups) {
    E.line("ZK_GEN_CALL Fr zk_lookup_terms(Fr value, Fr table_value, const uint4* Zc, const uint4* Ac, const uint4* Sc, const uint4* l0c,");
    E.line("                               const uint4* llc, const uint4* lac, const uint4* sc, size_t idx) {");
    E.line("  const uint32_t r_next = ROT(1), r_prev = ROT(-1);");
    E.line("  const Fr y = SC_Y, one = SC_ONE;");
    E.line("  const Fr l0 = Fr::load(l0c + 2 * idx), l_last = Fr::load(llc + 2 * idx), l_active = Fr::load(lac + 2 * idx);");
    E.line("  const Fr Z = Fr::load(Zc + 2 * idx), A = Fr::load(Ac + 2 * idx), Sv = Fr::load(Sc + 2 * idx);");
    E.line("  const Fr a_minus_s = ev_sub(A, Sv);");
    E.line("  value = ev_fma(value, y, ev_mul(ev_sub(one, Z), l0));");
    E.line("  value = ev_fma(value, y, ev_mul(ev_sub(ev_mul(Z, Z), Z), l_last));");
    E.line("  const Fr Zn = Fr::load(Zc + 2 * (size_t)r_next);");
    E.line("  const Fr t1 = ev_mul(ev_mul(Zn, ev_add(A, SC_BETA)), ev_add(Sv, SC_GAMMA));");
    E.line("  value = ev_fma(value, y, ev_mul(ev_sub(t1, ev_mul(Z, table_value)), l_active));");
    E.line("  value = ev_fma(value, y, ev_mul(a_minus_s, l0));");
    E.line("  const Fr Ap = Fr::load(Ac + 2 * (size_t)r_prev);");
    E.line("  value = ev_fma(value, y, ev_mul(ev_mul(a_minus_s, ev_sub(A, Ap)), l_active));");
    E.line("  return value;");
    E.line("}");
  }
  // one (column, sigma) step of a permutation set
  if (P.n_sets) {
    E.line("struct ZkPermStep { Fr left, right, delta; };");
    E.line("ZK_GEN_CALL ZkPermStep zk_perm_step(ZkPermStep st, const uint4* vc, const uint4* sgc, const uint4* sc, size_t idx) {");
    E.line("  const Fr v = Fr::load(vc + 2 * idx), sg = Fr::load(sgc + 2 * idx), gamma = SC_GAMMA;");
    E.line("  st.left = ev_mul(st.left, ev_add(ev_fma(SC_BETA, sg, v), gamma));");
    E.line("  st.right = ev_mul(st.right, ev_add(ev_add(v, st.delta), gamma));");
    E.line("  st.delta = ev_mul(st.delta, SC_DELTA);");
    E.line("  return st;");
    E.line("}");
  }
  // the row function
  if (host_flavour) {
    E.line("static Fr zk_evalh_row(const void* const* cols, const uint4* challenges, const uint4* sc, size_t idx, const Fr& beta_term) {");
  } else {
    E.line("__device__ __forceinline__ Fr zk_evalh_row(const uint4* res, const uint4* wit, const uint4* challenges, const uint4* sc, size_t idx, const Fr& beta_term) {");
  }
  E.line("  Fr value;");
  emit_graph(E, C, P.gates, M, "value");
  E.fmt("  const uint4* l0c = %s; const uint4* llc = %s; const uint4* lac = %s;\n", col_expr(M, P.off_l0).c_str(), col_expr(M, P.off_l_last).c_str(),
        col_expr(M, P.off_l_active).c_str());
  if (P.n_sets) {
    E.line("  {");
    E.line("    const Fr y = SC_Y, one = SC_ONE;");
    E.line("    const Fr l0 = Fr::load(l0c + 2 * idx), l_last = Fr::load(llc + 2 * idx), l_active = Fr::load(lac + 2 * idx);");
    E.fmt("    const uint32_t r_next = ROT(1), r_last = ROT(%d);\n", P.last_rotation);
    E.fmt("    { const Fr z0 = Fr::load(%s + 2 * idx); value = ev_fma(value, y, ev_mul(ev_sub(one, z0), l0)); }\n", col_expr(M, P.off_z).c_str());
    E.fmt("    { const Fr zl = Fr::load(%s + 2 * idx); value = ev_fma(value, y, ev_mul(ev_sub(ev_mul(zl, zl), zl), l_last)); }\n",
          col_expr(M, P.off_z + P.n_sets - 1).c_str());
    for (uint32_t s = 1; s < P.n_sets; s++)
      E.fmt("    { const Fr zi = Fr::load(%s + 2 * idx), zp = Fr::load(%s + 2 * (size_t)r_last); value = ev_fma(value, y, ev_mul(ev_sub(zi, zp), l0)); }\n",
            col_expr(M, P.off_z + s).c_str(), col_expr(M, P.off_z + s - 1).c_str());
    E.line("    ZkPermStep st; st.delta = ev_mul(SC_DELTA_START, beta_term);");
    for (uint32_t s = 0; s < P.n_sets; s++) {
      const uint32_t c0 = s * P.chunk_len, c1 = c0 + P.chunk_len < P.n_perm_cols ? c0 + P.chunk_len : P.n_perm_cols;
      E.fmt("    st.left = Fr::load(%s + 2 * (size_t)r_next); st.right = Fr::load(%s + 2 * idx);\n", col_expr(M, P.off_z + s).c_str(),
            col_expr(M, P.off_z + s).c_str());
      for (uint32_t c = c0; c < c1; c++)
        E.fmt("    st = zk_perm_step(st, %s, %s, sc, idx);\n", col_expr(M, C.code[P.perm_cols_off + c]).c_str(), col_expr(M, P.off_sigma + c).c_str());
      E.line("    value = ev_fma(value, y, ev_mul(ev_sub(st.left, st.right), l_active));");
    }
    E.line("  }");
  }
  for (uint32_t l = 0; l < P.n_lookups; l++) {
    E.line("  {");
    E.line("    Fr table_value;");
    emit_graph(E, C, P.lookups[l], M, "table_value");
    E.fmt("    value = zk_lookup_terms(value, table_value, %s, %s, %s, l0c, llc, lac, sc, idx);\n", col_expr(M, P.off_lookup + 3 * l).c_str(),
          col_expr(M, P.off_lookup + 3 * l + 1).c_str(), col_expr(M, P.off_lookup + 3 * l + 2).c_str());
    E.line("  }");
  }
  E.line("  return value;");
  E.line("}");
  if (host_flavour) {
    // rows [row0, row1); beta_term = ext_omega^idx, advanced by one multiplication per row
    E.line("extern \"C\" void zk_evalh_gen_host(const void* const* cols, const void* challenges, const void* sc, const unsigned* ext_omega,");
    E.line("                                  unsigned* out, size_t row0, size_t row1) {");
    E.line("  Fr w; memcpy(w.l, ext_omega, 32);");
    E.line("  Fr bt = Fr::pow_u64(w, (uint64_t)row0);");
    E.line("  for (size_t idx = row0; idx < row1; idx++) {");
    E.line("    Fr v = zk_evalh_row(cols, (const uint4*)challenges, (const uint4*)sc, idx, bt).reduce();");
    E.line("    memcpy(out + 8 * idx, v.l, 32);");
    E.line("    bt = bt * w;");
    E.line("  }");
    E.line("}");
  } else {
    E.line("extern \"C\" __global__ void __launch_bounds__(128, 4) zk_evalh_gen(const uint4* res, const uint4* wit, const uint4* challenges, const uint4* sc,");
    E.line("                                                                 const uint4* tw, uint4* out, unsigned long long row0, unsigned long long row1) {");
    E.line("  const size_t idx = row0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x;");
    E.line("  if (idx >= row1) return;");
    E.fmt("  const size_t half = %lluull;\n", (unsigned long long)((1ull << P.ext_k) >> 1));
    E.line("  Fr bt = SC_ONE;");
    if (P.ext_k > 0) {
      E.line("  bt = Fr::load(tw + 2 * (idx & (half - 1)));      // extended_omega^idx = +-table[idx mod ext/2]");
      E.line("  if (idx >= half) bt = bt.neg();");
    }
    E.line("  zk_evalh_row(res, wit, challenges, sc, idx, bt).store(out + 2 * idx);");
    E.line("}");
  }
  return E.s;
}

}  // namespace zk
