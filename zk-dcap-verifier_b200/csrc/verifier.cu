// verifier.cu -- halo2 verify_proof (KZG / SHPLONK / Blake2b) on the host, with a BN254 optimal-ate pairing.
//
// Restates halo2_proofs plonk/verifier.rs `verify_proof`, plonk/{vanishing,permutation,lookup}/verifier.rs,
// poly/kzg/multiopen/shplonk/verifier.rs and poly/kzg/strategy.rs (SingleStrategy -> DualMSM::check) ([UPSTREAM] @ c254c75f,
// Cargo.lock:1314-1316), as the reference runs them at circuits/src/sgx_dcap_verifier.rs:827-843 and
// bin/src/main.rs:244-251.  Verification is a few hundred field operations per
// column, ~150 scalar multiplications and two pairings: host code (as in the reference), built on the host
// compilation of ff.cuh / ec.cuh.  Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), Fq2 = Fq[i] / (i^2 + 1), xi = 9 + i = w^6.
#include <array>
#include <algorithm>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "transcript.hpp"

using namespace zk;

namespace {

// ---- small helpers ---------------------------------------------------------------------------------------------------
Fq fq_small(uint64_t v) { Fq r = Fq::zero(); r.l[0] = (uint32_t)v; r.l[1] = (uint32_t)(v >> 32); return r.to_mont(); }
Fr fr_small_(uint64_t v) { Fr r = Fr::zero(); r.l[0] = (uint32_t)v; r.l[1] = (uint32_t)(v >> 32); return r.to_mont(); }
template <class F> F f_canon_bytes(const uint8_t* p) { F v; memcpy(v.l, p, 32); return v.to_mont(); }
template <class F> bool f_is_canonical(const uint8_t* p) {
  F v, t, m = F::template kp<1>();
  memcpy(v.l, p, 32);
  return sub8(t.l, v.l, m.l) != 0;          // borrow <=> v < p
}
bool fr_lt(const Fr& a, const Fr& b) {
  Fr x = a.from_mont(), y = b.from_mont();
  for (int i = 7; i >= 0; i--) if (x.l[i] != y.l[i]) return x.l[i] < y.l[i];
  return false;
}
bool fr_eq(const Fr& a, const Fr& b) { return (a - b).is_zero(); }

// ---- Fq2 -----------------------------------------------------------------------------------------------------------
struct Fq2 { Fq c0, c1; };
Fq2 f2(const Fq& a, const Fq& b) { Fq2 r; r.c0 = a; r.c1 = b; return r; }
Fq2 f2_add(const Fq2& a, const Fq2& b) { return f2(a.c0 + b.c0, a.c1 + b.c1); }
Fq2 f2_sub(const Fq2& a, const Fq2& b) { return f2(a.c0 - b.c0, a.c1 - b.c1); }
Fq2 f2_mul(const Fq2& a, const Fq2& b) { return f2(a.c0 * b.c0 - a.c1 * b.c1, a.c0 * b.c1 + a.c1 * b.c0); }
Fq2 f2_neg(const Fq2& a) { return f2(a.c0.neg(), a.c1.neg()); }
Fq2 f2_conj(const Fq2& a) { return f2(a.c0, a.c1.neg()); }
Fq2 f2_scalar(const Fq2& a, const Fq& k) { return f2(a.c0 * k, a.c1 * k); }
Fq2 f2_inv(const Fq2& a) { Fq d = Fq::inv(a.c0 * a.c0 + a.c1 * a.c1); return f2(a.c0 * d, a.c1.neg() * d); }
bool f2_is_zero(const Fq2& a) { return a.c0.is_zero() && a.c1.is_zero(); }
bool f2_eq(const Fq2& a, const Fq2& b) { return f2_is_zero(f2_sub(a, b)); }

struct G2 { Fq2 x, y; bool inf; };
// slope of the line through p and q (tangent when equal); vertical = true when there is none
bool g2_slope(const G2& p, const G2& q, Fq2* m) {
  if (!f2_eq(p.x, q.x)) { *m = f2_mul(f2_sub(q.y, p.y), f2_inv(f2_sub(q.x, p.x))); return true; }
  if (f2_eq(p.y, q.y) && !f2_is_zero(p.y)) {
    Fq2 xx = f2_mul(p.x, p.x);
    *m = f2_mul(f2_add(f2_add(xx, xx), xx), f2_inv(f2_add(p.y, p.y)));
    return true;
  }
  return false;
}
G2 g2_add(const G2& p, const G2& q) {
  if (p.inf) return q;
  if (q.inf) return p;
  Fq2 m;
  if (!g2_slope(p, q, &m)) { G2 r; r.inf = true; r.x = r.y = f2(Fq::zero(), Fq::zero()); return r; }
  G2 r; r.inf = false;
  r.x = f2_sub(f2_sub(f2_mul(m, m), p.x), q.x);
  r.y = f2_sub(f2_mul(m, f2_sub(p.x, r.x)), p.y);
  return r;
}
bool g2_on_curve(const G2& q) {      // y^2 = x^3 + 3 / (9 + i)
  if (q.inf) return true;
  Fq2 b = f2_mul(f2(fq_small(3), Fq::zero()), f2_inv(f2(fq_small(9), Fq::one())));
  return f2_eq(f2_mul(q.y, q.y), f2_add(f2_mul(f2_mul(q.x, q.x), q.x), b));
}

// ---- Fq12 as 12 coefficients of w -----------------------------------------------------------------------------------------
typedef std::array<Fq, 12> F12;
F12 f12_one() { F12 r; for (auto& c : r) c = Fq::zero(); r[0] = Fq::one(); return r; }
F12 f12_mul(const F12& a, const F12& b) {
  static const Fq k18 = fq_small(18), k82 = fq_small(82);
  Fq t[23];
  for (auto& c : t) c = Fq::zero();
  for (int i = 0; i < 12; i++) {
    if (a[i].is_zero()) continue;
    for (int j = 0; j < 12; j++) {
      if (b[j].is_zero()) continue;
      t[i + j] = t[i + j] + a[i] * b[j];
    }
  }
  for (int k = 22; k >= 12; k--) {          // w^12 = 18 w^6 - 82
    if (t[k].is_zero()) continue;
    t[k - 6] = t[k - 6] + t[k] * k18;
    t[k - 12] = t[k - 12] - t[k] * k82;
  }
  F12 r;
  for (int i = 0; i < 12; i++) r[i] = t[i];
  return r;
}
bool f12_is_one(const F12& a) {
  if (!(a[0] - Fq::one()).is_zero()) return false;
  for (int i = 1; i < 12; i++) if (!a[i].is_zero()) return false;
  return true;
}
F12 embed(const Fq2& c) {               // i = w^6 - 9
  F12 r; for (auto& x : r) x = Fq::zero();
  r[0] = c.c0 - c.c1 * fq_small(9); r[6] = c.c1;
  return r;
}
F12 shift(const F12& a, int k) { F12 w; for (auto& x : w) x = Fq::zero(); w[k] = Fq::one(); return f12_mul(a, w); }
// value at the G1 point (px, py) of the line with untwisted slope m through the twist point t (vertical when !has_m)
F12 line(bool has_m, const Fq2& m, const G2& t, const Fq& px, const Fq& py) {
  F12 X1 = shift(embed(t.x), 2);
  F12 dx; for (int i = 0; i < 12; i++) dx[i] = X1[i].neg();
  dx[0] = dx[0] + px;
  if (!has_m) return dx;
  F12 Y1 = shift(embed(t.y), 3);
  F12 dy; for (int i = 0; i < 12; i++) dy[i] = Y1[i].neg();
  dy[0] = dy[0] + py;
  F12 mm = f12_mul(shift(embed(m), 1), dx), r;
  for (int i = 0; i < 12; i++) r[i] = mm[i] - dy[i];
  return r;
}
Fq fq_hex(const char* h) {               // canonical big-endian hex -> Montgomery
  Fq v = Fq::zero();
  size_t n = strlen(h);
  for (size_t i = 0; i < n; i++) {
    char c = h[n - 1 - i];
    uint32_t d = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
    v.l[i / 8] |= d << (4 * (i % 8));
  }
  return v.to_mont();
}
G2 g2_frobenius(const G2& q) {           // (conj(x) * xi^((p-1)/3), conj(y) * xi^((p-1)/2))
  static const Fq2 g2c = f2(fq_hex("2fb347984f7911f74c0bec3cf559b143b78cc310c2c3330c99e39557176f553d"),
                            fq_hex("16c9e55061ebae204ba4cc8bd75a079432ae2a1d0b7c9dce1665d51c640fcba2"));
  static const Fq2 g3c = f2(fq_hex("063cf305489af5dcdc5ec698b6e2f9b9dbaae0eda9c95998dc54014671a0135a"),
                            fq_hex("07c03cbcac41049a0704b5a7ec796f2b21807dc98fa25bd282d37f632623b0e3"));
  G2 r; r.inf = q.inf;
  r.x = f2_mul(f2_conj(q.x), g2c); r.y = f2_mul(f2_conj(q.y), g3c);
  return r;
}
F12 miller_loop(const G2& q, const Fq& px, const Fq& py) {
  const uint64_t ate = 29793968203157093288ULL;       // 6u + 2
  F12 f = f12_one();
  G2 r = q;
  for (int i = 63; i >= 0; i--) {
    Fq2 m; bool has = g2_slope(r, r, &m);
    f = f12_mul(f12_mul(f, f), line(has, m, r, px, py));
    r = g2_add(r, r);
    if ((ate >> i) & 1) {
      has = g2_slope(r, q, &m);
      f = f12_mul(f, line(has, m, r, px, py));
      r = g2_add(r, q);
    }
  }
  G2 q1 = g2_frobenius(q), q2 = g2_frobenius(q1);
  q2.y = f2_neg(q2.y);
  Fq2 m; bool has = g2_slope(r, q1, &m);
  f = f12_mul(f, line(has, m, r, px, py));
  r = g2_add(r, q1);
  has = g2_slope(r, q2, &m);
  f = f12_mul(f, line(has, m, r, px, py));
  return f;
}
const uint64_t kFinalExp[44] = {   // (p^12 - 1) / r, little-endian 64-bit words
    0x86964b64ca86f120ULL, 0x40a4efb7e54523a4ULL, 0x837fa97896e84abbULL, 0x361102b6b9b2b918ULL,
    0xc0de81def35692daULL, 0xbe04c7e8a6c3c760ULL, 0xd766f9c9d570bb7fULL, 0xc230974d83561841ULL,
    0x5bba1668c3be69a3ULL, 0x7f3811c410526294ULL, 0x29baee7ddadda71cULL, 0xbf813b8d145da900ULL,
    0x641bbadf423f9a2cULL, 0xa80bb4ea44eacc5eULL, 0xcd65664814fde37cULL, 0x4a0364b9580291d2ULL,
    0xee93dfb10826f0ddULL, 0x6b42db8dc5514724ULL, 0xbb10cf430b0f3785ULL, 0x40494e406f804216ULL,
    0x55cfe107acf3aafbULL, 0x2088ec80e0ebae87ULL, 0x846a3ed011a337a0ULL, 0x48a45a4a1e3a5195ULL,
    0xe5664568dfc50e16ULL, 0xab6a41294c0cc4ebULL, 0x82d0d602d268c7daULL, 0x6668449aed3cc48aULL,
    0x5062cd0fb2015dfcULL, 0x7f2940a8b1ddb3d1ULL, 0x77f5b63a2a226448ULL, 0xfef0781361e443aeULL,
    0xf977870e88d5c6c8ULL, 0x790364a61f676baaULL, 0x5887e72eceaddea3ULL, 0x1377e563a09a1b70ULL,
    0x0c54efee1bd8c3b2ULL, 0x3ec3d15ad524d8f7ULL, 0xdaf15466b2383a5dULL, 0xe1e30a73bb94fec0ULL,
    0x6a1c71015f3f7be2ULL, 0x842d43bf6369b1ffULL, 0x20fddadf107d20bcULL, 0x0000002f4b6dc970ULL,
};
F12 final_exponentiation(const F12& f) {
  F12 r = f12_one();
  for (int w = 43; w >= 0; w--)
    for (int b = 63; b >= 0; b--) {
      r = f12_mul(r, r);
      if ((kFinalExp[w] >> b) & 1) r = f12_mul(r, f);
    }
  return r;
}

// ---- G1 on the host ---------------------------------------------------------------------------------------------------
struct P1 { Fq x, y; bool inf; };
XYZZ g1_scalar_mul(const P1& p, const Fr& s) {
  XYZZ acc = XYZZ::identity();
  if (p.inf) return acc;
  Affine a; a.x = p.x; a.y = p.y;
  Fr c = s.from_mont();
  for (int i = 255; i >= 0; i--) {
    acc = acc.dbl();
    if ((c.l[i >> 5] >> (i & 31)) & 1) acc = acc.madd(a);
  }
  return acc;
}
bool g1_on_curve(const P1& p) { return p.inf || (p.y.sqr() - (p.x.sqr() * p.x + fq_small(3))).is_zero(); }

// ---- transcript read ------------------------------------------------------------------------------------------------------
struct TranscriptRead {
  TranscriptWrite t;         // state + the same absorb functions; its proof buffer is unused
  const uint8_t* buf; size_t len, pos = 0;
  explicit TranscriptRead(int kind) : t(kind) {}
  bool read_point(P1* out) {
    if (t.kind == ZK_TRANSCRIPT_EVM) {                           // 64 bytes: x, y big-endian
      if (pos + 64 > len) return false;
      uint8_t bx[32], by[32];
      for (int i = 0; i < 32; i++) { bx[i] = buf[pos + 31 - i]; by[i] = buf[pos + 63 - i]; }
      pos += 64;
      if (!f_is_canonical<Fq>(bx) || !f_is_canonical<Fq>(by)) return false;
      out->x = f_canon_bytes<Fq>(bx); out->y = f_canon_bytes<Fq>(by); out->inf = false;
      if (out->x.is_zero() && out->y.is_zero()) return false;
      if (!g1_on_curve(*out)) return false;
      t.common_point(out->x, out->y, false);
      return true;
    }
    if (pos + 32 > len) return false;
    uint8_t b[32]; memcpy(b, buf + pos, 32); pos += 32;
    bool allz = true; for (int i = 0; i < 32; i++) allz = allz && b[i] == 0;
    if (allz) return false;                                    // the identity cannot be absorbed
    uint32_t sign;
    if (t.kind == ZK_TRANSCRIPT_POSEIDON) {                      // halo2curves >= 0.4: bit 7 = identity flag, bit 6 = parity of y
      if (b[31] & 0x80) return false;
      sign = (b[31] >> 6) & 1u; b[31] &= 0x3F;
    } else {
      sign = b[31] >> 7; b[31] &= 0x7F;
    }
    if (!f_is_canonical<Fq>(b)) return false;
    Fq x = f_canon_bytes<Fq>(b);
    Fq rhs = x.sqr() * x + fq_small(3);
    const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
    Fq y = Fq::pow(rhs, e);
    if (!(y.sqr() - rhs).is_zero()) return false;
    if ((y.from_mont().l[0] & 1u) != sign) y = y.neg();
    out->x = x; out->y = y; out->inf = false;
    t.common_point(x, y, false);
    return true;
  }
  bool read_scalar(Fr* out) {
    if (pos + 32 > len) return false;
    uint8_t b[32];
    if (t.kind == ZK_TRANSCRIPT_EVM) for (int i = 0; i < 32; i++) b[i] = buf[pos + 31 - i];
    else memcpy(b, buf + pos, 32);
    if (!f_is_canonical<Fr>(b)) return false;
    *out = f_canon_bytes<Fr>(b); pos += 32;
    t.common_scalar(*out);
    return true;
  }
};

// ---- expression graphs evaluated at the opening point -----------------------------------------------------------------------
struct EvalCtx {
  const zk_circuit_desc* cd;
  const std::vector<Fr>* advice; const std::vector<Fr>* fixed; const std::vector<Fr>* instance; const std::vector<Fr>* challenges;
  Fr beta, gamma, theta, y;
  bool ok = true;
};
int query_index(const zk_query* q, uint32_t n, uint32_t col, int32_t rot) {
  for (uint32_t i = 0; i < n; i++) if (q[i].column == col && q[i].rotation == rot) return (int)i;
  return -1;
}
Fr graph_at_point(const zk_graph& g, EvalCtx& E, const Fr& prev) {
  std::vector<Fr> inter(g.num_intermediates + 1, Fr::zero());
  auto get = [&](const zk_value_source& s) -> Fr {
    switch (s.kind) {
      case ZK_VS_CONSTANT: return f_from_abi<Fr>(&g.constants[s.a]);
      case ZK_VS_INTERMEDIATE: return inter[s.a];
      case ZK_VS_FIXED: { int i = query_index(E.cd->fixed_queries, E.cd->n_fixed_queries, s.a, g.rotations[s.b]); if (i < 0) { E.ok = false; return Fr::zero(); } return (*E.fixed)[i]; }
      case ZK_VS_ADVICE: { int i = query_index(E.cd->advice_queries, E.cd->n_advice_queries, s.a, g.rotations[s.b]); if (i < 0) { E.ok = false; return Fr::zero(); } return (*E.advice)[i]; }
      case ZK_VS_INSTANCE: { int i = query_index(E.cd->instance_queries, E.cd->n_instance_queries, s.a, g.rotations[s.b]); if (i < 0) { E.ok = false; return Fr::zero(); } return (*E.instance)[i]; }
      case ZK_VS_CHALLENGE: return s.a < E.challenges->size() ? (*E.challenges)[s.a] : Fr::zero();
      case ZK_VS_BETA: return E.beta;
      case ZK_VS_GAMMA: return E.gamma;
      case ZK_VS_THETA: return E.theta;
      case ZK_VS_Y: return E.y;
      default: return prev;
    }
  };
  Fr last = Fr::zero();
  for (uint32_t i = 0; i < g.n_calcs; i++) {
    const zk_calculation& c = g.calcs[i];
    Fr a = get(c.a), r;
    switch (c.kind) {
      case ZK_CALC_ADD: r = a + get(c.b); break;
      case ZK_CALC_SUB: r = a - get(c.b); break;
      case ZK_CALC_MUL: r = a * get(c.b); break;
      case ZK_CALC_SQUARE: r = a * a; break;
      case ZK_CALC_DOUBLE: r = a + a; break;
      case ZK_CALC_NEGATE: r = a.neg(); break;
      case ZK_CALC_HORNER: {
        Fr f = get(c.b);
        r = a;
        for (uint32_t p = 0; p < c.parts_len; p++) r = r * f + get(g.horner_parts[c.parts_off + p]);
        break;
      }
      default: r = a; break;
    }
    inter[c.target] = r;
    last = r;
  }
  return last;
}

// The circuit description comes from the caller: every index graph_at_point / zk_verify_proof will follow is checked up
// front (the prover side rejects the same things in compile_graph / zk_keygen), so a malformed description is an argument
// error, never an out-of-bounds read.
bool source_ok(const zk_value_source& s, const zk_graph& g, const zk_evaluator_desc& D) {
  switch (s.kind) {
    case ZK_VS_CONSTANT: return s.a < g.n_constants && g.constants;
    case ZK_VS_INTERMEDIATE: return s.a <= g.num_intermediates;
    case ZK_VS_FIXED: return s.a < D.n_fixed && s.b < g.n_rotations;
    case ZK_VS_ADVICE: return s.a < D.n_advice && s.b < g.n_rotations;
    case ZK_VS_INSTANCE: return s.a < D.n_instance && s.b < g.n_rotations;
    case ZK_VS_CHALLENGE: case ZK_VS_BETA: case ZK_VS_GAMMA: case ZK_VS_THETA: case ZK_VS_Y: case ZK_VS_PREVIOUS_VALUE: return true;
    default: return false;
  }
}
bool graph_ok(const zk_graph& g, const zk_evaluator_desc& D) {
  if ((g.n_calcs && !g.calcs) || (g.n_rotations && !g.rotations) || (g.n_horner_parts && !g.horner_parts)) return false;
  for (uint32_t i = 0; i < g.n_calcs; i++) {
    const zk_calculation& c = g.calcs[i];
    if (c.target > g.num_intermediates || c.kind > ZK_CALC_STORE || !source_ok(c.a, g, D)) return false;
    const bool binary = c.kind == ZK_CALC_ADD || c.kind == ZK_CALC_SUB || c.kind == ZK_CALC_MUL || c.kind == ZK_CALC_HORNER;
    if (binary && !source_ok(c.b, g, D)) return false;
    if (c.kind == ZK_CALC_HORNER) {
      if ((uint64_t)c.parts_off + c.parts_len > g.n_horner_parts) return false;
      for (uint32_t p = 0; p < c.parts_len; p++) if (!source_ok(g.horner_parts[c.parts_off + p], g, D)) return false;
    }
  }
  return true;
}
bool circuit_ok(const zk_circuit_desc* cd, const zk_verifier_params* vp) {
  const zk_evaluator_desc& D = cd->ev;
  if ((D.n_fixed && !vp->fixed_commitments) || (D.n_perm_columns && (!vp->permutation_commitments || !D.perm_columns))) return false;
  if ((D.n_lookups && (!D.lookups || !cd->lookup_inputs || !cd->lookup_tables))) return false;
  if ((cd->n_advice_queries && !cd->advice_queries) || (cd->n_fixed_queries && !cd->fixed_queries) ||
      (cd->n_instance_queries && !cd->instance_queries)) return false;
  for (uint32_t q = 0; q < cd->n_advice_queries; q++) if (cd->advice_queries[q].column >= D.n_advice) return false;
  for (uint32_t q = 0; q < cd->n_fixed_queries; q++) if (cd->fixed_queries[q].column >= D.n_fixed) return false;
  for (uint32_t q = 0; q < cd->n_instance_queries; q++) if (cd->instance_queries[q].column >= D.n_instance) return false;
  for (uint32_t c = 0; c < D.n_perm_columns; c++) {
    const zk_column& pc = D.perm_columns[c];
    const uint32_t lim = pc.column_type == ZK_COL_ADVICE ? D.n_advice : pc.column_type == ZK_COL_FIXED ? D.n_fixed :
                         pc.column_type == ZK_COL_INSTANCE ? D.n_instance : 0;
    if (pc.index >= lim) return false;
  }
  if (!graph_ok(D.custom_gates, D)) return false;
  for (uint32_t l = 0; l < D.n_lookups; l++)
    if (!graph_ok(D.lookups[l], D) || !graph_ok(cd->lookup_inputs[l], D) || !graph_ok(cd->lookup_tables[l], D)) return false;
  return true;
}

std::vector<Fr> lagrange_interp(const std::vector<Fr>& pts, const std::vector<Fr>& evals) {
  size_t m = pts.size();
  std::vector<Fr> out(m, Fr::zero());
  for (size_t j = 0; j < m; j++) {
    std::vector<Fr> num(1, Fr::one());
    Fr den = Fr::one();
    for (size_t i = 0; i < m; i++) {
      if (i == j) continue;
      std::vector<Fr> nxt(num.size() + 1, Fr::zero());
      for (size_t t = 0; t < num.size(); t++) { nxt[t] = nxt[t] - num[t] * pts[i]; nxt[t + 1] = nxt[t + 1] + num[t]; }
      num.swap(nxt);
      den = den * (pts[j] - pts[i]);
    }
    Fr f = evals[j] * Fr::inv(den);
    for (size_t t = 0; t < num.size(); t++) out[t] = out[t] + num[t] * f;
  }
  return out;
}
Fr eval_small_poly(const std::vector<Fr>& c, const Fr& x) { Fr a = Fr::zero(); for (size_t i = c.size(); i-- > 0;) a = a * x + c[i]; return a; }

}  // namespace

// ---- G2 encodings (ParamsKZG.g2 / s_g2 as the SRS files hold them) -> the verifier's uncompressed form -----------------
// [UPSTREAM? halo2curves G2Affine::to_bytes, same macro as G1]: x = (c0, c1) as 2 x 32 little-endian bytes; the flag bits
// live in the last byte (byte 63): 0.3.1 -- parity of y's first byte (y.c0) in bit 7, identity = zeros; >= 0.4 -- parity in
// bit 6, identity flag in bit 7.  RawBytes = x.c0, x.c1, y.c0, y.c1 as Montgomery limbs (128 B).
namespace {
bool fq_sqrt(const Fq& a, Fq* out) {                 // p = 3 mod 4: a^((p+1)/4)
  uint32_t e[8] = {FqParams::M0 + 1, FqParams::M1, FqParams::M2, FqParams::M3, FqParams::M4, FqParams::M5, FqParams::M6, FqParams::M7};
  for (int i = 0; i < 8; i++) e[i] = (e[i] >> 2) | (i < 7 ? (e[i + 1] << 30) : 0);
  Fq r = Fq::pow(a, e);
  if (!((r * r) - a).is_zero()) return false;
  *out = r;
  return true;
}
bool f2_sqrt(const Fq2& a, Fq2* out) {
  Fq two_inv = Fq::inv(fq_small(2));
  Fq2 r;
  if (a.c1.is_zero()) {
    Fq t;
    if (fq_sqrt(a.c0, &t)) r = f2(t, Fq::zero());
    else if (fq_sqrt(a.c0.neg(), &t)) r = f2(Fq::zero(), t);
    else return false;
  } else {
    Fq n;
    if (!fq_sqrt(a.c0 * a.c0 + a.c1 * a.c1, &n)) return false;
    Fq x0;
    if (!fq_sqrt((a.c0 + n) * two_inv, &x0) && !fq_sqrt((a.c0 - n) * two_inv, &x0)) return false;
    r = f2(x0, a.c1 * Fq::inv(x0 + x0));
  }
  if (!f2_eq(f2_mul(r, r), a)) return false;
  *out = r;
  return true;
}
Fq2 g2_b() { return f2_mul(f2(fq_small(3), Fq::zero()), f2_inv(f2(fq_small(9), Fq::one()))); }
void put_canon(uint8_t* o, const Fq& v) { Fq c = v.from_mont(); memcpy(o, c.l, 32); }
}  // namespace

// The Poseidon permutation / sponge of the transcript (poseidon.hpp), exposed so that the parameter generation can be checked
// against published values (any width t <= 16) and the sponge against the CPU restatement.  Values in the ABI's Montgomery form.
extern "C" int zk_poseidon_permute(zk_fr* state, uint32_t t, uint32_t r_f, uint32_t r_p) {
  if (!state || t < 2 || t > 16 || (r_f & 1u) || r_f > 16 || r_p > 128) return ZK_ERR_ARG;
  const bool std_spec = t == (uint32_t)PoseidonSpec::T && r_f == (uint32_t)PoseidonSpec::RF && r_p == (uint32_t)PoseidonSpec::RP;
  PoseidonSpec own;
  if (!std_spec) own = poseidon_make_spec((int)t, (int)r_f, (int)r_p);
  const PoseidonSpec& S = std_spec ? poseidon_transcript_spec() : own;
  std::vector<Fr> st(t), plain(t);
  for (uint32_t i = 0; i < t; i++) plain[i] = st[i] = f_from_abi<Fr>(&state[i]);
  poseidon_permute(S, st.data());
  poseidon_permute_plain(S, plain.data());                    // self-check: the optimised form is the same function
  for (uint32_t i = 0; i < t; i++) if (st[i] != plain[i]) return ZK_ERR_STATE;
  for (uint32_t i = 0; i < t; i++) f_to_abi(st[i], &state[i]);
  return ZK_OK;
}
extern "C" int zk_poseidon_hash(const zk_fr* elements, size_t count, zk_fr* out, uint32_t squeezes) {
  if ((!elements && count) || !out) return ZK_ERR_ARG;
  PoseidonSponge sp;
  for (size_t i = 0; i < count; i++) sp.update(f_from_abi<Fr>(&elements[i]));
  for (uint32_t s = 0; s < (squeezes ? squeezes : 1u); s++) f_to_abi(sp.squeeze(), &out[s]);
  return ZK_OK;
}

extern "C" int zk_g2_decode(int format, const uint8_t* in, uint8_t out[128]) {
  if (!in || !out) return ZK_ERR_ARG;
  if (format == ZK_SERDE_RAW_BYTES) {
    Fq v[4];
    for (int i = 0; i < 4; i++) {
      if (!f_is_canonical<Fq>(in + 32 * i)) return ZK_ERR_ARG;      // Montgomery limbs are < p too
      memcpy(v[i].l, in + 32 * i, 32);
    }
    G2 q; q.inf = false; q.x = f2(v[0], v[1]); q.y = f2(v[2], v[3]);
    if (!g2_on_curve(q)) return ZK_ERR_ARG;
    for (int i = 0; i < 4; i++) put_canon(out + 32 * i, v[i]);
    return ZK_OK;
  }
  if (format != ZK_SERDE_PROCESSED && format != ZK_SERDE_PROCESSED_V04) return ZK_ERR_ARG;
  uint8_t b[64];
  memcpy(b, in, 64);
  const bool v04 = format == ZK_SERDE_PROCESSED_V04;
  int sign;
  if (v04) {
    if (b[63] & 0x80) return ZK_ERR_ARG;              // the identity is no SRS element
    sign = (b[63] >> 6) & 1; b[63] &= 0x3F;
  } else {
    sign = b[63] >> 7; b[63] &= 0x7F;
  }
  if (!f_is_canonical<Fq>(b) || !f_is_canonical<Fq>(b + 32)) return ZK_ERR_ARG;
  Fq2 x = f2(f_canon_bytes<Fq>(b), f_canon_bytes<Fq>(b + 32));
  Fq2 y;
  if (!f2_sqrt(f2_add(f2_mul(f2_mul(x, x), x), g2_b()), &y)) return ZK_ERR_ARG;
  if ((int)(y.c0.from_mont().l[0] & 1) != sign) y = f2_neg(y);
  put_canon(out, x.c0); put_canon(out + 32, x.c1); put_canon(out + 64, y.c0); put_canon(out + 96, y.c1);
  return ZK_OK;
}

extern "C" int zk_g2_encode(int format, const uint8_t in[128], uint8_t* out) {
  if (!in || !out) return ZK_ERR_ARG;
  for (int i = 0; i < 4; i++) if (!f_is_canonical<Fq>(in + 32 * i)) return ZK_ERR_ARG;
  G2 q; q.inf = false;
  q.x = f2(f_canon_bytes<Fq>(in), f_canon_bytes<Fq>(in + 32)); q.y = f2(f_canon_bytes<Fq>(in + 64), f_canon_bytes<Fq>(in + 96));
  if (!g2_on_curve(q)) return ZK_ERR_ARG;
  if (format == ZK_SERDE_RAW_BYTES) {
    memcpy(out, q.x.c0.reduce().l, 32); memcpy(out + 32, q.x.c1.reduce().l, 32);
    memcpy(out + 64, q.y.c0.reduce().l, 32); memcpy(out + 96, q.y.c1.reduce().l, 32);
    return ZK_OK;
  }
  if (format != ZK_SERDE_PROCESSED && format != ZK_SERDE_PROCESSED_V04) return ZK_ERR_ARG;
  memcpy(out, in, 64);
  const uint8_t sign = in[64] & 1;                    // parity of y.c0
  out[63] |= (uint8_t)(sign << (format == ZK_SERDE_PROCESSED_V04 ? 6 : 7));
  return ZK_OK;
}

extern "C" int zk_verify_proof(const zk_circuit_desc* cd, const zk_verifier_params* vp, const zk_fr* const* instance_values,
                               const uint32_t* instance_lens, const uint8_t* proof, size_t proof_len, int* accepted) {
  if (!cd || !vp || (!proof && proof_len) || !accepted || !vp->g2 || !vp->s_g2 || (cd->ev.n_instance && (!instance_values || !instance_lens)))
    return ZK_ERR_ARG;
  *accepted = 0;
  const zk_evaluator_desc& D = cd->ev;
  const uint32_t k = D.k, bf = D.blinding_factors, d = D.degree, chunk = d - 2;
  if (d < 3 || k > 27) return ZK_ERR_ARG;
  if (!circuit_ok(cd, vp)) return ZK_ERR_ARG;
  const uint64_t n = (uint64_t)1 << k;
  const uint32_t L = D.n_lookups, NP = D.n_perm_columns, NS = NP ? (NP + chunk - 1) / chunk : 0;
  // omega of the domain: ROOT_OF_UNITY^(2^(28-k))
  const uint32_t root_c[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
  Fr omega; memcpy(omega.l, root_c, 32); omega = omega.to_mont();
  for (uint32_t i = k; i < 28; i++) omega = omega * omega;
  Fr omega_inv = Fr::inv(omega);
  auto rotate = [&](const Fr& x, int32_t r) { return x * (r >= 0 ? Fr::pow_u64(omega, (uint64_t)r) : Fr::pow_u64(omega_inv, (uint64_t)(-(int64_t)r))); };
  const uint32_t delta_c[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
  Fr delta; memcpy(delta.l, delta_c, 32); delta = delta.to_mont();

  if (vp->transcript_kind != ZK_TRANSCRIPT_BLAKE2B && vp->transcript_kind != ZK_TRANSCRIPT_EVM && vp->transcript_kind != ZK_TRANSCRIPT_POSEIDON)
    return ZK_ERR_ARG;
  TranscriptRead tr(vp->transcript_kind); tr.buf = proof; tr.len = proof_len;
  tr.t.common_scalar(f_from_abi<Fr>(&cd->transcript_repr));
  for (uint32_t i = 0; i < D.n_instance; i++) {
    if (instance_lens[i] > n - (bf + 1)) return ZK_OK;
    for (uint32_t j = 0; j < instance_lens[i]; j++) tr.t.common_scalar(f_from_abi<Fr>(&instance_values[i][j]));
  }
#define RD_P(dst) do { if (!tr.read_point(&(dst))) return ZK_OK; } while (0)
#define RD_S(dst) do { if (!tr.read_scalar(&(dst))) return ZK_OK; } while (0)
  std::vector<P1> advice_c(D.n_advice);
  for (auto& p : advice_c) RD_P(p);
  std::vector<Fr> challenges(D.n_challenges);
  for (auto& c : challenges) c = tr.t.squeeze_challenge();
  Fr theta = tr.t.squeeze_challenge();
  std::vector<P1> pin_c(L), ptab_c(L), lz_c(L);
  for (uint32_t l = 0; l < L; l++) { RD_P(pin_c[l]); RD_P(ptab_c[l]); }
  Fr beta = tr.t.squeeze_challenge(), gamma = tr.t.squeeze_challenge();
  std::vector<P1> perm_c(NS);
  for (auto& p : perm_c) RD_P(p);
  for (uint32_t l = 0; l < L; l++) RD_P(lz_c[l]);
  P1 random_c; RD_P(random_c);
  Fr y = tr.t.squeeze_challenge();
  std::vector<P1> h_c(d - 1);
  for (auto& p : h_c) RD_P(p);
  Fr x = tr.t.squeeze_challenge();
  Fr xn = Fr::pow_u64(x, n);
  // l_i(x) = omega^i (x^n - 1) / (n (x - omega^i))
  Fr n_inv = Fr::inv(fr_small_(n));
  auto l_i = [&](int32_t i) { Fr wi = rotate(Fr::one(), i); return Fr::inv(x - wi) * ((xn - Fr::one()) * n_inv) * wi; };
  // instance evaluations (KZG: computed by the verifier)
  std::vector<Fr> instance_evals(cd->n_instance_queries, Fr::zero());
  for (uint32_t q = 0; q < cd->n_instance_queries; q++) {
    uint32_t col = cd->instance_queries[q].column; int32_t r = cd->instance_queries[q].rotation;
    if (col >= D.n_instance) return ZK_ERR_ARG;
    Fr acc = Fr::zero();
    for (uint32_t j = 0; j < instance_lens[col]; j++) acc = acc + f_from_abi<Fr>(&instance_values[col][j]) * l_i((int32_t)j - r);
    instance_evals[q] = acc;
  }
  std::vector<Fr> advice_evals(cd->n_advice_queries), fixed_evals(cd->n_fixed_queries), sigma_evals(NP);
  for (auto& e : advice_evals) RD_S(e);
  for (auto& e : fixed_evals) RD_S(e);
  Fr random_eval; RD_S(random_eval);
  for (auto& e : sigma_evals) RD_S(e);
  struct PE { Fr x, next, last; };
  std::vector<PE> pe(NS);
  for (uint32_t s = 0; s < NS; s++) { RD_S(pe[s].x); RD_S(pe[s].next); if (s + 1 < NS) RD_S(pe[s].last); }
  struct LE { Fr z, z_next, pin, pin_inv, ptab; };
  std::vector<LE> le(L);
  for (auto& e : le) { RD_S(e.z); RD_S(e.z_next); RD_S(e.pin); RD_S(e.pin_inv); RD_S(e.ptab); }

  // ---- expected h(x): the expressions folded with y in evaluate_h's order ---------------------------------------------------
  Fr l_last = l_i(-(int32_t)(bf + 1)), l_blind = Fr::zero(), l_0 = l_i(0);
  for (int32_t i = -(int32_t)bf; i < 0; i++) l_blind = l_blind + l_i(i);
  Fr active = Fr::one() - (l_last + l_blind);
  EvalCtx E; E.cd = cd; E.advice = &advice_evals; E.fixed = &fixed_evals; E.instance = &instance_evals; E.challenges = &challenges;
  E.beta = beta; E.gamma = gamma; E.theta = theta; E.y = y;
  Fr h = graph_at_point(D.custom_gates, E, Fr::zero());     // ends with Horner(PreviousValue = 0; gate polynomials; y)
  auto fold = [&](const Fr& e) { h = h * y + e; };
  auto any_eval = [&](const zk_column& c) -> Fr {
    int i = c.column_type == ZK_COL_ADVICE ? query_index(cd->advice_queries, cd->n_advice_queries, c.index, 0)
          : c.column_type == ZK_COL_FIXED ? query_index(cd->fixed_queries, cd->n_fixed_queries, c.index, 0)
                                          : query_index(cd->instance_queries, cd->n_instance_queries, c.index, 0);
    if (i < 0) { E.ok = false; return Fr::zero(); }
    return c.column_type == ZK_COL_ADVICE ? advice_evals[i] : c.column_type == ZK_COL_FIXED ? fixed_evals[i] : instance_evals[i];
  };
  if (NS) {
    fold(l_0 * (Fr::one() - pe[0].x));
    fold((pe[NS - 1].x * pe[NS - 1].x - pe[NS - 1].x) * l_last);
    for (uint32_t s = 1; s < NS; s++) fold((pe[s].x - pe[s - 1].last) * l_0);
    Fr cur = beta * x;
    for (uint32_t s = 0; s < NS; s++) {
      uint32_t c0 = s * chunk, c1 = std::min(c0 + chunk, NP);
      Fr left = pe[s].next, right = pe[s].x;
      for (uint32_t c = c0; c < c1; c++) {
        Fr v = any_eval(D.perm_columns[c]);
        left = left * (v + beta * sigma_evals[c] + gamma);
        right = right * (v + cur + gamma);
        cur = cur * delta;
      }
      fold((left - right) * active);
    }
  }
  for (uint32_t l = 0; l < L; l++) {
    Fr tv = graph_at_point(D.lookups[l], E, Fr::zero());      // (compressed inputs + beta) * (compressed tables + gamma)
    const LE& e = le[l];
    fold(l_0 * (Fr::one() - e.z));
    fold(l_last * (e.z * e.z - e.z));
    fold((e.z_next * (e.pin + beta) * (e.ptab + gamma) - e.z * tv) * active);
    fold(l_0 * (e.pin - e.ptab));
    fold((e.pin - e.ptab) * (e.pin - e.pin_inv) * active);
  }
  if (!E.ok) return ZK_ERR_ARG;
  Fr expected_h = h * Fr::inv(xn - Fr::one());

  // ---- queries: commitments as small MSMs, identified by an id (pointer equality upstream) -------------------------------------
  struct Term { Fr s; P1 p; };
  std::vector<std::vector<Term>> comm;
  struct Q { int c; Fr pt, ev; };
  std::vector<Q> queries;
  auto single = [&](const P1& p) { comm.push_back({{Fr::one(), p}}); return (int)comm.size() - 1; };
  auto vk_point = [&](const zk_g1_affine& a) { P1 p; p.x = f_from_abi<Fq>(&a.x); p.y = f_from_abi<Fq>(&a.y); p.inf = p.x.is_zero() && p.y.is_zero(); return p; };
  std::vector<int> adv_id(D.n_advice), perm_id(NS), lz_id(L), lpin_id(L), lptab_id(L), fix_id(D.n_fixed), sig_id(NP);
  for (uint32_t i = 0; i < D.n_advice; i++) adv_id[i] = single(advice_c[i]);
  for (uint32_t s = 0; s < NS; s++) perm_id[s] = single(perm_c[s]);
  for (uint32_t l = 0; l < L; l++) { lz_id[l] = single(lz_c[l]); lpin_id[l] = single(pin_c[l]); lptab_id[l] = single(ptab_c[l]); }
  for (uint32_t i = 0; i < D.n_fixed; i++) fix_id[i] = single(vk_point(vp->fixed_commitments[i]));
  for (uint32_t i = 0; i < NP; i++) sig_id[i] = single(vk_point(vp->permutation_commitments[i]));
  Fr x_next = rotate(x, 1), x_inv = rotate(x, -1), x_last = rotate(x, -(int32_t)(bf + 1));
  for (uint32_t q = 0; q < cd->n_advice_queries; q++)
    queries.push_back({adv_id[cd->advice_queries[q].column], rotate(x, cd->advice_queries[q].rotation), advice_evals[q]});
  for (uint32_t s = 0; s < NS; s++) { queries.push_back({perm_id[s], x, pe[s].x}); queries.push_back({perm_id[s], x_next, pe[s].next}); }
  for (uint32_t s = NS; s-- > 0;) if (s + 1 < NS) queries.push_back({perm_id[s], x_last, pe[s].last});
  for (uint32_t l = 0; l < L; l++) {
    queries.push_back({lz_id[l], x, le[l].z}); queries.push_back({lpin_id[l], x, le[l].pin}); queries.push_back({lptab_id[l], x, le[l].ptab});
    queries.push_back({lpin_id[l], x_inv, le[l].pin_inv}); queries.push_back({lz_id[l], x_next, le[l].z_next});
  }
  for (uint32_t q = 0; q < cd->n_fixed_queries; q++)
    queries.push_back({fix_id[cd->fixed_queries[q].column], rotate(x, cd->fixed_queries[q].rotation), fixed_evals[q]});
  for (uint32_t i = 0; i < NP; i++) queries.push_back({sig_id[i], x, sigma_evals[i]});
  {
    std::vector<Term> hm; Fr pw = Fr::one();
    for (auto& c : h_c) { hm.push_back({pw, c}); pw = pw * xn; }
    comm.push_back(hm);
    queries.push_back({(int)comm.size() - 1, x, expected_h});
    queries.push_back({single(random_c), x, random_eval});
  }

  // ---- VerifierSHPLONK ---------------------------------------------------------------------------------------------------------
  std::vector<Fr> super;                                  // BTreeSet of points
  auto insert_sorted = [&](std::vector<Fr>& v, const Fr& p) {
    for (auto& e : v) if (fr_eq(e, p)) return;
    v.push_back(p);
    std::sort(v.begin(), v.end(), fr_lt);
  };
  std::vector<std::pair<int, std::vector<Fr>>> cset;      // (commitment, its points ascending)
  for (auto& q : queries) {
    insert_sorted(super, q.pt);
    bool found = false;
    for (auto& cs : cset) if (cs.first == q.c) { insert_sorted(cs.second, q.pt); found = true; break; }
    if (!found) cset.push_back({q.c, {q.pt}});
  }
  struct RS { std::vector<Fr> pts; std::vector<int> comms; std::vector<std::vector<Fr>> evals; };
  std::vector<RS> sets;
  auto same_set = [&](const std::vector<Fr>& a, const std::vector<Fr>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) if (!fr_eq(a[i], b[i])) return false;
    return true;
  };
  for (auto& cs : cset) {
    size_t si = 0;
    for (; si < sets.size(); si++) if (same_set(sets[si].pts, cs.second)) break;
    if (si == sets.size()) { RS r; r.pts = cs.second; sets.push_back(r); }
    std::vector<Fr> ev;
    for (auto& p : sets[si].pts) {
      bool ok = false;
      for (auto& q : queries) if (q.c == cs.first && fr_eq(q.pt, p)) { ev.push_back(q.ev); ok = true; break; }
      if (!ok) return ZK_ERR_STATE;
    }
    sets[si].comms.push_back(cs.first);
    sets[si].evals.push_back(ev);
  }
  Fr yy = tr.t.squeeze_challenge(), v = tr.t.squeeze_challenge();
  P1 h1; RD_P(h1);
  Fr u = tr.t.squeeze_challenge();
  P1 h2; RD_P(h2);
  if (tr.pos != tr.len) return ZK_OK;
#undef RD_P
#undef RD_S
  auto vanishing = [&](const std::vector<Fr>& roots, const Fr& z) { Fr a = Fr::one(); for (auto& r : roots) a = a * (z - r); return a; };
  XYZZ right = XYZZ::identity();
  Fr r_outer = Fr::zero(), z_0 = Fr::zero(), z_0_diff_inv = Fr::zero(), pv = Fr::one();
  for (size_t i = 0; i < sets.size(); i++) {
    std::vector<Fr> diffs;
    for (auto& p : super) { bool in = false; for (auto& q : sets[i].pts) in = in || fr_eq(p, q); if (!in) diffs.push_back(p); }
    Fr z_diff = vanishing(diffs, u);
    if (i == 0) {
      z_0 = vanishing(sets[i].pts, u);
      if (z_diff.is_zero()) return ZK_OK;
      z_0_diff_inv = Fr::inv(z_diff);
      z_diff = Fr::one();
    } else {
      z_diff = z_diff * z_0_diff_inv;
    }
    Fr r_inner = Fr::zero(), py = Fr::one();
    for (size_t j = 0; j < sets[i].comms.size(); j++) {
      r_inner = r_inner + py * eval_small_poly(lagrange_interp(sets[i].pts, sets[i].evals[j]), u);
      for (auto& t : comm[sets[i].comms[j]]) right = right.add(g1_scalar_mul(t.p, t.s * py * pv * z_diff));
      py = py * yy;
    }
    r_outer = r_outer + pv * r_inner * z_diff;
    pv = pv * v;
  }
  P1 g1 = vk_point(vp->g1);
  right = right.add(g1_scalar_mul(g1, r_outer.neg()));
  right = right.add(g1_scalar_mul(h1, z_0.neg()));
  right = right.add(g1_scalar_mul(h2, u));
  // DualMSM::check: e(h2, [s]G2) * e(-right, G2) == 1
  auto g2_from = [&](const uint8_t* b, G2* q) {
    for (int i = 0; i < 4; i++) if (!f_is_canonical<Fq>(b + 32 * i)) return false;
    q->x = f2(f_canon_bytes<Fq>(b), f_canon_bytes<Fq>(b + 32)); q->y = f2(f_canon_bytes<Fq>(b + 64), f_canon_bytes<Fq>(b + 96));
    q->inf = false;
    return g2_on_curve(*q);
  };
  G2 g2, sg2;
  if (!g2_from(vp->g2, &g2) || !g2_from(vp->s_g2, &sg2)) return ZK_ERR_ARG;
  F12 f = miller_loop(sg2, h2.x, h2.y);
  if (!right.is_identity()) {
    Affine ra = right.to_affine();
    f = f12_mul(f, miller_loop(g2, ra.x, ra.y.neg()));
  }
  *accepted = f12_is_one(final_exponentiation(f)) ? 1 : 0;
  return ZK_OK;
}
