// Host-side harness for ff.cuh / ec.cuh: exposes the limb-level algorithms (compiled for
// the CPU through their 64-bit emulation of the PTX carry chains) to the Python tests.
#include <cstring>
#include <cstddef>
struct uint4 { unsigned x, y, z, w; };
#include "../../zk-dcap-verifier_b200/csrc/ff.cuh"
using namespace zk;
template <class F>
static void run(int op, const unsigned* a, const unsigned* b, unsigned* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    F x, y, z;
    memcpy(x.l, a + 8 * i, 32);
    if (b) memcpy(y.l, b + 8 * i, 32); else y = F::zero();
    switch (op) {
      case 0: z = x + y; break;
      case 1: z = x - y; break;
      case 2: z = x * y; break;
      case 3: z = x.neg(); break;
      case 4: z = F::inv(x); break;
      case 5: z = x.sqr(); break;
      case 6: z = x.to_mont(); break;
      case 7: z = x.from_mont(); break;
      case 8: z = x.dbl(); break;
      // lazy-range stress: feed (x + p) (still < 2p) and multiply
      case 9: { F xp; add8(xp.l, x.l, F::template kp<1>().l); F yp; add8(yp.l, y.l, F::template kp<1>().l); z = xp * yp; break; }
      case 10: { F xp; add8(xp.l, x.l, F::template kp<1>().l); z = xp - y; break; }
      case 11: { F xp; add8(xp.l, x.l, F::template kp<1>().l); F yp; add8(yp.l, y.l, F::template kp<1>().l); z = xp + yp; break; }
      case 12: { F xp; add8(xp.l, x.l, F::template kp<1>().l); z = xp.sqr(); break; }      // squaring of a lazy value in [p, 2p)
      case 13: { z = x.sqr().sqr(); break; }                                                 // chained: the input of the second is lazy
      default: z = x; break;
    }
    z = z.reduce();
    memcpy(out + 8 * i, z.l, 32);
  }
}
extern "C" void ffh_op(int field, int op, const unsigned* a, const unsigned* b, unsigned* out, size_t n) {
  if (field == 0) run<Fr>(op, a, b, out, n); else run<Fq>(op, a, b, out, n);
}

#include "../../zk-dcap-verifier_b200/csrc/ec.cuh"
// XYZZ accumulate: acc = sum_i (+/-) pts[i] using madd; then a second pass mixing add/dbl.
// signs[i] != 0 negates.  out = normalised G1 (96 bytes).
extern "C" void ech_sum(const unsigned* pts, const unsigned char* signs, size_t n, unsigned* out, int mode) {
  XYZZ acc = XYZZ::identity();
  if (mode == 0) {
    for (size_t i = 0; i < n; i++) acc = acc.madd(Affine::load(pts + 16 * i).cneg(signs && signs[i]));
  } else {
    // pairwise: (p0 + p1) + (p2 + p3) ... through XYZZ + XYZZ adds, and doubling when equal
    XYZZ a = XYZZ::identity(), b = XYZZ::identity();
    for (size_t i = 0; i < n; i++) {
      Affine p = Affine::load(pts + 16 * i).cneg(signs && signs[i]);
      if (i & 1) b = b.madd(p); else a = a.madd(p);
    }
    acc = a.add(b);
    if (mode == 2) acc = acc.add(acc);   // exercises the doubling branch of add
    if (mode == 3) acc = acc.add(acc.neg());
  }
  store_g1_normalised(out, acc.to_affine());
}
extern "C" void ech_from_g1(const unsigned* g1, unsigned* out) {
  store_g1_normalised(out, load_g1_as_xyzz(g1).to_affine());
}
// digits[w] for w < W
extern "C" void ech_digits(const unsigned* canon, int c, int W, int* digits) {
  int carry = 0;
  for (int w = 0; w < W; w++) digits[w] = signed_digit(canon, w, c, carry);
  digits[W] = carry;
}
// the streaming form the digit kernels use (registers only)
extern "C" void ech_digits_stream(const unsigned* canon, int c, int W, int* digits) {
  unsigned l[8];
  memcpy(l, canon, 32);
  int carry = 0;
  for (int w = 0; w < W; w++) digits[w] = next_signed_digit(l, c, carry);
  digits[W] = carry;
}

// ---- evaluate_h interpreter on the host -------------------------------------------------------
#include "../../zk-dcap-verifier_b200/csrc/evalh_compile.hpp"
static int g_optimise = 1;
extern "C" void evh_set_optimise(int on) { g_optimise = on; }
// cols: n_slots host pointers in slot order (see compile_program); returns 0 or -1 (error text in err)
extern "C" int evh_run(const zk_evaluator_desc* desc, const void* const* cols, const void* challenges,
                       const unsigned* y, const unsigned* beta, const unsigned* gamma, const unsigned* theta,
                       const unsigned* ext_omega, unsigned* out, unsigned long long* stats, char* err) {
  CompiledEval C;
  if (!compile_program(*desc, C, g_optimise != 0)) { strncpy(err, C.error.c_str(), 200); return -1; }
  if (C.max_intermediates > 512) { strncpy(err, "too many intermediates", 200); return -1; }
  EvalScalars S;
  memcpy(S.beta.l, beta, 32); memcpy(S.gamma.l, gamma, 32); memcpy(S.theta.l, theta, 32); memcpy(S.y.l, y, 32);
  const uint32_t zeta_c[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
  const uint32_t delta_c[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
  Fr z, dl; memcpy(z.l, zeta_c, 32); memcpy(dl.l, delta_c, 32);
  S.delta_start = S.beta * z.to_mont(); S.delta = dl.to_mont(); S.one = Fr::one();
  Fr w; memcpy(w.l, ext_omega, 32);
  static EvalState<512> st;
  size_t E = (size_t)1 << C.prog.ext_k;
  Fr bt = Fr::one();
  for (size_t idx = 0; idx < E; idx++) {
    Fr v = evalh_row<512>(C.prog, C.code.data(), (const uint4* const*)cols, (const uint4*)C.consts.data(),
                          (const uint4*)challenges, S, st, idx, bt);
    v = v.reduce();
    memcpy(out + 8 * idx, v.l, 32);
    bt = bt * w;
  }
  stats[0] = C.muls_per_row; stats[1] = C.columns_read; stats[2] = C.n_slots; stats[3] = C.max_intermediates;
  return 0;
}

// ---- transcript.hpp (Blake2b-512, Blake2bWrite, Fr::from_bytes_wide, point compression) on the host ----------------
#include "../../zk-dcap-verifier_b200/csrc/transcript.hpp"
// ops: a sequence of records  [kind (1 byte)] [payload]; kind 0 = squeeze (no payload; 32-byte challenge appended to
// `out`), 1 = write_point (64 bytes: x, y canonical little-endian), 2 = write_scalar (32 bytes canonical), 3 = common_scalar.
// Returns the number of bytes written to `out`; the proof bytes are copied to `proof` (proof_len receives their count).
extern "C" size_t transcript_replay(const unsigned char* ops, size_t n_ops_bytes, unsigned char* out, unsigned char* proof,
                                    size_t* proof_len, int kind) {
  TranscriptWrite t(kind);
  size_t pos = 0, w = 0;
  auto canon_to_mont_fr = [](const unsigned char* p) { Fr v; memcpy(v.l, p, 32); return v.to_mont(); };
  auto canon_to_mont_fq = [](const unsigned char* p) { Fq v; memcpy(v.l, p, 32); return v.to_mont(); };
  while (pos < n_ops_bytes) {
    unsigned char kind = ops[pos++];
    if (kind == 0) {
      Fr c = t.squeeze_challenge();
      f_to_repr(c, out + w); w += 32;
    } else if (kind == 1) {
      t.write_point(canon_to_mont_fq(ops + pos), canon_to_mont_fq(ops + pos + 32), false); pos += 64;
    } else if (kind == 2) {
      t.write_scalar(canon_to_mont_fr(ops + pos)); pos += 32;
    } else {
      t.common_scalar(canon_to_mont_fr(ops + pos)); pos += 32;
    }
  }
  memcpy(proof, t.proof.data(), t.proof.size());
  *proof_len = t.proof.size();
  return w;
}
