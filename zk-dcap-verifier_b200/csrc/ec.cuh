// ec.cuh -- BN254 G1 (y^2 = x^3 + 3 over Fq) group law for sm_100a.
//
// Replaces halo2curves 0.3.1 bn256::{G1Affine, G1} arithmetic (src/bn256/curve.rs,
// src/derive/curve.rs; Cargo.lock:1329-1331) inside the MSM kernels.  Memory formats at the
// C ABI are halo2curves': G1Affine = {x, y} (identity = (0,0)), G1 = Jacobian {x, y, z}
// (identity z = 0).  Internally buckets are accumulated in extended Jacobian "XYZZ"
// coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), whose mixed addition needs 8M + 2S and no
// inversion; every group result is normalised before it leaves the device, so outputs are
// canonical and bit-identical to the CPU prover's.
#pragma once
#include "ff.cuh"

namespace zk {

struct Affine {
  Fq x, y;
  ZK_HD static Affine load(const void* p) {
    Affine a;
    a.x = Fq::load(p);
    a.y = Fq::load(reinterpret_cast<const char*>(p) + 32);
    return a;
  }
  ZK_HD void store(void* p) const {
    x.store(p);
    y.store(reinterpret_cast<char*>(p) + 32);
  }
  ZK_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
  ZK_HD static Affine identity() { Affine a; a.x = Fq::zero(); a.y = Fq::zero(); return a; }
  ZK_HD Affine neg() const { Affine a; a.x = x; a.y = y.neg(); return a; }
  // conditional negation without divergence on the value
  ZK_HD Affine cneg(bool s) const {
    Affine a; a.x = x;
    Fq ny = y.neg();
#pragma unroll
    for (int i = 0; i < 8; i++) a.y.l[i] = s ? ny.l[i] : y.l[i];
    return a;
  }
};

struct XYZZ {
  Fq x, y, zz, zzz;
  ZK_HD static XYZZ identity() {
    XYZZ p; p.x = Fq::zero(); p.y = Fq::zero(); p.zz = Fq::zero(); p.zzz = Fq::zero();
    return p;
  }
  // zz of the identity is the literal 0 (identity(), zeroed bucket memory, 0 * x in the formulas); zz of any other point is
  // a product of non-zero field elements, whose lazy representative in [0, 2p) is never 0 -- so a raw limb test suffices
  // (the full test needs a conditional subtraction of p first: ~40 more instructions in every mixed addition)
  ZK_HD bool is_identity() const { return zz.is_zero_raw(); }
  ZK_HD static XYZZ from_affine(const Affine& a) {
    if (a.is_identity()) return identity();
    XYZZ p; p.x = a.x; p.y = a.y; p.zz = Fq::one(); p.zzz = Fq::one();
    return p;
  }
  ZK_HD static XYZZ load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    XYZZ r; r.x = Fq::load(c); r.y = Fq::load(c + 32); r.zz = Fq::load(c + 64); r.zzz = Fq::load(c + 96);
    return r;
  }
  ZK_HD void store(void* p) const {
    char* c = reinterpret_cast<char*>(p);
    x.store_raw(c); y.store_raw(c + 32); zz.store_raw(c + 64); zzz.store_raw(c + 96);
  }
  ZK_HD XYZZ neg() const { XYZZ p = *this; p.y = y.neg(); return p; }

  // 2 * (affine point), mdbl-2008-s-1 with a = 0
  ZK_HD static XYZZ dbl_affine(const Affine& a) {
    Fq u = a.y.dbl(), v = u.sqr(), w = u * v, s = a.x * v;
    Fq xx = a.x.sqr(), m = xx.dbl() + xx;
    XYZZ r;
    r.x = m.sqr() - s.dbl();
    r.y = m * (s - r.x) - w * a.y;
    r.zz = v; r.zzz = w;
    return r;
  }
  // 2 * P, dbl-2008-s-1 with a = 0
  ZK_HD XYZZ dbl() const {
    if (is_identity()) return *this;
    Fq u = y.dbl(), v = u.sqr(), w = u * v, s = x * v;
    Fq xx = x.sqr(), m = xx.dbl() + xx;
    XYZZ r;
    r.x = m.sqr() - s.dbl();
    r.y = m * (s - r.x) - w * y;
    r.zz = v * zz; r.zzz = w * zzz;
    return r;
  }
  // P + (affine Q), madd-2008-s; all exceptional cases handled
  // `q` comes from memory (canonical limbs) or from cneg of such a point (0 - 0 = 0), so its identity test is raw too
  ZK_HD XYZZ madd(const Affine& q) const {
    if (q.x.is_zero_raw() && q.y.is_zero_raw()) return *this;
    if (is_identity()) { XYZZ r; r.x = q.x; r.y = q.y; r.zz = Fq::one(); r.zzz = Fq::one(); return r; }
    Fq u2 = q.x * zz, s2 = q.y * zzz;
    Fq p = u2 - x, r_ = s2 - y;
    if (p.is_zero()) {
      if (r_.is_zero()) return dbl_affine(q);
      return identity();
    }
    Fq pp = p.sqr(), ppp = p * pp, qv = x * pp;
    XYZZ r;
    r.x = r_.sqr() - ppp - qv.dbl();
    r.y = r_ * (qv - r.x) - y * ppp;
    r.zz = zz * pp; r.zzz = zzz * ppp;
    return r;
  }
  // P + Q, add-2008-s
  ZK_HD XYZZ add(const XYZZ& q) const {
    if (q.is_identity()) return *this;
    if (is_identity()) return q;
    Fq u1 = x * q.zz, u2 = q.x * zz, s1 = y * q.zzz, s2 = q.y * zzz;
    Fq p = u2 - u1, r_ = s2 - s1;
    if (p.is_zero()) {
      if (r_.is_zero()) return dbl();
      return identity();
    }
    Fq pp = p.sqr(), ppp = p * pp, qv = u1 * pp;
    XYZZ r;
    r.x = r_.sqr() - ppp - qv.dbl();
    r.y = r_ * (qv - r.x) - s1 * ppp;
    r.zz = zz * q.zz * pp; r.zzz = zzz * q.zzz * ppp;
    return r;
  }
  // normalise (one field inversion)
  ZK_HD_NOINL Affine to_affine() const {
    if (is_identity()) return Affine::identity();
    Fq t = Fq::inv(zz * zzz);
    Affine a;
    a.x = x * (t * zzz);     // X / ZZ
    a.y = y * (t * zz);      // Y / ZZZ
    return a;
  }
};

// write a normalised halo2curves G1: (x, y, 1), or the identity (0, 1, 0)
ZK_HD void store_g1_normalised(void* out, const Affine& a) {
  char* c = reinterpret_cast<char*>(out);
  if (a.is_identity()) {
    Fq::zero().store(c); Fq::one().store(c + 32); Fq::zero().store(c + 64);
  } else {
    a.x.store(c); a.y.store(c + 32); Fq::one().store(c + 64);
  }
}
// read a halo2curves G1 (Jacobian x, y, z: affine = (x/z^2, y/z^3)) as XYZZ
ZK_HD XYZZ load_g1_as_xyzz(const void* in) {
  const char* c = reinterpret_cast<const char*>(in);
  Fq x = Fq::load(c), y = Fq::load(c + 32), z = Fq::load(c + 64);
  if (z.is_zero()) return XYZZ::identity();
  XYZZ p; p.x = x; p.y = y; p.zz = z.sqr(); p.zzz = p.zz * z;
  return p;
}

// ---- signed-digit recoding of a canonical 254-bit scalar -----------------------------------
// Window w of width c yields a digit in [-2^(c-1)+1, 2^(c-1)].  `limbs` is the canonical
// little-endian value (8 x u32).  Returns digit for window w given the carry from window w-1
// and updates the carry.  W = 254/c + 1 windows never leave a carry.
ZK_HD int signed_digit(const uint32_t* limbs, int w, int c, int& carry) {
  int bit = w * c;
  uint32_t v = 0;
  if (bit < 256) {
    int li = bit >> 5, sh = bit & 31;
    uint64_t two = limbs[li];
    if (li + 1 < 8) two |= (uint64_t)limbs[li + 1] << 32;
    v = (uint32_t)((two >> sh) & ((1u << c) - 1u));
  }
  int d = (int)v + carry;
  if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
  return d;
}


// The same recoding as a STREAM: windows are taken from the bottom while the value is shifted right by c bits, so every
// limb access has a compile-time index and the scalar stays in registers (signed_digit's `limbs[bit >> 5]` with a
// run-time window number made the digit kernels read the scalar back from local memory: 28 LDL per thread).
// 1 <= c <= 31; `l` is consumed.
ZK_HD int next_signed_digit(uint32_t* l, int c, int& carry) {
  const uint32_t v = l[0] & ((1u << c) - 1u);
#pragma unroll
  for (int i = 0; i < 7; i++) l[i] = (l[i] >> c) | (l[i + 1] << (32 - c));
  l[7] >>= c;
  int d = (int)v + carry;
  if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
  return d;
}

}  // namespace zk
