// poseidon.hpp -- host code: the Poseidon sponge behind snark-verifier's PoseidonTranscript (T = 5, RATE = 4, R_F = 8, R_P = 60 over
// bn256::Fr), the transcript of the reference's non-EVM proofs (crates/p256-ecdsa/src/base.rs:200-211 `gen_proof`; bin/src/main.rs:242
// `PoseidonTranscript::<NativeLoader, &[u8]>::new::<0>(proof)`).  The implementation is in un-vendored dependencies (snark-verifier 0.1.7
// `util/hash/poseidon.rs`, the `poseidon` crate's Grain parameter generation) and is restated from the published construction:
//   * parameters: Grain LFSR of the Poseidon paper -- 80-bit state (field = 1, sbox = 0, 254 field bits, t, R_F, R_P, thirty ones),
//     160 bits discarded, bits emitted in pairs (first bit 1 => output the second); round constants are 254-bit MSB-first integers with
//     rejection sampling, the MDS matrix is the Cauchy matrix 1 / (x_i + y_j) of 2t further elements taken mod r without rejection;
//   * permutation: constants, x^5 on all / the first element, MDS; R_F/2 + R_P + R_F/2 rounds.  Executed in the optimised form (sparse
//     matrices in the partial rounds, derived below); the plain form is kept and zk_poseidon_permute checks one against the other;
//   * sponge (`Poseidon::squeeze`): state (2^64, 0, 0, 0, 0); buffered elements are absorbed RATE at a time into state[1..], a short chunk
//     is followed by +1 in the next position, a buffer that is a multiple of RATE gets one more permutation of an empty chunk; output
//     state[1]; the state carries over to the next squeeze.
// The parameter generation and the permutation are pinned by two published circomlib values (tests/test_poseidon.py, through
// zk_poseidon_permute); the sponge framing has no vector in the reference tree.
#pragma once
#include <vector>

#include "ff.cuh"

namespace zk {

inline Fr fr_from_bytes_wide(const uint8_t b[64]);      // transcript.hpp, which is how this header gets included

// Host arithmetic for the permutation: 4 x 64-bit limbs, same Montgomery representation (R = 2^256) as Fr, always fully reduced.
// (The 8 x 32-bit limb code of ff.cuh is written for the GPU's integer pipe; on the host it costs ~300 us per permutation,
// i.e. ~30 ms per full-chain proof of transcript hashing -- this form takes a sixth of that.)
namespace poseidon_detail {
constexpr uint64_t neg_inv64(uint64_t m0) {     // -m^-1 mod 2^64 (Newton iteration on the low limb)
  uint64_t x = 1;
  for (int i = 0; i < 7; i++) x *= 2 - m0 * x;
  return (uint64_t)(0 - x);
}
}  // namespace poseidon_detail
struct PFr {
  uint64_t l[4];
  static constexpr uint64_t kM[4] = {((uint64_t)FrParams::M1 << 32) | FrParams::M0, ((uint64_t)FrParams::M3 << 32) | FrParams::M2,
                                     ((uint64_t)FrParams::M5 << 32) | FrParams::M4, ((uint64_t)FrParams::M7 << 32) | FrParams::M6};
  static constexpr uint64_t kInv = poseidon_detail::neg_inv64(((uint64_t)FrParams::M1 << 32) | FrParams::M0);
  static_assert((uint32_t)kInv == FrParams::INV, "64-bit Montgomery constant disagrees with ff.cuh");
  static PFr from(const Fr& v) { PFr r; Fr c = v.reduce(); memcpy(r.l, c.l, 32); return r; }
  Fr to_fr() const { Fr r; memcpy(r.l, l, 32); return r; }
  // t (with an extra top word `hi`) -> t - r if t >= r, branch-free
  static void cond_sub(uint64_t* t, uint64_t hi) {
    uint64_t d[4];
    unsigned __int128 bw = 0;
    for (int i = 0; i < 4; i++) { unsigned __int128 x = (unsigned __int128)t[i] - kM[i] - (uint64_t)bw; d[i] = (uint64_t)x; bw = (x >> 64) & 1; }
    const uint64_t take = (uint64_t)0 - (uint64_t)((hi != 0) | (bw == 0));      // all ones when t >= r
    for (int i = 0; i < 4; i++) t[i] = (d[i] & take) | (t[i] & ~take);
  }
  friend PFr operator+(const PFr& a, const PFr& b) {
    PFr r; unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) { c += (unsigned __int128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    cond_sub(r.l, 0);                           // a + b < 2r < 2^255: no carry out of limb 3
    return r;
  }
  friend PFr operator-(const PFr& a, const PFr& b) {          // a + (r - b); b is reduced
    PFr nb; unsigned __int128 bw = 0;
    for (int i = 0; i < 4; i++) { unsigned __int128 x = (unsigned __int128)kM[i] - b.l[i] - (uint64_t)bw; nb.l[i] = (uint64_t)x; bw = (x >> 64) & 1; }
    return a + nb;                                              // r - 0 = r is folded by the conditional subtraction of operator+
  }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  static PFr zero() { PFr r; r.l[0] = r.l[1] = r.l[2] = r.l[3] = 0; return r; }
  static PFr one() { return from(Fr::one()); }
  PFr inverse() const { return from(Fr::inv(to_fr())); }
  friend PFr operator*(const PFr& a, const PFr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      unsigned __int128 c = 0;
      for (int j = 0; j < 4; j++) { c += (unsigned __int128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
      const uint64_t q = t[0] * kInv;
      c = (unsigned __int128)q * kM[0] + t[0]; c >>= 64;
      for (int j = 1; j < 4; j++) { c += (unsigned __int128)q * kM[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    PFr r; memcpy(r.l, t, 32);
    cond_sub(r.l, t[4]);
    return r;
  }
};

struct PoseidonSpec {
  static constexpr int T = 5, RATE = 4, RF = 8, RP = 60;
  std::vector<Fr> rc;        // (RF + RP) * t round constants, Montgomery form
  std::vector<Fr> mds;       // t * t
  std::vector<PFr> rc64, mds64;
  // optimised form (derivation below): the matrix of the last full round before the partial rounds, the one full constant vector
  // added before them, per partial round a scalar added to state[0] after the s-box and a sparse matrix (m00, first row, first column)
  std::vector<PFr> pre_sparse, first, post, sp_m00, sp_row, sp_col;
  int t = T, rf = RF, rp = RP;
};

namespace poseidon_detail {
struct Grain {
  bool s[80];
  int head = 0;              // ring buffer: s[(head + i) % 80] is b_i
  bool new_bit() {
    auto b = [&](int i) { return s[(head + i) % 80]; };
    bool n = b(62) ^ b(51) ^ b(38) ^ b(23) ^ b(13) ^ b(0);
    s[head] = n;             // b_0 leaves, the new bit becomes b_79
    head = (head + 1) % 80;
    return n;
  }
  bool bit() {
    while (!new_bit()) new_bit();
    return new_bit();
  }
  Grain(uint32_t t, uint32_t rf, uint32_t rp) {
    int n = 0;
    auto put = [&](int bits, uint32_t v) { for (int i = bits - 1; i >= 0; i--) s[n++] = (v >> i) & 1u; };
    put(2, 1); put(4, 0); put(12, 254); put(12, t); put(10, rf); put(10, rp); put(30, 0x3fffffffu);
    for (int i = 0; i < 160; i++) new_bit();
  }
  // 254 bits, most significant first, as a little-endian 256-bit integer
  void integer(uint32_t out[8]) {
    for (int i = 0; i < 8; i++) out[i] = 0;
    for (int i = 253; i >= 0; i--)
      if (bit()) out[i / 32] |= 1u << (i % 32);
  }
  static Fr to_field(const uint32_t v[8]) {          // any 256-bit integer -> v mod r, Montgomery form
    uint8_t wide[64];
    memset(wide, 0, sizeof wide);
    memcpy(wide, v, 32);
    return fr_from_bytes_wide(wide);
  }
  Fr field_element() {
    for (;;) {
      uint32_t v[8], t_[8];
      integer(v);
      Fr m = Fr::kp<1>();
      if (sub8(t_, v, m.l)) return to_field(v);       // borrow: v < r
    }
  }
  Fr field_element_without_rejection() {
    uint32_t v[8];
    integer(v);
    return to_field(v);
  }
};
}  // namespace poseidon_detail

// The optimised form of the permutation (the same function; what snark-verifier executes, derived here independently).
// Partial round p is  s += c_p;  s[0] = s[0]^5;  s = M s.
//  (1) constants: c_p = M (M^-1 c_p), and M^-1 c_p =: d can be added BEFORE the multiplication by M of round p-1; its components
//      1.. commute with the s-box of round p-1 (which touches s[0] only) and join c_(p-1), its component 0 stays behind the s-box as a
//      scalar.  Backwards over the partial rounds: one full vector before the first partial round, one scalar per round.
//  (2) matrices: M = S D with D = diag(1, M^) (M^ = M without row and column 0) and S = [[m00, v^T M^^-1], [w, I]] sparse.  D commutes
//      with everything a partial round does to s[0], so it moves to the previous round's matrix: (D M) is factored again, and so on;
//      what is left after the first partial round's matrix joins the matrix of the preceding full round (`pre_sparse` = D_0 M, and
//      the full vector becomes D_0 c).  A partial round then costs 2t - 1 multiplications instead of t^2.
namespace poseidon_detail {
inline std::vector<PFr> mat_inverse(std::vector<PFr> a, int n) {          // Gauss-Jordan; a is n x n row-major and invertible
  std::vector<PFr> inv((size_t)n * n, PFr::zero());
  for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = PFr::one();
  for (int c = 0; c < n; c++) {
    int piv = c;
    while (piv < n && a[(size_t)piv * n + c].is_zero()) piv++;
    if (piv == n) return {};
    for (int j = 0; j < n; j++) { std::swap(a[(size_t)c * n + j], a[(size_t)piv * n + j]); std::swap(inv[(size_t)c * n + j], inv[(size_t)piv * n + j]); }
    const PFr f = a[(size_t)c * n + c].inverse();
    for (int j = 0; j < n; j++) { a[(size_t)c * n + j] = a[(size_t)c * n + j] * f; inv[(size_t)c * n + j] = inv[(size_t)c * n + j] * f; }
    for (int r = 0; r < n; r++) {
      if (r == c || a[(size_t)r * n + c].is_zero()) continue;
      const PFr g = a[(size_t)r * n + c];
      for (int j = 0; j < n; j++) {
        a[(size_t)r * n + j] = a[(size_t)r * n + j] - g * a[(size_t)c * n + j];
        inv[(size_t)r * n + j] = inv[(size_t)r * n + j] - g * inv[(size_t)c * n + j];
      }
    }
  }
  return inv;
}
inline void optimise(PoseidonSpec& S) {
  const int t = S.t, half = S.rf / 2, rp = S.rp, m = t - 1;
  const std::vector<PFr>& M = S.mds64;
  const std::vector<PFr> Minv = mat_inverse(M, t);
  std::vector<PFr> part(S.rc64.begin() + (size_t)half * t, S.rc64.begin() + (size_t)(half + rp) * t);
  S.post.assign(rp, PFr::zero());
  for (int p = rp - 1; p >= 1; p--) {
    std::vector<PFr> d(t, PFr::zero());
    for (int i = 0; i < t; i++)
      for (int j = 0; j < t; j++) d[i] = d[i] + Minv[(size_t)i * t + j] * part[(size_t)p * t + j];
    for (int i = 1; i < t; i++) part[(size_t)(p - 1) * t + i] = part[(size_t)(p - 1) * t + i] + d[i];
    S.post[p - 1] = d[0];
  }
  S.sp_m00.assign(rp, PFr::zero()); S.sp_row.assign((size_t)rp * m, PFr::zero()); S.sp_col.assign((size_t)rp * m, PFr::zero());
  std::vector<PFr> cur = M, D((size_t)t * t, PFr::zero());
  for (int p = rp - 1; p >= 0; p--) {
    std::vector<PFr> sub((size_t)m * m);
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) sub[(size_t)i * m + j] = cur[(size_t)(i + 1) * t + j + 1];
    const std::vector<PFr> sinv = mat_inverse(sub, m);
    S.sp_m00[p] = cur[0];
    for (int j = 0; j < m; j++) {
      PFr acc = PFr::zero();
      for (int k = 0; k < m; k++) acc = acc + cur[k + 1] * sinv[(size_t)k * m + j];
      S.sp_row[(size_t)p * m + j] = acc;
      S.sp_col[(size_t)p * m + j] = cur[(size_t)(j + 1) * t];
    }
    std::fill(D.begin(), D.end(), PFr::zero());
    D[0] = PFr::one();
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) D[(size_t)(i + 1) * t + j + 1] = sub[(size_t)i * m + j];
    std::vector<PFr> next((size_t)t * t, PFr::zero());
    for (int i = 0; i < t; i++)
      for (int j = 0; j < t; j++)
        for (int k = 0; k < t; k++) next[(size_t)i * t + j] = next[(size_t)i * t + j] + D[(size_t)i * t + k] * M[(size_t)k * t + j];
    cur.swap(next);
  }
  S.pre_sparse = cur;                                   // D_0 M
  S.first.assign(t, PFr::zero());                       // D_0 c
  for (int i = 0; i < t; i++)
    for (int j = 0; j < t; j++) S.first[i] = S.first[i] + D[(size_t)i * t + j] * part[j];
}
}  // namespace poseidon_detail

inline PoseidonSpec poseidon_make_spec(int t, int rf, int rp) {
  PoseidonSpec S;
  S.t = t; S.rf = rf; S.rp = rp;
  poseidon_detail::Grain g((uint32_t)t, (uint32_t)rf, (uint32_t)rp);
  S.rc.resize((size_t)(rf + rp) * t);
  for (auto& c : S.rc) c = g.field_element();
  std::vector<Fr> xs(t), ys(t);
  for (auto& x : xs) x = g.field_element_without_rejection();
  for (auto& y : ys) y = g.field_element_without_rejection();
  S.mds.resize((size_t)t * t);
  for (int i = 0; i < t; i++)
    for (int j = 0; j < t; j++) S.mds[(size_t)i * t + j] = Fr::inv(xs[i] + ys[j]).reduce();
  for (auto& c : S.rc) S.rc64.push_back(PFr::from(c));
  for (auto& c : S.mds) S.mds64.push_back(PFr::from(c));
  poseidon_detail::optimise(S);
  return S;
}

inline const PoseidonSpec& poseidon_transcript_spec() {
  static const PoseidonSpec S = poseidon_make_spec(PoseidonSpec::T, PoseidonSpec::RF, PoseidonSpec::RP);   // thread-safe (C++11)
  return S;
}

inline void poseidon_permute_plain(const PoseidonSpec& S, Fr* state) {
  const int t = S.t, half = S.rf / 2;
  PFr st[16], nx[16];
  for (int i = 0; i < t; i++) st[i] = PFr::from(state[i]);
  for (int r = 0; r < S.rf + S.rp; r++) {
    const PFr* rc = &S.rc64[(size_t)r * t];
    for (int i = 0; i < t; i++) st[i] = st[i] + rc[i];
    const int nbox = (r < half || r >= half + S.rp) ? t : 1;
    for (int i = 0; i < nbox; i++) {
      PFr x2 = st[i] * st[i];
      st[i] = x2 * x2 * st[i];
    }
    for (int i = 0; i < t; i++) {
      const PFr* row = &S.mds64[(size_t)i * t];
      PFr acc = row[0] * st[0];
      for (int j = 1; j < t; j++) acc = acc + row[j] * st[j];
      nx[i] = acc;
    }
    for (int i = 0; i < t; i++) st[i] = nx[i];
  }
  for (int i = 0; i < t; i++) state[i] = st[i].to_fr();
}

inline void poseidon_permute(const PoseidonSpec& S, Fr* state) {
  const int t = S.t, half = S.rf / 2, m = t - 1;
  PFr st[16], nx[16];
  for (int i = 0; i < t; i++) st[i] = PFr::from(state[i]);
  auto full_round = [&](const PFr* rc, const PFr* mat) {
    for (int i = 0; i < t; i++) {
      PFr x = st[i] + rc[i], x2 = x * x;
      st[i] = x2 * x2 * x;
    }
    for (int i = 0; i < t; i++) {
      const PFr* row = mat + (size_t)i * t;
      PFr acc = row[0] * st[0];
      for (int j = 1; j < t; j++) acc = acc + row[j] * st[j];
      nx[i] = acc;
    }
    for (int i = 0; i < t; i++) st[i] = nx[i];
  };
  for (int r = 0; r < half; r++) full_round(&S.rc64[(size_t)r * t], r == half - 1 ? S.pre_sparse.data() : S.mds64.data());
  for (int i = 0; i < t; i++) st[i] = st[i] + S.first[i];
  for (int p = 0; p < S.rp; p++) {
    PFr x2 = st[0] * st[0];
    const PFr s0 = x2 * x2 * st[0] + S.post[p];
    const PFr* row = &S.sp_row[(size_t)p * m];
    const PFr* col = &S.sp_col[(size_t)p * m];
    PFr acc = S.sp_m00[p] * s0;
    for (int j = 0; j < m; j++) acc = acc + row[j] * st[j + 1];
    for (int j = 0; j < m; j++) st[j + 1] = st[j + 1] + col[j] * s0;
    st[0] = acc;
  }
  for (int r = 0; r < half; r++) full_round(&S.rc64[(size_t)(half + S.rp + r) * t], S.mds64.data());
  for (int i = 0; i < t; i++) state[i] = st[i].to_fr();
}

struct PoseidonSponge {
  static constexpr int T = PoseidonSpec::T, RATE = PoseidonSpec::RATE;
  Fr state[T];
  std::vector<Fr> buf;
  PoseidonSponge() {
    uint8_t wide[64];
    memset(wide, 0, sizeof wide);
    wide[8] = 1;                                      // 2^64
    state[0] = fr_from_bytes_wide(wide);
    for (int i = 1; i < T; i++) state[i] = Fr::zero();
  }
  void update(const Fr& v) { buf.push_back(v); }
  void absorb(const Fr* chunk, size_t len) {
    for (size_t i = 0; i < len; i++) state[1 + i] = state[1 + i] + chunk[i];
    if (len < (size_t)RATE) state[1 + len] = state[1 + len] + Fr::one();
    poseidon_permute(poseidon_transcript_spec(), state);
  }
  Fr squeeze() {
    std::vector<Fr> b;
    b.swap(buf);
    for (size_t i = 0; i < b.size(); i += RATE) absorb(b.data() + i, std::min((size_t)RATE, b.size() - i));
    if (b.size() % RATE == 0) absorb(nullptr, 0);
    return state[1];
  }
};

}  // namespace zk
