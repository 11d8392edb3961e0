// poseidon.hpp -- host code: the Poseidon sponge behind snark-verifier's PoseidonTranscript (T = 5, RATE = 4, R_F = 8, R_P = 60 over
// bn256::Fr), the transcript of the reference's non-EVM proofs (crates/p256-ecdsa/src/base.rs:200-211 `gen_proof`; bin/src/main.rs:242
// `PoseidonTranscript::<NativeLoader, &[u8]>::new::<0>(proof)`).  The implementation is in un-vendored dependencies (snark-verifier 0.1.7
// `util/hash/poseidon.rs`, the `poseidon` crate's Grain parameter generation) and is restated from the published construction:
//   * parameters: Grain LFSR of the Poseidon paper -- 80-bit state (field = 1, sbox = 0, 254 field bits, t, R_F, R_P, thirty ones),
//     160 bits discarded, bits emitted in pairs (first bit 1 => output the second); round constants are 254-bit MSB-first integers with
//     rejection sampling, the MDS matrix is the Cauchy matrix 1 / (x_i + y_j) of 2t further elements taken mod r without rejection;
//   * permutation: plain form (constants, x^5 on all / the first element, MDS), R_F/2 + R_P + R_F/2 rounds -- the same function as the
//     optimised form snark-verifier executes;
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
  return S;
}

inline const PoseidonSpec& poseidon_transcript_spec() {
  static const PoseidonSpec S = poseidon_make_spec(PoseidonSpec::T, PoseidonSpec::RF, PoseidonSpec::RP);   // thread-safe (C++11)
  return S;
}

inline void poseidon_permute(const PoseidonSpec& S, Fr* state) {
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
