// transcript.hpp -- host side of the prover: Blake2b-512 and halo2's Blake2bWrite transcript.
//
// Restates halo2_proofs transcript.rs `Blake2bWrite<Vec<u8>, G1Affine, Challenge255<G1Affine>>`
// ([UPSTREAM] @ c254c75f, Cargo.lock:1314-1316; instantiated at circuits/src/sgx_dcap_verifier.rs:813):
// state = Blake2b-512 personalised "Halo2-Transcript"; prefixes 0 = challenge, 1 = point, 2 = scalar;
// a challenge is the state's digest (state cloned) reduced by Fr::from_bytes_wide.
// Blake2b itself follows RFC 7693.  Pure C++ (no CUDA); field arithmetic is ff.cuh's host build.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "ff.cuh"

namespace zk {

struct Blake2b {
  uint64_t h[8];
  uint64_t t0 = 0, t1 = 0;
  uint8_t buf[128];
  size_t buflen = 0;

  static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  static uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian hosts only

  // digest_size 64, no key, 16-byte personalisation (parameter block bytes 48..63)
  explicit Blake2b(const char personal[16]) {
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    uint8_t param[64];
    memset(param, 0, sizeof param);
    param[0] = 64; param[2] = 1; param[3] = 1;       // digest length, fanout, depth
    memcpy(param + 48, personal, 16);
    for (int i = 0; i < 8; i++) h[i] = iv[i] ^ load64(param + 8 * i);
  }
  void compress(const uint8_t* block, bool last) {
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    static const uint8_t sigma[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) m[i] = load64(block + 8 * i);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = iv[i]; }
    v[12] ^= t0; v[13] ^= t1;
    if (last) v[14] = ~v[14];
#define ZK_B2G(a, b, c, d, x, y)                   \
  v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); \
  v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 24);       \
  v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16); \
  v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; r++) {
      const uint8_t* s = sigma[r];
      ZK_B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) ZK_B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
      ZK_B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) ZK_B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
      ZK_B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) ZK_B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
      ZK_B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) ZK_B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef ZK_B2G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
  }
  void update(const uint8_t* p, size_t len) {
    while (len) {
      if (buflen == 128) {                 // the buffered block is not the last one
        t0 += 128; if (t0 < 128) t1++;
        compress(buf, false);
        buflen = 0;
      }
      size_t take = 128 - buflen < len ? 128 - buflen : len;
      memcpy(buf + buflen, p, take);
      buflen += take; p += take; len -= take;
    }
  }
  // digest of a COPY of the state (the transcript keeps absorbing afterwards)
  void digest(uint8_t out[64]) const {
    Blake2b c = *this;
    c.t0 += c.buflen; if (c.t0 < c.buflen) c.t1++;
    memset(c.buf + c.buflen, 0, 128 - c.buflen);
    c.compress(c.buf, true);
    memcpy(out, c.h, 64);
  }
};

// ---- host field helpers -----------------------------------------------------------------------------------
template <class F>
inline F f_from_abi(const void* p) { F v; memcpy(v.l, p, 32); return v; }
template <class F>
inline void f_to_abi(const F& v, void* p) { F r = v.reduce(); memcpy(p, r.l, 32); }
// canonical little-endian bytes (to_repr)
template <class F>
inline void f_to_repr(const F& v, uint8_t out[32]) { F c = v.from_mont(); memcpy(out, c.l, 32); }
// Fr::from_bytes_wide: 512-bit little-endian integer mod r, returned in Montgomery form
inline Fr fr_from_bytes_wide(const uint8_t b[64]) {
  Fr lo, hi;
  memcpy(lo.l, b, 32); memcpy(hi.l, b + 32, 32);
  // lo and hi are arbitrary 256-bit integers: one Montgomery product with R^2 (resp. R^3) both reduces them and
  // moves them to Montgomery form -- lo*R + hi*R*2^256  (values < 2^256 < 6r are fine as REDC inputs: the lazy
  // multiplier only needs a*b < R*4p... so reduce the operands first)
  auto cond_sub = [](Fr v) {       // bring an arbitrary 256-bit value below 2p by subtracting 2p while possible
    for (int i = 0; i < 3; i++) {
      Fr t, m = Fr::kp<2>();
      uint32_t bw = sub8(t.l, v.l, m.l);
      if (!bw) v = t;
    }
    return v;
  };
  lo = cond_sub(lo); hi = cond_sub(hi);
  Fr r2 = Fr::r2();
  Fr r3 = r2 * r2;                 // R^2 * R^2 / R = R^3
  return (lo * r2 + hi * r3).reduce();
}

}  // namespace zk
#include "poseidon.hpp"      // needs fr_from_bytes_wide (above)
namespace zk {

// Keccak-256 (padding 0x01, as Ethereum uses), FIPS 202 permutation
struct Keccak256 {
  static uint64_t rol(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
  static void f1600(uint64_t a[25]) {
    static const uint64_t rc[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
                                    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                                    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
                                    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                                    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
                                    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int rot[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
    for (int r = 0; r < 24; r++) {          // lane (x, y) lives at a[x + 5 y]
      uint64_t c[5], d[5], b[25];
      for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
      for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
      for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
      for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], rot[x][y]);
      for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & b[(x + 2) % 5 + 5 * y]);
      a[0] ^= rc[r];
    }
  }
  static void hash(const uint8_t* data, size_t len, uint8_t out[32]) {
    const size_t rate = 136;
    std::vector<uint8_t> m(data, data + len);
    m.push_back(0x01);
    while (m.size() % rate) m.push_back(0);
    m.back() |= 0x80;
    uint64_t a[25];
    memset(a, 0, sizeof a);
    for (size_t off = 0; off < m.size(); off += rate) {
      for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, &m[off + 8 * i], 8); a[i] ^= w; }
      f1600(a);
    }
    memcpy(out, a, 32);
  }
};

constexpr int kTranscriptBlake2b = 0, kTranscriptEvm = 1, kTranscriptPoseidon = 2;      // == zk_transcript_kind of the C ABI

// kind 0: halo2's Blake2bWrite / Challenge255 (above).  kind 1: snark-verifier's EvmTranscript ([UPSTREAM?]
// system/halo2/transcript/evm.rs; gen_evm_proof_shplonk at crates/p256-ecdsa/src/base.rs:193-199): a byte buffer hashed
// with Keccak-256, points uncompressed (x, y big-endian), scalars big-endian, challenge = keccak(buf ++ [1 if |buf| = 32])
// read big-endian mod r, the digest becomes the new buffer.  kind 2: snark-verifier's PoseidonTranscript over NativeLoader
// (poseidon.hpp; `gen_proof`, crates/p256-ecdsa/src/base.rs:200-211): a point is absorbed as (x mod r, y mod r), a scalar as itself,
// a challenge is the sponge's output; points are written compressed in the halo2curves >= 0.4 encoding that stack pins (y's parity
// in bit 6 of byte 31 -- the encoding of the reference's bin/assets/proof.bin), scalars as canonical little-endian bytes.
struct TranscriptWrite {
  int kind = kTranscriptBlake2b;
  Blake2b state;
  std::vector<uint8_t> evm_buf;
  PoseidonSponge sponge;
  std::vector<uint8_t> proof;
  explicit TranscriptWrite(int kind_ = kTranscriptBlake2b) : kind(kind_), state("Halo2-Transcript") {}
  static void be32(const uint8_t le[32], uint8_t out[32]) { for (int i = 0; i < 32; i++) out[i] = le[31 - i]; }
  static Fr fq_to_fr(const Fq& v) {           // fe_to_fe: the canonical integer mod r
    uint8_t wide[64];
    memset(wide, 0, sizeof wide);
    f_to_repr(v, wide);
    return fr_from_bytes_wide(wide);
  }
  Fr squeeze_challenge() {
    if (kind == kTranscriptPoseidon) return sponge.squeeze();
    if (kind == kTranscriptEvm) {
      std::vector<uint8_t> data = evm_buf;
      if (evm_buf.size() == 32) data.push_back(1);
      uint8_t h[32], wide[64];
      Keccak256::hash(data.data(), data.size(), h);
      evm_buf.assign(h, h + 32);
      memset(wide, 0, sizeof wide);
      be32(h, wide);                          // big-endian digest -> little-endian integer, then mod r
      return fr_from_bytes_wide(wide);
    }
    uint8_t p = 0, d[64];
    state.update(&p, 1);
    state.digest(d);
    return fr_from_bytes_wide(d);
  }
  void common_scalar(const Fr& s) {
    if (kind == kTranscriptPoseidon) { sponge.update(s.reduce()); return; }
    uint8_t b[32];
    f_to_repr(s, b);
    if (kind == kTranscriptEvm) { uint8_t r[32]; be32(b, r); evm_buf.insert(evm_buf.end(), r, r + 32); return; }
    uint8_t p = 2;
    state.update(&p, 1); state.update(b, 32);
  }
  void write_scalar(const Fr& s) {
    common_scalar(s);
    uint8_t b[32]; f_to_repr(s, b);
    if (kind == kTranscriptEvm) { uint8_t r[32]; be32(b, r); proof.insert(proof.end(), r, r + 32); return; }
    proof.insert(proof.end(), b, b + 32);
  }
  // absorb only (the verifier replays what it has read)
  bool common_point(const Fq& x, const Fq& y, bool identity) {
    if (identity) return false;
    if (kind == kTranscriptPoseidon) { sponge.update(fq_to_fr(x)); sponge.update(fq_to_fr(y)); return true; }
    uint8_t bx[32], by[32];
    f_to_repr(x, bx); f_to_repr(y, by);
    if (kind == kTranscriptEvm) {
      uint8_t r[32];
      be32(bx, r); evm_buf.insert(evm_buf.end(), r, r + 32);
      be32(by, r); evm_buf.insert(evm_buf.end(), r, r + 32);
      return true;
    }
    uint8_t p = 1;
    state.update(&p, 1); state.update(bx, 32); state.update(by, 32);
    return true;
  }
  // affine point in Montgomery form; returns false for the identity (upstream: io::Error "cannot write points at
  // infinity to the transcript").  Blake2b: compressed form, x little-endian, y parity in bit 7 of byte 31 [UPSTREAM?].
  bool write_point(const Fq& x, const Fq& y, bool identity) {
    if (!common_point(x, y, identity)) return false;
    uint8_t bx[32], by[32];
    f_to_repr(x, bx); f_to_repr(y, by);
    if (kind == kTranscriptEvm) {
      uint8_t r[32];
      be32(bx, r); proof.insert(proof.end(), r, r + 32);
      be32(by, r); proof.insert(proof.end(), r, r + 32);
      return true;
    }
    bx[31] |= (uint8_t)((by[0] & 1) << (kind == kTranscriptPoseidon ? 6 : 7));
    proof.insert(proof.end(), bx, bx + 32);
    return true;
  }
};

}  // namespace zk
