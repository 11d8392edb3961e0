"""
poseidon.py -- TEST INFRASTRUCTURE: the Poseidon sponge and transcript that the reference's non-EVM proof path uses.

Reference call sites: bin/src/main.rs:239-252 (`PoseidonTranscript::<NativeLoader, &[u8]>::new::<0>(proof)` in
`inner_verify_proof`), crates/p256-ecdsa/src/base.rs:200-211 (`gen_proof`), :227-244 (verification).  The algorithm lives in
dependencies that are NOT vendored in the reference tree: snark-verifier / snark-verifier-sdk 0.1.7 (axiom-crypto @ 4b733e0c,
Cargo.lock:2577-2615; `util/hash/poseidon.rs`, `system/halo2/transcript/halo2.rs`, the SDK's `T = 5, RATE = 4, R_F = 8,
R_P = 60`) and the `poseidon` crate it builds its `Spec` from.  Restated here from the published construction:

* parameters by the Poseidon paper's Grain LFSR (the reference `generate_parameters_grain.sage`): 80-bit state seeded with
  (field = 1, sbox = 0, field bits, t, R_F, R_P, thirty ones), 160 bits discarded, output bits taken in pairs (first bit 1 =>
  emit the second); round constants by rejection sampling of 254-bit MSB-first integers; the MDS matrix is the Cauchy matrix
  1 / (x_i + y_j) of 2t more elements taken WITHOUT rejection (reduced mod r);
* the permutation in its plain form (add constants, x^5 on all / on the first element, multiply by the MDS matrix; R_F / 2 full,
  R_P partial, R_F / 2 full rounds).  snark-verifier runs the "optimized" form (pre-computed constants, sparse matrices), which is
  the same function;
* the sponge of snark-verifier's `Poseidon`: state = (2^64, 0, ...); `squeeze` absorbs the buffered elements RATE at a time, adding 1
  after the last element of a short chunk, runs one more permutation on an empty chunk when the buffer was a multiple of RATE, and
  returns state[1];
* the transcript: a point is absorbed as (x mod r, y mod r), a scalar as itself; points are written compressed, scalars as their
  canonical little-endian bytes; a challenge is the full squeezed field element.

PIN: with t = 3, R_F = 8, R_P = 57 and the circomlib convention (state (0, a, b), output state[0]) this code reproduces the
widely published value poseidon(1, 2) = 0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a, and with t = 5,
R_P = 60 -- the transcript's own parameter set -- poseidon(1, 2, 3, 4) = 0x299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465
(tests/test_poseidon.py; both quoted from memory of circomlib's test-suite, a 254-bit match cannot be accidental).  That pins the
parameter generation and the permutation.  The sponge / transcript framing has NO vector in the reference tree (the golden
`bin/assets/proof.bin` cannot be verified without the un-shipped ./params): that part stays "parity unpinned".

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from . import pymodel as pm

R = pm.R_MOD
P = pm.P_MOD
FIELD_BITS = 254

T, RATE, R_F, R_P = 5, 4, 8, 60        # snark-verifier-sdk's PoseidonTranscript


class Grain:
    def __init__(self, t, r_f, r_p, field_bits=FIELD_BITS):
        bits = []

        def put(n, v):
            for i in reversed(range(n)):
                bits.append((v >> i) & 1)
        put(2, 1); put(4, 0); put(12, field_bits); put(12, t); put(10, r_f); put(10, r_p); put(30, (1 << 30) - 1)
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self.new_bit()

    def new_bit(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self):
        while not self.new_bit():
            self.new_bit()
        return self.new_bit()

    def integer(self, nbits=FIELD_BITS):
        v = 0
        for _ in range(nbits):
            v = (v << 1) | self.bit()
        return v

    def field_element(self):
        while True:
            v = self.integer()
            if v < R:
                return v

    def field_element_without_rejection(self):
        return self.integer() % R


_SPECS = {}


def spec(t=T, r_f=R_F, r_p=R_P):
    """(round constants [(r_f + r_p)][t], mds [t][t])"""
    key = (t, r_f, r_p)
    if key not in _SPECS:
        g = Grain(t, r_f, r_p)
        rc = [[g.field_element() for _ in range(t)] for _ in range(r_f + r_p)]
        xs = [g.field_element_without_rejection() for _ in range(t)]
        ys = [g.field_element_without_rejection() for _ in range(t)]
        mds = [[pow((x + y) % R, R - 2, R) for y in ys] for x in xs]
        _SPECS[key] = (rc, mds)
    return _SPECS[key]


def permute(state, t=T, r_f=R_F, r_p=R_P):
    rc, mds = spec(t, r_f, r_p)
    st = list(state)
    for r in range(r_f + r_p):
        st = [(a + c) % R for a, c in zip(st, rc[r])]
        if r < r_f // 2 or r >= r_f // 2 + r_p:
            st = [pow(a, 5, R) for a in st]
        else:
            st[0] = pow(st[0], 5, R)
        st = [sum(m * a for m, a in zip(row, st)) % R for row in mds]
    return st


class Sponge:
    """snark-verifier `util::hash::Poseidon<F, L, T, RATE>`"""

    def __init__(self):
        self.state = [1 << 64] + [0] * (T - 1)
        self.buf = []

    def update(self, elements):
        self.buf.extend(int(e) % R for e in elements)

    def _absorb(self, chunk):
        st = self.state
        for i, v in enumerate(chunk):
            st[1 + i] = (st[1 + i] + v) % R
        if len(chunk) < RATE:
            st[1 + len(chunk)] = (st[1 + len(chunk)] + 1) % R
        self.state = permute(st)

    def squeeze(self):
        buf, self.buf = self.buf, []
        for i in range(0, len(buf), RATE):
            self._absorb(buf[i:i + RATE])
        if len(buf) % RATE == 0:
            self._absorb([])
        return self.state[1]


def hash_elements(elements):
    s = Sponge()
    s.update(elements)
    return s.squeeze()


# ---- the optimised form (what snark-verifier's `Poseidon::permutation` executes): one full constant vector before the partial rounds,
# a scalar constant per partial round, sparse matrices in the partial rounds.  Derived independently here (see csrc/poseidon.hpp for
# the derivation); tests check that it is the same function as `permute`.
def _mat_inv(m):
    n = len(m)
    a = [list(row) + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(m)]
    for c in range(n):
        piv = next(r for r in range(c, n) if a[r][c] % R)
        a[c], a[piv] = a[piv], a[c]
        inv = pow(a[c][c], R - 2, R)
        a[c] = [v * inv % R for v in a[c]]
        for r in range(n):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(x - f * y) % R for x, y in zip(a[r], a[c])]
    return [row[n:] for row in a]


def _mat_mul(a, b):
    return [[sum(a[i][k] * b[k][j] for k in range(len(b))) % R for j in range(len(b[0]))] for i in range(len(a))]


def _mat_vec(m, v):
    return [sum(x * y for x, y in zip(row, v)) % R for row in m]


_OPT = {}


def optimized_spec(t=T, r_f=R_F, r_p=R_P):
    key = (t, r_f, r_p)
    if key in _OPT:
        return _OPT[key]
    rc, mds = spec(t, r_f, r_p)
    half = r_f // 2
    minv = _mat_inv(mds)
    part = [list(c) for c in rc[half:half + r_p]]
    post = [0] * r_p
    for p in range(r_p - 1, 0, -1):
        d = _mat_vec(minv, part[p])
        for i in range(1, t):
            part[p - 1][i] = (part[p - 1][i] + d[i]) % R
        post[p - 1] = d[0]
        part[p] = [0] * t
    sparse = [None] * r_p                                       # (m00, row[1..], col[1..])
    cur = [list(r) for r in mds]
    for p in range(r_p - 1, -1, -1):
        sub = [row[1:] for row in cur[1:]]
        sub_inv = _mat_inv(sub)
        v = cur[0][1:]
        vrow = [sum(v[k] * sub_inv[k][j] for k in range(t - 1)) % R for j in range(t - 1)]
        sparse[p] = (cur[0][0], vrow, [cur[i][0] for i in range(1, t)])
        dmat = [[1] + [0] * (t - 1)] + [[0] + sub[i] for i in range(t - 1)]
        last_d = dmat
        cur = _mat_mul(dmat, mds)
    pre_sparse = cur
    first = _mat_vec(last_d, part[0])
    _OPT[key] = {"full_first": rc[:half], "full_last": rc[half + r_p:], "pre_sparse": pre_sparse, "first": first, "post": post,
                 "sparse": sparse, "mds": mds}
    return _OPT[key]


def permute_optimized(state, t=T, r_f=R_F, r_p=R_P):
    o = optimized_spec(t, r_f, r_p)
    st = list(state)
    half = r_f // 2
    for r in range(half):
        st = [pow((a + c) % R, 5, R) for a, c in zip(st, o["full_first"][r])]
        st = _mat_vec(o["pre_sparse"] if r == half - 1 else o["mds"], st)
    st = [(a + c) % R for a, c in zip(st, o["first"])]
    for p in range(r_p):
        m00, vrow, col = o["sparse"][p]
        s0 = (pow(st[0], 5, R) + o["post"][p]) % R
        new0 = (m00 * s0 + sum(x * y for x, y in zip(vrow, st[1:]))) % R
        st = [new0] + [(c * s0 + s) % R for c, s in zip(col, st[1:])]
    for r in range(half):
        st = [pow((a + c) % R, 5, R) for a, c in zip(st, o["full_last"][r])]
        st = _mat_vec(o["mds"], st)
    return st
