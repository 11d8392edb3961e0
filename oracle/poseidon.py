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
