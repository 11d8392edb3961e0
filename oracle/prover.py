"""
prover.py -- TEST INFRASTRUCTURE: CPU restatement of halo2's KZG/SHPLONK prover and key generation.

PARITY UNPINNED (see oracle/zk_oracle.c): restated from halo2_proofs 0.2.0 @ zkwebauthn/halo2 c254c75f
(Cargo.lock:1314-1316) -- plonk/prover.rs `create_proof`, plonk/keygen.rs `keygen_vk`/`keygen_pk`,
plonk/{lookup,permutation,vanishing}/prover.rs, poly/kzg/multiopen/shplonk/prover.rs, transcript.rs
(Blake2bWrite / Challenge255) -- whose sources are not vendored in the reference tree.  Reference call sites:
circuits/src/sgx_dcap_verifier.rs:799 (gen_srs), :803 (keygen_vk), :807 (keygen_pk), :813 (Blake2bWrite::init),
:814-822 (create_proof), :823 (finalize).  What pins it: oracle/verifier.py accepts its proofs (pairing check with
the real G2 elements), and every commitment is a best_multiexp / every transform a best_fft of the pinned oracle.

Every n-sized loop runs in oracle/libzkoracle.so (zk_oracle.c, prover_tail.inc.h); this file holds the
order-sensitive orchestration: RNG draw order, transcript order (3.2, A.8), query order and the
SHPLONK set construction (A.5).  The circuit arrives as plain data: the `ConstraintSystem` object (expression trees,
query lists) plus the ctypes graph images the host layer builds from it.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module.
"""
import hashlib

import numpy as np

from . import pymodel as pm

R = pm.R_MOD
P = pm.P_MOD


def L(v):
    """canonical int -> Montgomery limbs (4,)"""
    return pm.to_mont_limbs([v], R)[0]


def I(limbs):
    """Montgomery limbs -> canonical int"""
    return pm.from_mont_limbs(np.asarray(limbs).reshape(1, 4), R)[0]


# ---- transcript.rs: Blake2bWrite / Blake2bRead with Challenge255 -------------------------------------
def compress_point(pt):
    """G1Affine::to_bytes of halo2curves 0.3.1 [UPSTREAM?]: x little-endian, y's parity in bit 7 of byte 31;
    the identity is all zeros."""
    if pt is None:
        return bytes(32)
    b = bytearray(pt[0].to_bytes(32, "little"))
    b[31] |= (pt[1] & 1) << 7
    return bytes(b)


def decompress_point(b):
    if b == bytes(32):
        return None
    sign = b[31] >> 7
    x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x7F]), "little")
    if x >= P:
        raise ValueError("invalid point encoding")
    y = pow((x * x * x + 3) % P, (P + 1) // 4, P)
    if (y * y - x * x * x - 3) % P:
        raise ValueError("not on curve")
    if (y & 1) != sign:
        y = P - y
    return (x, y)


def compress_point_v04(pt):
    """G1Affine::to_bytes of halo2curves >= 0.4 / halo2curves-axiom [UPSTREAM?] (the halo2-axiom stack, Cargo.lock:1181-1199):
    parity of y in bit 6 of byte 31, identity = flag in bit 7.  The sign position is pinned by the flagged commitments of
    the reference's bin/assets/proof.bin (tests/golden/ref_proof_points.json)."""
    if pt is None:
        return bytes(31) + b"\x80"
    b = bytearray(pt[0].to_bytes(32, "little"))
    b[31] |= (pt[1] & 1) << 6
    return bytes(b)


def decompress_point_v04(b):
    if b[31] & 0x80:
        if bytes(b[:31]) != bytes(31) or b[31] != 0x80:
            raise ValueError("invalid point encoding")
        return None
    sign = (b[31] >> 6) & 1
    x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x3F]), "little")
    if x >= P:
        raise ValueError("invalid point encoding")
    y = pow((x * x * x + 3) % P, (P + 1) // 4, P)
    if (y * y - x * x * x - 3) % P:
        raise ValueError("not on curve")
    if (y & 1) != sign:
        y = P - y
    return (x, y)


class Blake2bWrite:
    """Blake2bWrite<Vec<u8>, G1Affine, Challenge255<G1Affine>> (sgx_dcap_verifier.rs:813)"""

    def __init__(self):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.proof = bytearray()

    def squeeze_challenge(self):
        self.state.update(b"\x00")
        return int.from_bytes(self.state.copy().digest(), "little") % R      # Fr::from_bytes_wide

    def common_point(self, pt):
        if pt is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.state.update(b"\x01" + pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little"))

    def common_scalar(self, s):
        self.state.update(b"\x02" + s.to_bytes(32, "little"))

    def write_point(self, pt):
        self.common_point(pt)
        self.proof += compress_point(pt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += s.to_bytes(32, "little")

    def finalize(self):
        return bytes(self.proof)


class Blake2bRead(Blake2bWrite):
    def __init__(self, proof):
        super().__init__()
        self.buf, self.pos = bytes(proof), 0

    def read_point(self):
        b = self.buf[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof too short")
        self.pos += 32
        pt = decompress_point(b)
        self.common_point(pt)
        return pt

    def read_scalar(self):
        b = self.buf[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof too short")
        self.pos += 32
        s = int.from_bytes(b, "little")
        if s >= R:
            raise ValueError("invalid scalar encoding")
        self.common_scalar(s)
        return s


class EvmWrite:
    """snark-verifier `EvmTranscript<G1Affine, NativeLoader, _, Vec<u8>>` [UPSTREAM? snark-verifier
    system/halo2/transcript/evm.rs, the transcript of gen_evm_proof_shplonk at crates/p256-ecdsa/src/base.rs:193-199]:
    a byte buffer hashed with Keccak-256; points are absorbed / written uncompressed as big-endian x, y, scalars as
    32 big-endian bytes; a challenge is keccak(buf ++ [1 if len(buf) == 32]) read big-endian mod r, and the digest
    becomes the new buffer."""

    def __init__(self):
        from .keccak import keccak256
        self._h = keccak256
        self.buf = bytearray()
        self.proof = bytearray()

    def squeeze_challenge(self):
        data = bytes(self.buf) + (b"\x01" if len(self.buf) == 32 else b"")
        digest = self._h(data)
        self.buf = bytearray(digest)
        return int.from_bytes(digest, "big") % R

    def common_point(self, pt):
        if pt is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.buf += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    def common_scalar(self, s):
        self.buf += s.to_bytes(32, "big")

    def write_point(self, pt):
        self.common_point(pt)
        self.proof += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += s.to_bytes(32, "big")

    def finalize(self):
        return bytes(self.proof)


class EvmRead(EvmWrite):
    def __init__(self, proof):
        super().__init__()
        self.rbuf, self.pos = bytes(proof), 0

    def read_point(self):
        b = self.rbuf[self.pos:self.pos + 64]
        if len(b) != 64:
            raise ValueError("proof too short")
        self.pos += 64
        x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big")
        if x >= P or y >= P or (y * y - x * x * x - 3) % P:
            raise ValueError("invalid point")
        self.common_point((x, y))
        return (x, y)

    def read_scalar(self):
        b = self.rbuf[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof too short")
        self.pos += 32
        s = int.from_bytes(b, "big")
        if s >= R:
            raise ValueError("invalid scalar encoding")
        self.common_scalar(s)
        return s

    @property
    def buf_len(self):
        return len(self.rbuf)


class PoseidonWrite:
    """snark-verifier `PoseidonTranscript<G1Affine, NativeLoader, Vec<u8>, 5, 4, 8, 60>` (oracle/poseidon.py; the transcript of
    `gen_proof`, crates/p256-ecdsa/src/base.rs:200-211, and of the CLI's verifier, bin/src/main.rs:242): a point is absorbed as
    (x mod r, y mod r), a scalar as itself, a challenge is the sponge's output element.  Points are written compressed in the
    encoding of the halo2curves version that stack pins (y's parity in bit 6: `compress_point_v04`, the encoding of the golden
    bin/assets/proof.bin), scalars as canonical little-endian bytes."""

    def __init__(self):
        from .poseidon import Sponge
        self.sponge = Sponge()
        self.proof = bytearray()

    def squeeze_challenge(self):
        return self.sponge.squeeze()

    def common_point(self, pt):
        if pt is None:
            raise ValueError("Invalid elliptic curve point encoding in proof")
        self.sponge.update([pt[0] % R, pt[1] % R])

    def common_scalar(self, s):
        self.sponge.update([s])

    def write_point(self, pt):
        self.common_point(pt)
        self.proof += compress_point_v04(pt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += s.to_bytes(32, "little")

    def finalize(self):
        return bytes(self.proof)


class PoseidonRead(PoseidonWrite):
    def __init__(self, proof):
        super().__init__()
        self.rbuf, self.pos = bytes(proof), 0

    def read_point(self):
        b = self.rbuf[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof too short")
        self.pos += 32
        pt = decompress_point_v04(b)
        self.common_point(pt)
        return pt

    def read_scalar(self):
        b = self.rbuf[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof too short")
        self.pos += 32
        s = int.from_bytes(b, "little")
        if s >= R:
            raise ValueError("invalid scalar encoding")
        self.common_scalar(s)
        return s

    @property
    def buf_len(self):
        return len(self.rbuf)


class Randoms:
    """The caller's RNG, pre-drawn: element i is the i-th `Fr::random(&mut rng)` create_proof would make."""

    def __init__(self, arr):
        self.arr, self.pos = np.asarray(arr, dtype=np.uint64).reshape(-1, 4), 0

    def take(self, count):
        if self.pos + count > self.arr.shape[0]:
            raise ValueError("random stream exhausted")
        out = self.arr[self.pos:self.pos + count]
        self.pos += count
        return out


def randoms_needed(cs, k):
    """Number of Fr::random draws of one create_proof."""
    n, bf, d = 1 << k, cs.blinding_factors(), cs.degree()
    chunk = d - 2
    sets = (len(cs.permutation_columns) + chunk - 1) // chunk if cs.permutation_columns else 0
    a, l = cs.num_advice_columns, len(cs.lookups)
    return a * (bf + 1) + a + l * (2 * (bf + 1) + 2) + sets * (bf + 1) + l * (bf + 1) + n + 1 + (d - 1)


# ---- ParamsKZG::setup with a known trapdoor (poly/kzg/commitment.rs) -----------------------------------
def setup_params(orc, k, s):
    """g[i] = s^i * G, g_lagrange[i] = L_i(s) * G (closed form), G = (1, 2)."""
    n = 1 << k
    gen = pm.affine_to_limbs([(1, 2)])[0]
    w = pm.omega_for(k)
    g = np.empty((n, 8), dtype=np.uint64)
    gl = np.empty((n, 8), dtype=np.uint64)
    num = (pow(s, n, R) - 1) * pow(n, -1, R) % R
    sp, wp = 1, 1
    for i in range(n):
        g[i] = orc.g1_to_affine(orc.g1_mul(gen, L(sp)))[0]
        li = num * wp % R * pow((s - wp) % R, -1, R) % R
        gl[i] = orc.g1_to_affine(orc.g1_mul(gen, L(li)))[0]
        sp = sp * s % R
        wp = wp * w % R
    return {"k": k, "n": n, "g": g, "g_lagrange": gl, "s": s}


def rotate_omega(k, x, rot):
    w = pm.omega_for(k)
    return x * pow(w, rot, R) % R          # negative exponents invert (omega_inv)


# ---- keygen_vk / keygen_pk (plonk/keygen.rs) -----------------------------------------------------------
def keygen(orc, params, circ, fixed_values, mapping, transcript_repr, threads=None):
    """circ: host-layer Circuit (cs, k, ctypes images).  fixed_values: Lagrange columns AFTER selector compression.
    mapping: (n_perm_columns, n, 2) uint32 -- cell (i, j) maps to (column', row') in the permutation cycles."""
    cs, k = circ.cs, circ.k
    n, bf = 1 << k, cs.blinding_factors()
    dom = orc.Domain(cs.degree(), k, threads)
    commit_l = lambda v: pm.limbs_to_affine(orc.g1_to_affine(orc.best_multiexp(v, params["g_lagrange"], threads)))[0]
    pk = {"domain": dom, "transcript_repr": transcript_repr, "k": k}
    pk["fixed_values"] = [np.ascontiguousarray(v, dtype=np.uint64) for v in fixed_values]
    pk["fixed_polys"] = [dom.lagrange_to_coeff(v) for v in pk["fixed_values"]]
    pk["fixed_cosets"] = [dom.coeff_to_extended(p) for p in pk["fixed_polys"]]
    npc = len(cs.permutation_columns)
    sig = orc.permutation_sigma(k, mapping, dom.omega) if npc else np.zeros((0, n, 4), np.uint64)
    pk["sigma_values"] = [sig[i] for i in range(npc)]
    pk["sigma_polys"] = [dom.lagrange_to_coeff(v) for v in pk["sigma_values"]]
    pk["sigma_cosets"] = [dom.coeff_to_extended(p) for p in pk["sigma_polys"]]
    one, zero = L(1), L(0)
    l0 = np.tile(zero, (n, 1)); l0[0] = one
    l_blind = np.tile(zero, (n, 1)); l_blind[n - bf:] = one
    l_last = np.tile(zero, (n, 1)); l_last[n - bf - 1] = one
    ext = lambda v: dom.coeff_to_extended(dom.lagrange_to_coeff(v))
    pk["l0"], pk["l_last"] = ext(l0), ext(l_last)
    lb = ext(l_blind)
    ones = np.tile(one, (dom.ext_n, 1))
    pk["l_active_row"] = orc.field_op("fr", "sub", ones, orc.field_op("fr", "add", pk["l_last"], lb))
    pk["vk"] = {"fixed_commitments": [commit_l(v) for v in pk["fixed_values"]],
                "permutation_commitments": [commit_l(v) for v in pk["sigma_values"]],
                "transcript_repr": transcript_repr}
    return pk


# ---- small host-side polynomial helpers (canonical ints) --------------------------------------------
def lagrange_interpolate(points, evals):
    """coefficients (low -> high) of the polynomial of degree < len(points) through (points, evals)"""
    m = len(points)
    out = [0] * m
    for j in range(m):
        # numerator prod_{i != j} (X - x_i), denominator prod (x_j - x_i)
        num, den = [1], 1
        for i in range(m):
            if i == j:
                continue
            nxt = [0] * (len(num) + 1)
            for t, c in enumerate(num):
                nxt[t] = (nxt[t] - c * points[i]) % R
                nxt[t + 1] = (nxt[t + 1] + c) % R
            num = nxt
            den = den * (points[j] - points[i]) % R
        f = evals[j] * pow(den, -1, R) % R
        for t, c in enumerate(num):
            out[t] = (out[t] + c * f) % R
    return out


def eval_small(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def evaluate_vanishing_polynomial(roots, z):
    acc = 1
    for r in roots:
        acc = acc * (z - r) % R
    return acc


def construct_intermediate_sets(queries):
    """shplonk.rs construct_intermediate_sets.  queries: list of (commitment id, point, eval) in query order.
    Returns (rotation_sets, super_point_set): rotation_sets = [(points ascending, [(commitment, evals)])]."""
    super_points = sorted({q[1] for q in queries})                   # BTreeSet<F>
    commitment_sets = []                                             # Vec<(commitment, BTreeSet<point>)>
    for c, pt, _ in queries:
        for ent in commitment_sets:
            if ent[0] == c:
                ent[1].add(pt)
                break
        else:
            commitment_sets.append((c, {pt}))
    rot_map = []                                                     # Vec<(BTreeSet<point>, Vec<commitment>)>
    for c, pts in commitment_sets:
        for ent in rot_map:
            if ent[0] == pts:
                ent[1].append(c)
                break
        else:
            rot_map.append((set(pts), [c]))

    def get_eval(c, pt):
        for qc, qp, qe in queries:
            if qc == c and qp == pt:
                return qe
        raise KeyError
    sets = []
    for pts, cs_ in rot_map:
        pv = sorted(pts)
        sets.append((pv, [(c, [get_eval(c, p) for p in pv]) for c in cs_]))
    return sets, super_points


# ---- plonk/prover.rs create_proof -----------------------------------------------------------------------
def create_proof(orc, params, pk, circ, advice_values, instances, randoms, transcript=None, threads=None, trace=None):
    """One circuit instance, first-phase advice only.  advice_values: Lagrange columns as left by synthesize
    (blinding rows are overwritten here, as upstream); instances: list of lists of canonical ints;
    randoms: (count, 4) Montgomery Fr in draw order.  Returns the proof bytes."""
    cs, k = circ.cs, circ.k
    n, bf, d = 1 << k, cs.blinding_factors(), cs.degree()
    dom = pk["domain"]
    tr = transcript or Blake2bWrite()
    rng = Randoms(randoms)
    aff = lambda g1: pm.limbs_to_affine(orc.g1_to_affine(g1))[0]
    commit_l = lambda v: aff(orc.best_multiexp(v, params["g_lagrange"], threads))
    commit_c = lambda v: aff(orc.best_multiexp(v, params["g"][:len(v)], threads))
    ev = lambda poly, x: I(orc.eval_polynomial(poly, L(x)))
    note = (lambda name, v: trace.__setitem__(name, v)) if trace is not None else (lambda name, v: None)
    import time as _time
    _t = [_time.perf_counter()]

    def section(name):                                                # wall-clock per prover section (bench.py's reference arm)
        now = _time.perf_counter()
        note("t_" + name, now - _t[0])
        _t[0] = now

    tr.common_scalar(pk["transcript_repr"])                          # vk.hash_into
    assert len(instances) == cs.num_instance_columns
    instance_values = []
    for vals in instances:
        assert len(vals) <= n - (bf + 1)
        col = np.tile(L(0), (n, 1))
        if len(vals):
            col[:len(vals)] = pm.to_mont_limbs(vals, R)
        for v in vals:
            tr.common_scalar(v % R)                                   # KZG: QUERY_INSTANCE == false
        instance_values.append(col)
    instance_polys = [dom.lagrange_to_coeff(v) for v in instance_values]

    # advice: blinding rows, commitments
    advice = [np.array(a, dtype=np.uint64).reshape(n, 4).copy() for a in advice_values]
    assert len(advice) == cs.num_advice_columns
    for col in advice:
        col[n - (bf + 1):] = rng.take(bf + 1)
    rng.take(len(advice))                                             # one Blind per column (unused by KZG)
    for col in advice:
        tr.write_point(commit_l(col))
    challenges = [tr.squeeze_challenge() for _ in range(cs.num_challenges)]
    ch_l = pm.to_mont_limbs(challenges, R) if challenges else []
    theta = tr.squeeze_challenge()
    section("advice_commitments_s")

    # lookups: commit_permuted
    usable = n - (bf + 1)
    lk = []
    for li in range(len(cs.lookups)):
        cin = orc.eval_graph_rows(circ.lookup_input_graphs[li], k, pk["fixed_values"], advice, instance_values, ch_l, L(theta), threads)
        ctab = orc.eval_graph_rows(circ.lookup_table_graphs[li], k, pk["fixed_values"], advice, instance_values, ch_l, L(theta), threads)
        a_s, s_s = orc.permute_expression_pair(cin[:usable], ctab[:usable], usable)
        pin = np.concatenate([a_s, rng.take(bf + 1)])
        ptab = np.concatenate([s_s, rng.take(bf + 1)])
        pin_poly = dom.lagrange_to_coeff(pin); rng.take(1); pin_c = commit_l(pin)
        ptab_poly = dom.lagrange_to_coeff(ptab); rng.take(1); ptab_c = commit_l(ptab)
        tr.write_point(pin_c)
        tr.write_point(ptab_c)
        lk.append({"cin": cin, "ctab": ctab, "pin": pin, "ptab": ptab, "pin_poly": pin_poly, "ptab_poly": ptab_poly})
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    section("lookup_permuted_s")

    # permutation.commit
    chunk = d - 2
    pcols = cs.permutation_columns
    col_values = lambda ct, ix: (advice, pk["fixed_values"], instance_values)[ct][ix]
    z_polys, deltaomega, last_z = [], 1, L(1)
    for c0 in range(0, len(pcols), chunk):
        cols = pcols[c0:c0 + chunk]
        z = orc.permutation_product(k, [col_values(ct, ix) for ct, ix in cols], pk["sigma_values"][c0:c0 + len(cols)],
                                    L(beta), L(gamma), dom.omega, L(deltaomega), last_z, threads)
        deltaomega = deltaomega * pow(pm.DELTA, len(cols), R) % R
        z[n - bf:] = rng.take(bf)
        last_z = z[n - (bf + 1)].copy()
        rng.take(1)
        zc = commit_l(z)
        z_polys.append(dom.lagrange_to_coeff(z))
        tr.write_point(zc)

    section("permutation_product_s")
    # lookups: commit_product
    for e in lk:
        z = orc.lookup_product(k, e["cin"], e["ctab"], e["pin"], e["ptab"], L(beta), L(gamma))
        z[n - bf:] = rng.take(bf)
        rng.take(1)
        zc = commit_l(z)
        e["z_poly"] = dom.lagrange_to_coeff(z)
        tr.write_point(zc)

    section("lookup_product_s")
    # vanishing.commit: random polynomial of degree n - 1 [UPSTREAM? SURVEY.md A.10 (2)]
    random_poly = rng.take(n).copy()
    rng.take(1)
    tr.write_point(commit_c(random_poly))
    y = tr.squeeze_challenge()

    advice_polys = [dom.lagrange_to_coeff(a) for a in advice]
    flat_lookup = [p for e in lk for p in (e["z_poly"], e["pin_poly"], e["ptab_poly"])]
    h = orc.evaluate_h(circ.desc, pk["fixed_cosets"], pk["sigma_cosets"], pk["l0"], pk["l_last"], pk["l_active_row"],
                       advice_polys, instance_polys, ch_l, L(y), L(beta), L(gamma), L(theta), flat_lookup, z_polys,
                       dom.ext_n, threads)
    if trace is not None and trace.get("keep_h_ext", True):
        note("h_ext", h)
    section("evaluate_h_with_transforms_s")
    # vanishing.construct
    hc = dom.extended_to_coeff(dom.divide_by_vanishing_poly(h))
    pieces = [np.ascontiguousarray(hc[i * n:(i + 1) * n]) for i in range(hc.shape[0] // n)]
    rng.take(len(pieces))
    for p in pieces:
        tr.write_point(commit_c(p))
    x = tr.squeeze_challenge()
    xn = pow(x, n, R)
    section("quotient_commitments_s")

    rot = lambda r: rotate_omega(k, x, r)
    advice_evals = [ev(advice_polys[c], rot(r)) for c, r in cs.advice_queries]
    for e_ in advice_evals:
        tr.write_scalar(e_)
    fixed_evals = [ev(pk["fixed_polys"][c], rot(r)) for c, r in cs.fixed_queries]
    for e_ in fixed_evals:
        tr.write_scalar(e_)
    # vanishing.evaluate
    h_poly = orc.poly_lincomb(pieces, [L(pow(xn, i, R)) for i in range(len(pieces))], threads)
    random_eval = ev(random_poly, x)
    tr.write_scalar(random_eval)
    # pk.permutation.evaluate
    sigma_evals = [ev(p, x) for p in pk["sigma_polys"]]
    for e_ in sigma_evals:
        tr.write_scalar(e_)
    # permutation.evaluate
    x_next, x_last = rot(1), rot(-(bf + 1))
    perm_evals = []
    for si, zp in enumerate(z_polys):
        pe = {"x": ev(zp, x), "next": ev(zp, x_next)}
        tr.write_scalar(pe["x"]); tr.write_scalar(pe["next"])
        if si + 1 < len(z_polys):
            pe["last"] = ev(zp, x_last)
            tr.write_scalar(pe["last"])
        perm_evals.append(pe)
    # lookups.evaluate
    x_inv = rot(-1)
    for e in lk:
        e["evals"] = [ev(e["z_poly"], x), ev(e["z_poly"], x_next), ev(e["pin_poly"], x), ev(e["pin_poly"], x_inv),
                      ev(e["ptab_poly"], x)]
        for v in e["evals"]:
            tr.write_scalar(v)

    section("evaluations_s")
    # queries, in create_proof's order
    polys, queries = [], []

    def q(poly, pt, evalv):
        for i, p in enumerate(polys):
            if p is poly:
                break
        else:
            polys.append(poly); i = len(polys) - 1
        queries.append((i, pt, evalv))
    for (c, r), e_ in zip(cs.advice_queries, advice_evals):
        q(advice_polys[c], rot(r), e_)
    for zp, pe in zip(z_polys, perm_evals):                      # permutation.open
        q(zp, x, pe["x"]); q(zp, x_next, pe["next"])
    for zp, pe in list(zip(z_polys, perm_evals))[::-1][1:]:
        q(zp, x_last, pe["last"])
    for e in lk:                                                  # lookups.open
        z_, zn, a_, ai, s_ = e["evals"]
        q(e["z_poly"], x, z_); q(e["pin_poly"], x, a_); q(e["ptab_poly"], x, s_)
        q(e["pin_poly"], x_inv, ai); q(e["z_poly"], x_next, zn)
    for (c, r), e_ in zip(cs.fixed_queries, fixed_evals):
        q(pk["fixed_polys"][c], rot(r), e_)
    for p, e_ in zip(pk["sigma_polys"], sigma_evals):             # pk.permutation.open
        q(p, x, e_)
    # vanishing.open: the verifier recomputes h(x); the prover's eval is only used for r(X)
    q(h_poly, x, ev(h_poly, x)); q(random_poly, x, random_eval)

    shplonk_create_proof(orc, params, tr, polys, queries, n, threads, commit_c)
    section("shplonk_s")
    assert rng.pos == randoms_needed(cs, k), (rng.pos, randoms_needed(cs, k))
    return tr.finalize()


def shplonk_create_proof(orc, params, tr, polys, queries, n, threads, commit_c):
    """poly/kzg/multiopen/shplonk/prover.rs ProverSHPLONK::create_proof"""
    sets, super_points = construct_intermediate_sets(queries)
    y = tr.squeeze_challenge()
    v = tr.squeeze_challenge()
    zero = np.tile(L(0), (n, 1))

    def sub_low(poly, low):
        out = poly.copy()
        cur = pm.from_mont_limbs(out[:len(low)], R)
        out[:len(low)] = pm.to_mont_limbs([(a - b) % R for a, b in zip(cur, low)], R)
        return out
    quotients, combined, lows = [], [], []
    for pts, comms in sets:
        ypow = [pow(y, j, R) for j in range(len(comms))]
        acc = orc.poly_lincomb([polys[c] for c, _ in comms], [L(p) for p in ypow], threads)      # sum_j y^j P_ij
        r_polys = [lagrange_interpolate(pts, evs) for _, evs in comms]                           # low_degree_equivalent
        low = [sum(ypow[j] * r_polys[j][t] for j in range(len(comms))) % R for t in range(len(pts))]
        num = sub_low(acc, low)
        for pt in pts:                                                                           # div_by_vanishing
            num = orc.kate_division(num, L(pt))
        qpoly = zero.copy(); qpoly[:num.shape[0]] = num
        quotients.append(qpoly); combined.append(acc); lows.append((ypow, r_polys))
    h_x = orc.poly_lincomb(quotients, [L(pow(v, i, R)) for i in range(len(sets))], threads)
    tr.write_point(commit_c(h_x))
    u = tr.squeeze_challenge()
    lin, z_diffs = [], []
    for (pts, comms), acc, (ypow, r_polys) in zip(sets, combined, lows):
        z_i = evaluate_vanishing_polynomial([p for p in super_points if p not in pts], u)
        r_u = sum(ypow[j] * eval_small(r_polys[j], u) for j in range(len(comms))) % R
        lin.append(sub_low(acc, [r_u])); z_diffs.append(z_i)
    l_x = orc.poly_lincomb(lin + [h_x], [L(pow(v, i, R) * z_diffs[i] % R) for i in range(len(sets))] +
                           [L(-evaluate_vanishing_polynomial(super_points, u) % R)], threads)
    assert I(orc.eval_polynomial(l_x, L(u))) == 0
    h2 = orc.kate_division(l_x, L(u))
    qpoly = zero.copy(); qpoly[:h2.shape[0]] = h2
    h2 = orc.poly_lincomb([qpoly], [L(pow(z_diffs[0], -1, R))], threads)
    tr.write_point(commit_c(h2))
