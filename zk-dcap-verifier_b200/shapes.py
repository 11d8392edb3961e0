"""
shapes.py -- constraint-system SHAPES of the reference's circuits (no witness logic).

The reference's circuits are authored in Rust against halo2-base / halo2-ecc / halo2-dynamic-sha256
(circuits/src/sgx_dcap_verifier.rs:139-238); what the GPU hot path depends on is only the
resulting ConstraintSystem: column counts, gate polynomials, lookup arguments and the
equality-enabled columns.  These builders restate that shape (counts
marked there as approximate are the same here).

  qe3_shape()        BASELINE config 1: the QE3 demo circuit (k = 17)
  full_chain_shape() BASELINE config 2: synthetic full certificate chain = same columns, k = 19
  toy_shape()        a small circuit with one gate, one lookup and a permutation, for tests
"""
from . import _abi
from .halo2 import ConstraintSystem, Expression as E


def _flex_gate(cs, q_fixed, advice_col):
    # halo2-base FlexGate "vertical" basic gate: q * (a + b * c - out), a,b,c,out = rotations 0..3 of one column
    a, b, c, out = (E.Advice(advice_col, r) for r in range(4))
    cs.create_gate([E.Fixed(q_fixed) * (a + b * c - out)])


def qe3_shape():
    """SgxDcapVerifierConfig::configure (circuits/src/sgx_dcap_verifier.rs:139-238)."""
    # advice: 12 bit-decomposition + encoded/decoded/decoded_without_gap (:140-147)
    #         FpConfig 4 gate + 1 lookup advice (configs/ecdsa_circuit.config:1)
    #         RangeConfig 3 gate + 1 lookup advice (:183-192), SHA-256 spread 2 x (dense, spread) (:195-202)
    BIT0, ENC, DEC, DECNG = 0, 12, 13, 14
    FP_GATE, FP_LOOKUP = 15, 19
    RG_GATE, RG_LOOKUP = 20, 23
    SHA = 24
    n_advice = 28
    # fixed: 7 table columns (table/mod.rs:25-31), 2 range tables, 2 spread tables (dense, spread), 2 constants,
    #        7 simple selectors (one per basic-gate column), 1 complex selector (:149)
    TAB, RT_FP, RT_RG, SPREAD_T, CONST, QSEL, QDEC = 0, 7, 8, 9, 11, 13, 20
    n_fixed = 21
    cs = ConstraintSystem(num_fixed=n_fixed, num_advice=n_advice)
    for i in range(4):
        _flex_gate(cs, QSEL + i, FP_GATE + i)
    for i in range(3):
        _flex_gate(cs, QSEL + 4 + i, RG_GATE + i)
    q = E.Fixed(QDEC)
    one_minus_q = E.Constant(1) - q
    zero = E.Constant(0)

    def gated(col):       # q * value + (1 - q) * default   (create_bit_lookup, :97-111)
        return q * E.Advice(col) + one_minus_q * zero
    # 4 lookups of a 6-bit encoded value + its 3 two-bit limbs into table columns (:216-226)
    for i in range(4):
        cols = [ENC] + [BIT0 + 3 * i + j for j in range(3)]
        cs.lookup([gated(c) for c in cols], [E.Fixed(TAB + j) for j in range(4)])
    # 3 lookups of an 8-bit decoded value + its 4 two-bit limbs (:227-236)
    for i in range(3):
        cols = [DEC] + [BIT0 + 4 * i + j for j in range(4)]
        cs.lookup([gated(c) for c in cols], [E.Fixed(TAB + 2 + j) for j in range(5)])
    # range-check lookups: lookup-advice column into the range table
    cs.lookup([E.Advice(FP_LOOKUP)], [E.Fixed(RT_FP)])
    cs.lookup([E.Advice(RG_LOOKUP)], [E.Fixed(RT_RG)])
    # SHA-256 spread tables: (dense, spread) pairs
    for i in range(2):
        cs.lookup([E.Advice(SHA + 2 * i), E.Advice(SHA + 2 * i + 1)], [E.Fixed(SPREAD_T), E.Fixed(SPREAD_T + 1)])
    # equality-enabled columns (:151-153 + the chips' gate / lookup / constant columns)
    for c in (ENC, DEC, DECNG):
        cs.enable_equality(_abi.COL_ADVICE, c)
    for c in list(range(FP_GATE, FP_GATE + 5)) + list(range(RG_GATE, RG_GATE + 4)) + list(range(SHA, SHA + 4)):
        cs.enable_equality(_abi.COL_ADVICE, c)
    for c in (CONST, CONST + 1):
        cs.enable_equality(_abi.COL_FIXED, c)
    return cs


def full_chain_shape():
    """Root -> intermediate -> PCK leaf + QE3 + ISV report signature (4 x ECDSA-P256 + SHA-256): roadmap only in
    the reference (README.md:10-51); BASELINE config 2 defines it as the config-1 column set at k = 19."""
    return qe3_shape()


def toy_shape():
    """1 gate (q*(a + b*c - out)), 1 two-column lookup with a degree-2 input, 3 equality columns."""
    cs = ConstraintSystem(num_fixed=4, num_advice=3, num_instance=1, num_challenges=1)
    _flex_gate(cs, 0, 0)
    cs.create_gate([E.Fixed(0) * (E.Advice(1) * E.Advice(1, -1) - E.Instance(0) * E.Challenge(0)) * 3])
    q = E.Fixed(1)
    cs.lookup([q * E.Advice(1), q * E.Advice(2) + (E.Constant(1) - q) * E.Constant(5)], [E.Fixed(2), E.Fixed(3)])
    cs.enable_equality(_abi.COL_ADVICE, 0)
    cs.enable_equality(_abi.COL_ADVICE, 1)
    cs.enable_equality(_abi.COL_FIXED, 2)
    cs.enable_equality(_abi.COL_INSTANCE, 0)
    return cs


def counts(cs):
    """(A, F, I, L, P, degree, blinding_factors) -- P = number of permutation sets"""
    d = cs.degree()
    chunk = d - 2
    p = (len(cs.permutation_columns) + chunk - 1) // chunk if cs.permutation_columns else 0
    return {"advice": cs.num_advice_columns, "fixed": cs.num_fixed_columns, "instance": cs.num_instance_columns,
            "lookups": len(cs.lookups), "perm_columns": len(cs.permutation_columns), "perm_sets": p,
            "degree": d, "blinding_factors": cs.blinding_factors()}


def hot_path_calls(cs):
    """Number of hot-path calls one create_proof makes for this shape:
    MSM(n) x [A + 3L + P + 1 + (d-1) + 2], iNTT(n) x [A + 3L + P], coset-NTT(ext) x [A + I + P + 3L],
    iNTT(ext) x 1, h(X) over ext rows."""
    c = counts(cs)
    A, L, P, I, d = c["advice"], c["lookups"], c["perm_sets"], c["instance"], c["degree"]
    return {"msm_lagrange": A + 3 * L + P, "msm_coeff": 1 + (d - 1) + 2, "intt_n": A + 3 * L + P,
            "coset_ntt_ext": A + I + P + 3 * L, "intt_ext": 1}


# ---- synthetic satisfying assignment of the QE3 / full-chain shape ----------------------------------------
# BASELINE configs 1/2 ask for "synthetic witnesses of the named circuit shapes" (SURVEY.md 8d).  The assignment
# below satisfies every gate, lookup and copy constraint of qe3_shape(), so the proof it leads to verifies.  It is
# workload generation, not part of the prover: the field arithmetic it needs (out = a + b*c on the enabled rows) is
# done by the `fops` object the caller passes (the device field ops in bench.py).
def _rand_fr(rng, shape_n):
    import numpy as np
    v = rng.integers(0, 1 << 63, size=(shape_n, 4), dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(shape_n, 4), dtype=np.uint64)
    v[:, 3] %= np.uint64(0x30644e72e131a029)      # < r: every such limb vector is a valid Montgomery residue
    return v


def synth_qe3_witness(cs, k, seed, fops):
    """-> (fixed_values [21 x (n,4)], advice_values [28 x (n,4)], mapping (18, n, 2) uint32).
    fops: object with mul(a, b), add(a, b), to_mont(raw) on (m, 4) uint64 arrays."""
    import numpy as np
    rng = np.random.default_rng(seed)
    n = 1 << k
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    BIT0, ENC, DEC, DECNG, FP_GATE, FP_LOOKUP, RG_GATE, RG_LOOKUP, SHA = 0, 12, 13, 14, 15, 19, 20, 23, 24
    TAB, RT_FP, RT_RG, SPREAD_T, CONST, QSEL, QDEC = 0, 7, 8, 9, 11, 13, 20
    zero = np.zeros((n, 4), dtype=np.uint64)

    def small(vals):                               # canonical small integers -> Montgomery column
        raw = np.zeros((n, 4), dtype=np.uint64)
        raw[:len(vals), 0] = np.asarray(vals, dtype=np.uint64)
        return fops.to_mont(raw)
    one = small([1] * n)
    fixed = [zero.copy() for _ in range(21)]
    advice = [_rand_fr(rng, n) for _ in range(28)]
    # lookup tables: the seven bit-table columns all hold t for t < 256 (then the default row 0), see SURVEY 8d
    tab_rows = min(256, usable)
    tvals = np.arange(n, dtype=np.uint64)
    tvals[tab_rows:] = 0
    tcol = small(tvals)
    for j in range(7):
        fixed[TAB + j] = tcol.copy()
    range_rows = min(1 << max(k - 1, 1), usable)
    rvals = np.arange(n, dtype=np.uint64)
    rvals[range_rows:] = 0
    fixed[RT_FP] = small(rvals)
    fixed[RT_RG] = fixed[RT_FP].copy()
    # spread table: dense t, spread(t) = bits of t interleaved with zeros
    srows = min(1 << min(k - 1, 16), usable)
    dense = np.arange(n, dtype=np.uint64)
    dense[srows:] = 0
    spread = np.zeros(n, dtype=np.uint64)
    for b in range(20):
        spread |= ((dense >> np.uint64(b)) & np.uint64(1)) << np.uint64(2 * b)
    fixed[SPREAD_T], fixed[SPREAD_T + 1] = small(dense), small(spread)
    fixed[CONST], fixed[CONST + 1] = _rand_fr(rng, n), _rand_fr(rng, n)
    # complex selector q on ~3/4 of the usable rows; looked-up cells take one table row per witness row
    qdec = (rng.integers(0, 4, size=n) != 0)
    qdec[usable:] = False
    fixed[QDEC][qdec] = one[qdec]
    trow = rng.integers(0, tab_rows, size=n)
    for c in list(range(BIT0, BIT0 + 12)) + [ENC, DEC]:
        advice[c][qdec] = tcol[trow][qdec]
    advice[FP_LOOKUP][:usable] = fixed[RT_FP][rng.integers(0, range_rows, size=usable)]
    advice[RG_LOOKUP][:usable] = fixed[RT_RG][rng.integers(0, range_rows, size=usable)]
    for i in range(2):
        t = rng.integers(0, srows, size=usable)
        advice[SHA + 2 * i][:usable] = fixed[SPREAD_T][t]
        advice[SHA + 2 * i + 1][:usable] = fixed[SPREAD_T + 1][t]
    # vertical basic gates q * (a + b*c - out) on rows 4g .. 4g+3, wired together by copy constraints
    pcol = {col: i for i, col in enumerate(cs.permutation_columns)}
    mapping = np.empty((len(cs.permutation_columns), n, 2), dtype=np.uint32)
    mapping[:, :, 0] = np.arange(len(cs.permutation_columns), dtype=np.uint32)[:, None]
    mapping[:, :, 1] = np.arange(n, dtype=np.uint32)[None, :]

    def tie(ca, ra, cb, rb):                        # 2-cycles between distinct cells (vectorised)
        mapping[ca, ra, 0], mapping[ca, ra, 1] = cb, rb
        mapping[cb, rb, 0], mapping[cb, rb, 1] = ca, ra
    groups = usable // 4
    gate_cols = [(FP_GATE + i, QSEL + i, FP_LOOKUP) for i in range(4)] + [(RG_GATE + i, QSEL + 4 + i, RG_LOOKUP) for i in range(3)]
    even, odd = np.arange(0, groups, 2), np.arange(1, groups, 2)
    lk_free = {FP_LOOKUP: rng.permutation(usable), RG_LOOKUP: rng.permutation(usable)}
    const_free = [rng.permutation(usable), rng.permutation(usable)]
    used = {FP_LOOKUP: 0, RG_LOOKUP: 0, "c0": 0, "c1": 0}
    for gi, (col, sel, lcol) in enumerate(gate_cols):
        fixed[sel][4 * np.arange(groups)] = one[:groups]
        # b cells of every 4th group come from a constant column, c cells of every 4th group from the lookup column
        gb = np.arange(1, groups, 4)
        cc = gi % 2
        rows_c = const_free[cc][used["c%d" % cc]:used["c%d" % cc] + len(gb)]
        used["c%d" % cc] += len(gb)
        advice[col][4 * gb + 1] = fixed[CONST + cc][rows_c]
        tie(pcol[(_abi.COL_ADVICE, col)], 4 * gb + 1, pcol[(_abi.COL_FIXED, CONST + cc)], rows_c)
        gc = np.arange(2, groups, 4)
        rows_l = lk_free[lcol][used[lcol]:used[lcol] + len(gc)]
        used[lcol] += len(gc)
        advice[col][4 * gc + 2] = advice[lcol][rows_l]
        tie(pcol[(_abi.COL_ADVICE, col)], 4 * gc + 2, pcol[(_abi.COL_ADVICE, lcol)], rows_l)
    # even groups first (free a), then odd groups take a = out of a random even group of the previous gate column
    for col, _, _ in gate_cols:
        r = 4 * even
        advice[col][r + 3] = fops.add(advice[col][r], fops.mul(advice[col][r + 1], advice[col][r + 2]))
    for gi, (col, _, _) in enumerate(gate_cols):
        src_col = gate_cols[(gi + 1) % len(gate_cols)][0]
        src = rng.permutation(even)[:len(odd)]
        r = 4 * odd[:len(src)]
        advice[col][r] = advice[src_col][4 * src + 3]
        tie(pcol[(_abi.COL_ADVICE, col)], r, pcol[(_abi.COL_ADVICE, src_col)], 4 * src + 3)
        r = 4 * odd
        advice[col][r + 3] = fops.add(advice[col][r], fops.mul(advice[col][r + 1], advice[col][r + 2]))
    # DECNG mirrors DEC on a random third of the rows; the SHA dense columns are tied pairwise on a few rows
    rows = rng.permutation(usable)[:usable // 3]
    dst = rng.permutation(usable)[:usable // 3]
    advice[DECNG][dst] = advice[DEC][rows]
    tie(pcol[(_abi.COL_ADVICE, DECNG)], dst, pcol[(_abi.COL_ADVICE, DEC)], rows)
    rows = rng.permutation(usable)[:usable // 8]
    advice[ENC][rows] = advice[ENC][rows]          # ENC stays unconstrained by copies (kept for clarity)
    return fixed, advice, mapping


def toy_prover_shape():
    """Small satisfiable circuit that exercises what the QE3 shape does not: an instance column (also in the
    permutation), a fixed column in the permutation, a rotation -1 query and a lookup with gated inputs."""
    cs = ConstraintSystem(num_fixed=4, num_advice=3, num_instance=1)
    _flex_gate(cs, 0, 0)                                            # F0 * (a + a[1]*a[2] - a[3])
    cs.create_gate([E.Fixed(1) * (E.Advice(1) - E.Instance(0))])    # F1 * (b - instance)
    cs.create_gate([E.Fixed(1) * (E.Advice(2, -1) - E.Advice(2, -1)) * 7])   # vanishes identically; queries c at -1
    q = E.Fixed(1)
    cs.lookup([q * E.Advice(1), q * E.Advice(2) + (E.Constant(1) - q) * E.Constant(5)], [E.Fixed(2), E.Fixed(3)])
    cs.enable_equality(_abi.COL_ADVICE, 0)
    cs.enable_equality(_abi.COL_ADVICE, 1)
    cs.enable_equality(_abi.COL_FIXED, 2)
    cs.enable_equality(_abi.COL_INSTANCE, 0)
    return cs


def synth_toy_witness(cs, k, seed, fops, n_instance=5):
    """-> (fixed_values, advice_values, mapping, instances [one (len, 4) array]) for toy_prover_shape()"""
    import numpy as np
    rng = np.random.default_rng(seed)
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    zero = np.zeros((n, 4), dtype=np.uint64)

    def small(vals):
        raw = np.zeros((n, 4), dtype=np.uint64)
        raw[:len(vals), 0] = np.asarray(vals, dtype=np.uint64)
        return fops.to_mont(raw)
    one = small([1] * n)
    fixed = [zero.copy() for _ in range(4)]
    advice = [_rand_fr(rng, n) for _ in range(3)]
    trows = min(usable, 16)
    t0 = _rand_fr(rng, n); t1 = _rand_fr(rng, n)
    t0[0] = 0; t1[0] = small([5])[0]
    t0[trows:] = t0[0]; t1[trows:] = t1[0]
    fixed[2], fixed[3] = t0, t1
    n_instance = min(n_instance, usable, trows - 1)
    rows = np.arange(n_instance)
    pick = 1 + rng.integers(0, trows - 1, size=n_instance)
    fixed[1][rows] = one[rows]
    instance = t0[pick].copy()
    advice[1][rows] = t0[pick]
    advice[2][rows] = t1[pick]
    groups = usable // 4
    g = np.arange(groups)
    fixed[0][4 * g] = one[:groups]
    pcol = {col: i for i, col in enumerate(cs.permutation_columns)}
    mapping = np.empty((len(cs.permutation_columns), n, 2), dtype=np.uint32)
    mapping[:, :, 0] = np.arange(len(cs.permutation_columns), dtype=np.uint32)[:, None]
    mapping[:, :, 1] = np.arange(n, dtype=np.uint32)[None, :]

    def tie(ca, ra, cb, rb):
        mapping[ca, ra, 0], mapping[ca, ra, 1] = cb, rb
        mapping[cb, rb, 0], mapping[cb, rb, 1] = ca, ra
    # a-cells of the first gates come from table cells (fixed column in the permutation)
    m = min(groups, trows)
    advice[0][4 * np.arange(m)] = t0[np.arange(m)]
    tie(pcol[(_abi.COL_ADVICE, 0)], 4 * np.arange(m), pcol[(_abi.COL_FIXED, 2)], np.arange(m))
    advice[0][4 * g + 3] = fops.add(advice[0][4 * g], fops.mul(advice[0][4 * g + 1], advice[0][4 * g + 2]))
    # instance cells are tied to the b cells they equal
    tie(pcol[(_abi.COL_INSTANCE, 0)], rows, pcol[(_abi.COL_ADVICE, 1)], rows)
    return fixed, advice, mapping, [instance]
