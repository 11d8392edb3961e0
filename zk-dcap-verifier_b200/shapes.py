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
