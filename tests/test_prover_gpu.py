"""GPU: zk_keygen / zk_create_proof (csrc/prover.cu) against the CPU restatement of halo2's prover
(oracle/prover.py) on the same circuit, witness and pre-drawn RNG stream: the proof BYTES must be identical, the
vk commitments identical, and the oracle's verifier (real pairing) must accept the device's proof."""
import numpy as np
import pytest

import zkdcap_b200 as zk
from oracle import oracle as orc
from oracle import prover as op
from oracle import pymodel as pm
from oracle import verifier as ov
from tests import prover_common as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return zk.Context(0)


def first_diff(a, b):
    for i in range(0, max(len(a), len(b)), 32):
        if a[i:i + 32] != b[i:i + 32]:
            return i // 32
    return -1


@pytest.mark.parametrize("shape,k,seed", [("toy", 5, 1), ("toy", 8, 2), ("qe3", 9, 1), ("qe3", 11, 3)])
def test_proof_bytes_match_the_cpu_prover(ctx, shape, k, seed):
    case = pc.make_case(shape, k, seed)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    opk = pc.oracle_key(case)
    fc, pcm = key.vk_commitments()
    assert pm.limbs_to_affine(fc) == opk["vk"]["fixed_commitments"]
    assert pm.limbs_to_affine(pcm) == opk["vk"]["permutation_commitments"]
    assert key.randoms_needed == op.randoms_needed(case["cs"], k)
    want = op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    got = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    assert len(got) == len(want) and got == want, "first differing 32-byte item: %d of %d" % (first_diff(got, want), len(want) // 32)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], case["cs"], k, case["instances_int"], got)
    # the library's own verifier, fed with the device's vk commitments
    from tests.test_verifier_host import g2_bytes
    vp = pc.verifier_params()
    assert zk.verify_proof(case["circ"], fc, pcm, pm.affine_to_limbs([(1, 2)])[0], g2_bytes(vp["g2"]), g2_bytes(vp["s_g2"]),
                           case["instances_limbs"], got)
    # a second proof on the same key (work area reuse) with another RNG stream
    rnd2 = orc.gen_scalars(4242, key.randoms_needed)
    got2 = key.create_proof(case["advice"], case["instances_limbs"], rnd2)
    assert got2 == op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], rnd2)


def test_lookup_failure_is_reported(ctx):
    case = pc.make_case("toy", 5, 4)
    p = pc.params_for(5)
    params = zk.ParamsKZG(5, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    adv = [a.copy() for a in case["advice"]]
    adv[2][0] = pm.to_mont_limbs([123456789], pm.R_MOD)[0]          # (b, c) on row 0 is no longer a table row
    with pytest.raises(zk.ZkError, match="ConstraintSystemFailure"):
        key.create_proof(adv, case["instances_limbs"], case["randoms"])
    with pytest.raises(zk.ZkError):
        key.create_proof(case["advice"], case["instances_limbs"], case["randoms"][:10])


def test_sort_fallback_with_colliding_high_digits(ctx):
    """Lookup inputs that agree on their top 64 bits but differ below them force the full-width radix sort."""
    case = pc.make_case("toy", 6, 5)
    cs, k = case["cs"], 6
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    # table values t0[j] = BIG + j: the compressed key (theta*t0 + t1) still varies in its low digits only when theta
    # is small -- instead exercise the path directly through a single-expression shape
    from importlib import import_module
    E = zk.Expression
    cs2 = zk.ConstraintSystem(num_fixed=1, num_advice=1)
    cs2.lookup([E.Advice(0)], [E.Fixed(0)])
    cs2.enable_equality(0, 0)
    # pairs of table values share bits 64..190 (the 8 most significant varying digits and more) and differ only in
    # their low 64 bits: the top-digit passes cannot order them, so the full-width sort has to run
    hrng = np.random.default_rng(3)
    hi = [int(hrng.integers(1, 1 << 60)) * (1 << 130) + int(hrng.integers(1, 1 << 60)) * (1 << 70) for _ in range(n // 2)]
    tab = [(hi[j // 2] + int(hrng.integers(0, 1 << 62))) % pm.R_MOD for j in range(n)]
    tab[usable:] = [tab[0]] * (n - usable)
    rng = np.random.default_rng(9)
    inp = [tab[int(t)] for t in rng.integers(0, usable, size=n)]
    fixed = [pm.to_mont_limbs(tab, pm.R_MOD)]
    advice = [pm.to_mont_limbs(inp, pm.R_MOD)]
    mapping = np.zeros((1, n, 2), dtype=np.uint32)
    mapping[0, :, 1] = np.arange(n)
    circ = zk.Circuit(cs2, k, transcript_repr=77)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, circ, fixed, mapping)
    opk = op.keygen(orc, p, circ, fixed, mapping, circ.transcript_repr)
    rnd = orc.gen_scalars(5, key.randoms_needed)
    want = op.create_proof(orc, p, opk, circ, advice, [], rnd)
    got = key.create_proof(advice, [], rnd)
    assert got == want, "first differing item %d" % first_diff(got, want)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], cs2, k, [], got)


def _random_witness_case(cs, k, seed, lookup_table_cols=()):
    """Random fixed / advice columns (no satisfying assignment: the provers do not check satisfiability, byte equality
    of the two provers is what is tested); lookup inputs are forced into their tables by the caller."""
    rng = np.random.default_rng(seed)
    n = 1 << k
    rnd = lambda: pc.shapes._rand_fr(rng, n)
    fixed = [rnd() for _ in range(cs.num_fixed_columns)]
    advice = [rnd() for _ in range(cs.num_advice_columns)]
    mapping = np.empty((len(cs.permutation_columns), n, 2), dtype=np.uint32)
    mapping[:, :, 0] = np.arange(len(cs.permutation_columns), dtype=np.uint32)[:, None]
    mapping[:, :, 1] = np.arange(n, dtype=np.uint32)[None, :]
    return fixed, advice, mapping


def _both_provers(ctx, cs, k, fixed, advice, mapping, instances_int, seed):
    circ = zk.Circuit(cs, k, transcript_repr=1000 + seed)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, circ, fixed, mapping)
    opk = op.keygen(orc, p, circ, fixed, mapping, circ.transcript_repr)
    rnd = orc.gen_scalars(seed, key.randoms_needed)
    want = op.create_proof(orc, p, opk, circ, advice, instances_int, rnd)
    got = key.create_proof(advice, [pm.to_mont_limbs(i, pm.R_MOD) for i in instances_int], rnd)
    assert got == want, "first differing item %d of %d" % (first_diff(got, want), len(want) // 32)
    return got, opk


def test_circuit_without_lookups_or_permutation(ctx):
    """only custom gates: no lookup argument, no permutation argument (empty product / z lists everywhere)"""
    E = zk.Expression
    cs = zk.ConstraintSystem(num_fixed=2, num_advice=3)
    cs.create_gate([E.Fixed(0) * (E.Advice(0) * E.Advice(1) - E.Advice(2)), E.Fixed(1) * (E.Advice(0, 1) - E.Advice(2, -1)) * 5])
    k = 6
    n = 1 << k
    fixed, advice, mapping = _random_witness_case(cs, k, 11)
    # make it satisfying so that the verifier can be asked too: q0 on even rows with c = a*b, q1 off
    one = pm.to_mont_limbs([1], pm.R_MOD)[0]
    fixed[0][:] = 0; fixed[1][:] = 0
    fixed[0][0:n - 8:2] = one
    advice[2] = orc.field_op("fr", "mul", advice[0], advice[1])
    got, opk = _both_provers(ctx, cs, k, fixed, advice, mapping, [], 21)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], cs, k, [], got)


def test_circuit_with_permutation_only(ctx):
    E = zk.Expression
    cs = zk.ConstraintSystem(num_fixed=1, num_advice=2)
    cs.create_gate([E.Fixed(0) * (E.Advice(0) - E.Advice(1))])
    cs.enable_equality(0, 0)
    cs.enable_equality(0, 1)
    k = 5
    n = 1 << k
    fixed, advice, mapping = _random_witness_case(cs, k, 12)
    fixed[0][:] = 0
    advice[1][:10] = advice[0][:10]
    for r in range(10):      # cell (0, r) <-> cell (1, r)
        mapping[0, r] = (1, r); mapping[1, r] = (0, r)
    got, opk = _both_provers(ctx, cs, k, fixed, advice, mapping, [], 22)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], cs, k, [], got)


def test_rotations_that_wrap_onto_the_same_opening_point(ctx):
    """at small n two rotations can be congruent mod n (here Advice(0, 1) and Advice(0, -7) at n = 8, and
    -(bf + 1) = -6 = 2 mod 8 against the query rotation 2).  halo2 keys opening points by their field value, so the two
    rotations share one point; the device prover used to refuse such circuits."""
    E = zk.Expression
    cs = zk.ConstraintSystem(num_fixed=2, num_advice=2)
    cs.create_gate([E.Fixed(0) * (E.Advice(0, 1) - E.Advice(0, -7)), E.Fixed(0) * (E.Advice(1, 2) - E.Advice(1, 2))])
    cs.enable_equality(0, 0)
    cs.enable_equality(0, 1)
    cs.enable_equality(1, 1)                 # degree 3 -> one column per permutation set -> three sets, x_last is opened
    k = 3
    assert (-(cs.blinding_factors() + 1)) % (1 << k) in {r % (1 << k) for _, r in cs.advice_queries}
    fixed, advice, mapping = _random_witness_case(cs, k, 31)
    got, opk = _both_provers(ctx, cs, k, fixed, advice, mapping, [], 31)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], cs, k, [], got)
    from tests.test_verifier_host import g2_bytes
    fc = pm.affine_to_limbs(opk["vk"]["fixed_commitments"]); pcm = pm.affine_to_limbs(opk["vk"]["permutation_commitments"])
    vp = pc.verifier_params()
    assert zk.verify_proof(zk.Circuit(cs, k, transcript_repr=1000 + 31), fc, pcm, pm.affine_to_limbs([(1, 2)])[0], g2_bytes(vp["g2"]),
                           g2_bytes(vp["s_g2"]), [], got)


def test_first_phase_challenge_and_instance(ctx):
    """shapes.toy_shape(): a gate that reads a challenge and an instance column, instance column in the permutation"""
    cs = pc.shapes.toy_shape()
    k = 6
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    fixed, advice, mapping = _random_witness_case(cs, k, 13)
    # the lookup is (F1*A1, F1*A2 + (1-F1)*5) in (F2, F3): switch it off (F1 = 0) and put (0, 5) into the table
    fixed[1][:] = 0
    fixed[2][0] = 0
    fixed[3][0] = pm.to_mont_limbs([5], pm.R_MOD)[0]
    inst = [[3, 1, 4, 1, 5, 9, 2, 6]]
    got, opk = _both_provers(ctx, cs, k, fixed, advice, mapping, inst, 23)
    assert len(got) > 0


@pytest.mark.parametrize("shape,k", [("toy", 6), ("qe3", 9)])
def test_poseidon_transcript_proof_bytes(ctx, shape, k):
    """zk_pkey_set_transcript(POSEIDON): snark-verifier's PoseidonTranscript (gen_proof, crates/p256-ecdsa/src/base.rs:200-211):
    bytes equal to the CPU prover's under the same transcript, accepted by both verifiers"""
    case = pc.make_case(shape, k, 12)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    key.set_transcript("poseidon")
    opk = pc.oracle_key(case)
    want = op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"], transcript=op.PoseidonWrite())
    got = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    assert got == want, "first differing item %d" % first_diff(got, want)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], case["cs"], k, case["instances_int"], got, transcript="poseidon")
    fc, pcm = key.vk_commitments()
    vp = pc.verifier_params()
    g2b = lambda q: b"".join(v.to_bytes(32, "little") for v in (q[0][0], q[0][1], q[1][0], q[1][1]))
    assert zk.verify_proof(case["circ"], fc, pcm, pm.affine_to_limbs([(1, 2)])[0], g2b(vp["g2"]), g2b(vp["s_g2"]),
                           case["instances_limbs"], got, transcript="poseidon")
    assert not zk.verify_proof(case["circ"], fc, pcm, pm.affine_to_limbs([(1, 2)])[0], g2b(vp["g2"]), g2b(vp["s_g2"]),
                               case["instances_limbs"], got, transcript="blake2b")


def test_evm_transcript_proof_bytes(ctx):
    """zk_pkey_set_transcript(EVM): Keccak-256 transcript, uncompressed big-endian points"""
    case = pc.make_case("toy", 6, 8)
    p = pc.params_for(6)
    params = zk.ParamsKZG(6, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    key.set_transcript("evm")
    opk = pc.oracle_key(case)
    want = op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"], transcript=op.EvmWrite())
    got = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    assert got == want, "first differing item %d" % first_diff(got, want)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], case["cs"], 6, case["instances_int"], got, transcript="evm")
    key.set_transcript("blake2b")
    assert key.create_proof(case["advice"], case["instances_limbs"], case["randoms"]) == \
        op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"])


# ---- random constraint systems ------------------------------------------------------------------------------------------
def _random_expression(rng, cs_dims, depth, max_degree):
    """random expression tree over the columns; degree kept <= max_degree"""
    E = zk.Expression
    nf, na, ni, nc = cs_dims

    def leaf():
        kinds = ["fixed", "advice", "advice", "const"] + (["instance"] if ni else []) + (["challenge"] if nc else [])
        kind = kinds[rng.integers(0, len(kinds))]
        rot = int(rng.integers(-2, 3))
        if kind == "fixed":
            return E.Fixed(int(rng.integers(0, nf)), rot), 1
        if kind == "advice":
            return E.Advice(int(rng.integers(0, na)), rot), 1
        if kind == "instance":
            return E.Instance(int(rng.integers(0, ni)), rot), 1
        if kind == "challenge":
            return E.Challenge(int(rng.integers(0, nc))), 0
        return E.Constant(int(rng.integers(0, 1 << 62))), 0

    def build(d):
        if d == 0 or rng.integers(0, 4) == 0:
            return leaf()
        op_ = rng.integers(0, 5)
        a, da = build(d - 1)
        if op_ == 0:
            return -a, da
        if op_ == 1:
            return a * int(rng.integers(1, 1 << 30)), da
        b, db = build(d - 1)
        if op_ == 2 and da + db <= max_degree:
            return a * b, da + db
        if op_ == 3:
            return a - b, max(da, db)
        return a + b, max(da, db)
    return build(depth)


def _random_circuit(seed):
    rng = np.random.default_rng(seed)
    E = zk.Expression
    nf, na = int(rng.integers(3, 6)), int(rng.integers(2, 6))
    ni, nc = int(rng.integers(0, 3)), int(rng.integers(0, 2))
    cs = zk.ConstraintSystem(num_fixed=nf, num_advice=na, num_instance=ni, num_challenges=nc)
    max_degree = int(rng.integers(2, 10))
    for _ in range(int(rng.integers(1, 4))):
        polys = []
        for _ in range(int(rng.integers(1, 4))):
            e, d = _random_expression(rng, (nf, na, ni, nc), 3, max_degree - 1)
            for _ in range(int(rng.integers(0, max(1, max_degree - d)))):      # raise the degree with more column factors
                e = e * E.Advice(int(rng.integers(0, na)), int(rng.integers(-1, 2)))
            polys.append(E.Fixed(int(rng.integers(0, nf))) * e)
        cs.create_gate(polys)
    # lookups: q * advice columns into fixed table columns (fixed column 0 is the selector, the last ones are tables)
    lookups = []
    for _ in range(int(rng.integers(0, 4))):
        width = int(rng.integers(1, 3))
        a_cols = [int(c) for c in rng.choice(na, size=min(width, na), replace=False)]
        t_cols = [int(c) for c in rng.choice(np.arange(1, nf), size=len(a_cols), replace=False)]
        q = E.Fixed(0)
        cs.lookup([q * E.Advice(a) for a in a_cols], [E.Fixed(t) for t in t_cols])
        lookups.append((a_cols, t_cols))
    cols = [(0, i) for i in range(na)] + [(1, i) for i in range(nf)] + [(2, i) for i in range(ni)]
    for ci in rng.permutation(len(cols))[:int(rng.integers(0, len(cols) + 1))]:
        cs.enable_equality(*cols[int(ci)])
    return cs, lookups, rng


@pytest.mark.parametrize("seed", list(range(101, 125)))
def test_random_constraint_systems(ctx, seed):
    """random gates (rotations -2..2, challenges, instances), 0-3 lookup arguments, random equality columns and random
    copy cycles: the two provers must write the same bytes.  The gates are not satisfied (neither prover checks them);
    lookup inputs are drawn from their tables so that permute_expression_pair succeeds."""
    cs, lookups, rng = _random_circuit(seed)
    k = int(rng.integers(5, 8))
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    if usable < 4:
        pytest.skip("degree too high for this k")
    fixed, advice, mapping = _random_witness_case(cs, k, seed)
    if lookups:
        # selector column 0: random bits; every table gets the all-zero row 0 (the inputs of unselected rows)
        one = pm.to_mont_limbs([1], pm.R_MOD)[0]
        sel = rng.integers(0, 2, size=n).astype(bool)
        fixed[0][:] = 0
        fixed[0][sel] = one
        tab_rows = int(rng.integers(2, max(3, usable // 2)))
        claimed = set()
        for a_cols, t_cols in lookups:
            for t in t_cols:
                fixed[t][0] = 0
                fixed[t][tab_rows:] = 0
            a_free = [a for a in a_cols if a not in claimed]
            pick = rng.integers(0, tab_rows, size=n)
            for a, t in zip(a_cols, t_cols):
                if a in a_free:
                    advice[a][sel] = fixed[t][pick][sel]
                    claimed.add(a)
        # a column shared by two lookups may miss its second table; rebuild consistently: every advice column used by
        # more than one lookup is zeroed on selected rows (row 0 of every table is zero)
        seen = {}
        for a_cols, t_cols in lookups:
            for a in a_cols:
                seen[a] = seen.get(a, 0) + 1
        shared = {a for a, c in seen.items() if c > 1}
        for a_cols, t_cols in lookups:
            if any(a in shared for a in a_cols):
                for a in a_cols:
                    advice[a][sel] = 0
        # tables are fixed columns that may themselves be selector 0?  no: t_cols are drawn from 1..nf-1
    # random copy cycles (2-cycles between random cells of the equality columns)
    P = len(cs.permutation_columns)
    if P:
        cells = rng.permutation(P * n)[:2 * int(rng.integers(0, 3 * n // 4))]
        for a, b in zip(cells[0::2], cells[1::2]):
            ca, ra, cb, rb = int(a) // n, int(a) % n, int(b) // n, int(b) % n
            mapping[ca, ra] = (cb, rb)
            mapping[cb, rb] = (ca, ra)
    inst = [[int(v) for v in rng.integers(0, 1 << 62, size=int(rng.integers(0, min(usable, 9))))] for _ in range(cs.num_instance_columns)]
    _both_provers(ctx, cs, k, fixed, advice, mapping, inst, seed)


@pytest.mark.parametrize("seed", list(range(201, 211)))
def test_random_satisfied_circuits(ctx, seed):
    """random circuits with a satisfying witness: same bytes as the CPU prover, accepted by the Python verifier and by
    zk_verify_proof; a broken gate output is rejected"""
    from tests.test_verifier_host import native_verify
    case = pc.random_satisfied_case(seed)
    k = case["k"]
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    opk = pc.oracle_key(case)
    want = op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    got = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    assert got == want, "first differing item %d" % first_diff(got, want)
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], case["cs"], k, case["instances_int"], got)
    assert native_verify(case, opk, case["instances_limbs"], got)
    fc, pcm = key.vk_commitments()
    assert pm.limbs_to_affine(fc) == list(opk["vk"]["fixed_commitments"])
    assert pm.limbs_to_affine(pcm) == list(opk["vk"]["permutation_commitments"])
    if case["first_gate_rows"]:
        adv = [a.copy() for a in case["advice"]]
        adv[case["first_output"]][case["first_gate_rows"][0], 0] ^= np.uint64(1)
        bad = key.create_proof(adv, case["instances_limbs"], case["randoms"])
        assert not native_verify(case, opk, case["instances_limbs"], bad)
