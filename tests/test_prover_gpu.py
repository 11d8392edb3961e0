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
