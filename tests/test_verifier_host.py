"""CPU: zk_verify_proof (csrc/verifier.cu: transcript replay, expected h(x), SHPLONK, its own BN254 pairing) against the
oracle's provers and verifier.  The C++ verifier and oracle/verifier.py share no code (C++ tower arithmetic on the
Montgomery limbs vs Python integers), so agreement on acceptance AND on every rejection below pins both."""
import random

import numpy as np
import pytest

import zkdcap_b200 as zk
from oracle import oracle as orc
from oracle import prover as op
from oracle import pymodel as pm
from oracle import verifier as ov
from tests import prover_common as pc


def g2_bytes(q):
    return b"".join(v.to_bytes(32, "little") for v in (q[0][0], q[0][1], q[1][0], q[1][1]))


def native_verify(case, pk, instances_limbs, proof, s_g2=None, transcript="blake2b"):
    vp = pc.verifier_params()
    fc = pm.affine_to_limbs(pk["vk"]["fixed_commitments"])
    pcm = pm.affine_to_limbs(pk["vk"]["permutation_commitments"])
    g1 = pm.affine_to_limbs([(1, 2)])[0]
    return zk.verify_proof(case["circ"], fc, pcm, g1, g2_bytes(vp["g2"]), g2_bytes(s_g2 or vp["s_g2"]), instances_limbs, proof,
                           transcript=transcript)


@pytest.mark.parametrize("shape,k", [("toy", 5), ("qe3", 9)])
def test_native_verifier_agrees_with_the_oracle(shape, k):
    case = pc.make_case(shape, k, 3)
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(k), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    assert ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], k, case["instances_int"], proof)
    assert native_verify(case, pk, case["instances_limbs"], proof)
    rnd = random.Random(k)
    for _ in range(12 if shape == "toy" else 4):          # any flipped bit: both reject
        bad = bytearray(proof)
        bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        assert not native_verify(case, pk, case["instances_limbs"], bytes(bad))
    assert not native_verify(case, pk, case["instances_limbs"], proof[:-32])          # truncated
    assert not native_verify(case, pk, case["instances_limbs"], proof + bytes(32))    # trailing bytes
    assert not native_verify(case, pk, case["instances_limbs"], b"")
    # a different SRS secret on the G2 side must break the pairing equation
    from oracle import pairing
    assert not native_verify(case, pk, case["instances_limbs"], proof, s_g2=pairing.g2_mul(pc.TRAPDOOR + 1, pairing.G2_GEN))
    if case["instances_limbs"]:
        wrong = [case["instances_limbs"][0].copy()]
        wrong[0][0] = pm.to_mont_limbs([12345], pm.R_MOD)[0]
        assert not native_verify(case, pk, wrong, proof)


def test_native_verifier_rejects_an_unsatisfied_witness():
    case = pc.make_case("toy", 5, 4)
    pk = pc.oracle_key(case)
    adv = [a.copy() for a in case["advice"]]
    adv[0][3] = adv[0][7]
    proof = op.create_proof(orc, pc.params_for(5), pk, case["circ"], adv, case["instances_int"], case["randoms"])
    assert not native_verify(case, pk, case["instances_limbs"], proof)


def test_evm_transcript_proofs():
    """the Keccak / EVM transcript (uncompressed big-endian points): oracle prover -> both verifiers"""
    case = pc.make_case("toy", 5, 6)
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(5), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"],
                            transcript=op.EvmWrite())
    assert ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], 5, case["instances_int"], proof, transcript="evm")
    assert native_verify(case, pk, case["instances_limbs"], proof, transcript="evm")
    assert not native_verify(case, pk, case["instances_limbs"], proof, transcript="blake2b")
    bad = bytearray(proof); bad[200] ^= 2
    assert not native_verify(case, pk, case["instances_limbs"], bytes(bad), transcript="evm")
    assert not ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], 5, case["instances_int"], bytes(bad), transcript="evm")


@pytest.mark.parametrize("seed", list(range(201, 211)))
def test_random_satisfied_circuits_are_accepted_by_both_verifiers(seed):
    """random gates / lookups / copy constraints WITH a satisfying witness (tests/prover_common.random_satisfied_case):
    both verifiers accept the proof; breaking one gate output, one looked-up cell or one copied cell makes both reject"""
    case = pc.random_satisfied_case(seed)
    k = case["k"]
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(k), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    assert ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], k, case["instances_int"], proof)
    assert native_verify(case, pk, case["instances_limbs"], proof)
    # break a gate: an output cell on a row whose selector is on
    cs = case["cs"]

    def rejected(adv):
        bad = op.create_proof(orc, pc.params_for(k), pk, case["circ"], adv, case["instances_int"], case["randoms"])
        return not ov.verify_proof(pc.verifier_params(), pk["vk"], cs, k, case["instances_int"], bad) and \
            not native_verify(case, pk, case["instances_limbs"], bad)
    if case["first_gate_rows"]:
        adv = [a.copy() for a in case["advice"]]
        adv[case["first_output"]][case["first_gate_rows"][0], 0] ^= np.uint64(1)
        assert rejected(adv)
    # break a copy constraint: one cell of a 2-cycle between advice cells
    pcols = cs.permutation_columns
    cyc = [(c, r) for c in range(len(pcols)) for r in range(1 << k)
           if tuple(case["mapping"][c, r]) != (c, r) and pcols[c][0] == 0]
    if cyc:
        c, r = cyc[0]
        adv = [a.copy() for a in case["advice"]]
        adv[pcols[c][1]][r, 0] ^= np.uint64(1)
        assert rejected(adv)


@pytest.mark.parametrize("seed", [211, 212, 213])
def test_random_satisfied_circuits_with_the_evm_transcript(seed):
    """the same random circuits through the Keccak / EVM transcript: both verifiers accept, neither accepts the proof
    under the other transcript"""
    case = pc.random_satisfied_case(seed)
    k = case["k"]
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(k), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"],
                            transcript=op.EvmWrite())
    assert ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], k, case["instances_int"], proof, transcript="evm")
    assert native_verify(case, pk, case["instances_limbs"], proof, transcript="evm")
    assert not native_verify(case, pk, case["instances_limbs"], proof, transcript="blake2b")
    assert not native_verify(case, pk, case["instances_limbs"], proof[:-1], transcript="evm")


# ---- ParamsKZG.g2 / s_g2 encodings (csrc/verifier.cu: zk_g2_decode / zk_g2_encode) -------------------------------------------
def _g2_compress(q, v04=False):
    """[UPSTREAM? halo2curves G2Affine::to_bytes] x.c0 | x.c1 little-endian, parity of y.c0 in bit 7 (0.3.1) / bit 6 (>= 0.4)"""
    b = bytearray(q[0][0].to_bytes(32, "little") + q[0][1].to_bytes(32, "little"))
    b[63] |= (q[1][0] & 1) << (6 if v04 else 7)
    return bytes(b)


def test_g2_encodings_round_trip():
    from oracle import pairing
    rnd = random.Random(5)
    pts = [pairing.G2_GEN, pairing.g2_mul(pc.TRAPDOOR, pairing.G2_GEN)] + [pairing.g2_mul(rnd.randrange(pm.R_MOD), pairing.G2_GEN) for _ in range(6)]
    negs = [(q[0], ((-q[1][0]) % pm.P_MOD, (-q[1][1]) % pm.P_MOD)) for q in pts]
    for q in pts + negs:
        want = g2_bytes(q)
        for fmt, v04 in (("processed", False), ("processed_v04", True)):
            enc = _g2_compress(q, v04)
            assert zk.g2_decode(enc, fmt) == want
            assert zk.g2_encode(want, fmt) == enc
        raw = b"".join(pm.to_mont_limbs([v], pm.P_MOD)[0].tobytes() for v in (q[0][0], q[0][1], q[1][0], q[1][1]))
        assert zk.g2_decode(raw, "raw_bytes") == want and zk.g2_encode(want, "raw_bytes") == raw
    # x with no point on the twist, a non-canonical coordinate, wrong lengths
    x = 5
    bad = None
    while bad is None:
        cand = x.to_bytes(32, "little") + bytes(32)
        try:
            zk.g2_decode(cand, "processed")
        except zk.ZkError:
            bad = cand
        x += 1
    with pytest.raises(zk.ZkError):
        zk.g2_decode(pm.P_MOD.to_bytes(32, "little") + bytes(32), "processed")
    with pytest.raises(zk.ZkError):
        zk.g2_decode(bytes(128), "processed")
    with pytest.raises(zk.ZkError):
        zk.g2_encode(bytes(127) + b"\x01", "processed")          # not on the curve


def test_verify_proof_rejects_file_encoded_g2():
    """the 64-byte compressed (or Montgomery-limb) g2 of an SRS file must not reach zk_verify_proof"""
    case = pc.make_case("toy", 5, 3)
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(5), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    vp = pc.verifier_params()
    fc = pm.affine_to_limbs(pk["vk"]["fixed_commitments"]); pcm = pm.affine_to_limbs(pk["vk"]["permutation_commitments"])
    g1 = pm.affine_to_limbs([(1, 2)])[0]
    c2, cs2 = _g2_compress(vp["g2"]), _g2_compress(vp["s_g2"])
    with pytest.raises(zk.ZkError, match="128 uncompressed bytes"):
        zk.verify_proof(case["circ"], fc, pcm, g1, c2, cs2, case["instances_limbs"], proof)
    assert zk.verify_proof(case["circ"], fc, pcm, g1, zk.g2_decode(c2), zk.g2_decode(cs2), case["instances_limbs"], proof)


def test_encode_calldata():
    """instances as 32-byte big-endian words followed by the proof (snark-verifier's encode_calldata)"""
    inst = [pm.to_mont_limbs([1, 2, pm.R_MOD - 1], pm.R_MOD), pm.to_mont_limbs([7], pm.R_MOD)]
    data = zk.encode_calldata(inst, b"\xaa\xbb")
    words = [int.from_bytes(data[32 * i:32 * i + 32], "big") for i in range(4)]
    assert words == [1, 2, pm.R_MOD - 1, 7] and data[128:] == b"\xaa\xbb" and len(data) == 130
    assert zk.encode_calldata([], b"") == b""
