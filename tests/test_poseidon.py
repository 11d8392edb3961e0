"""CPU: the Poseidon transcript of the reference's non-EVM proofs (bin/src/main.rs:239-252, crates/p256-ecdsa/src/base.rs:200-211).

* parameter generation + permutation, Python (oracle/poseidon.py) and C++ (csrc/poseidon.hpp through zk_poseidon_permute), against
  two published values of circomlib's Poseidon -- one of them at the transcript's own parameter set (t = 5, R_F = 8, R_P = 60);
* the sponge framing (snark-verifier `Poseidon::squeeze`): C++ against Python on every buffer length around the RATE boundary;
* complete proofs written with the transcript by the oracle prover: accepted by the Python verifier and by zk_verify_proof,
  rejected under any other transcript kind and after any corruption."""
import ctypes as C
import random

import numpy as np
import pytest

import zkdcap_b200 as zk
from oracle import oracle as orc
from oracle import poseidon as ps
from oracle import prover as op
from oracle import pymodel as pm
from oracle import verifier as ov
from tests import prover_common as pc
from tests.test_verifier_host import native_verify

R = pm.R_MOD
# circomlibjs test/poseidon.js: poseidon([1, 2]) (t = 3, 8 + 57 rounds) and poseidon([1, 2, 3, 4]) (t = 5, 8 + 60 rounds);
# circomlib's state is (0, inputs...), its output state[0]
KAT_T3 = 0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a
KAT_T5 = 0x299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465


def lib():
    return zk._abi.load()


def native_permute(state, t, r_f, r_p):
    buf = np.ascontiguousarray(pm.to_mont_limbs([int(v) % R for v in state], R))
    assert lib().zk_poseidon_permute(buf.ctypes.data, t, r_f, r_p) == 0
    return pm.from_mont_limbs(buf, R)


def native_hash(elements, squeezes=1):
    buf = np.ascontiguousarray(pm.to_mont_limbs([int(v) % R for v in elements], R)) if elements else np.zeros((1, 4), np.uint64)
    out = np.zeros((squeezes, 4), dtype=np.uint64)
    assert lib().zk_poseidon_hash(buf.ctypes.data if elements else None, len(elements), out.ctypes.data, squeezes) == 0
    return pm.from_mont_limbs(out, R)


def test_permutation_known_answers_python():
    assert ps.permute([0, 1, 2], t=3, r_f=8, r_p=57)[0] == KAT_T3
    assert ps.permute([0, 1, 2, 3, 4], t=5, r_f=8, r_p=60)[0] == KAT_T5


def test_permutation_known_answers_native():
    assert native_permute([0, 1, 2], 3, 8, 57)[0] == KAT_T3
    assert native_permute([0, 1, 2, 3, 4], 5, 8, 60)[0] == KAT_T5


def test_parameters_are_well_formed():
    rc, mds = ps.spec()
    assert len(rc) == 68 and all(len(r) == 5 for r in rc)
    flat = [c for r in rc for c in r]
    assert len(set(flat)) == len(flat) and all(0 < c < R for c in flat)
    # Cauchy matrix: every square submatrix of size 1 and 2 is invertible (spot check of the MDS property)
    for i in range(5):
        for j in range(5):
            assert mds[i][j] % R
            for i2 in range(i + 1, 5):
                for j2 in range(j + 1, 5):
                    assert (mds[i][j] * mds[i2][j2] - mds[i][j2] * mds[i2][j]) % R


@pytest.mark.parametrize("t,r_f,r_p", [(5, 8, 60), (3, 8, 57), (4, 8, 56), (2, 8, 56)])
def test_native_permutation_equals_python(t, r_f, r_p):
    rnd = random.Random(t * 1000 + r_p)
    for _ in range(4):
        st = [rnd.randrange(R) for _ in range(t)]
        assert native_permute(st, t, r_f, r_p) == ps.permute(st, t, r_f, r_p)
    edge = [0] * t
    assert native_permute(edge, t, r_f, r_p) == ps.permute(edge, t, r_f, r_p)
    edge = [R - 1] * t
    assert native_permute(edge, t, r_f, r_p) == ps.permute(edge, t, r_f, r_p)


@pytest.mark.parametrize("t,r_f,r_p", [(5, 8, 60), (3, 8, 57), (2, 8, 56)])
def test_optimised_form_is_the_same_function(t, r_f, r_p):
    """sparse-matrix form of the partial rounds (what snark-verifier executes; derived independently in oracle/poseidon.py and in
    csrc/poseidon.hpp, where zk_poseidon_permute also checks it against the plain rounds on every call)"""
    rnd = random.Random(31 * t)
    for _ in range(3):
        st = [rnd.randrange(R) for _ in range(t)]
        assert ps.permute_optimized(st, t, r_f, r_p) == ps.permute(st, t, r_f, r_p)
    o = ps.optimized_spec(t, r_f, r_p)
    assert len(o["sparse"]) == r_p and len(o["post"]) == r_p and o["post"][-1] == 0


def test_native_permute_rejects_bad_arguments():
    buf = np.zeros((5, 4), dtype=np.uint64)
    assert lib().zk_poseidon_permute(None, 5, 8, 60) != 0
    assert lib().zk_poseidon_permute(buf.ctypes.data, 1, 8, 60) != 0
    assert lib().zk_poseidon_permute(buf.ctypes.data, 5, 7, 60) != 0
    assert lib().zk_poseidon_permute(buf.ctypes.data, 17, 8, 60) != 0


def test_sponge_native_equals_python():
    rnd = random.Random(77)
    for count in range(0, 14):                      # empty buffer, short chunks, exact multiples of RATE
        elems = [rnd.randrange(R) for _ in range(count)]
        s = ps.Sponge()
        s.update(elems)
        want = [s.squeeze(), s.squeeze(), s.squeeze()]          # later squeezes run on an empty buffer and the carried state
        assert native_hash(elems, 3) == want
    assert ps.hash_elements([1, 2, 3, 4]) != ps.hash_elements([1, 2, 3, 4, 0])          # the padding separates lengths
    assert ps.hash_elements([1, 2, 3]) != ps.hash_elements([1, 2, 3, 0])


def test_sponge_framing():
    """the restated `squeeze`: state (2^64, 0, ..), chunk into state[1..], +1 after a short chunk, an extra permutation after an
    exact multiple of RATE, output state[1]"""
    st = [1 << 64, 0, 0, 0, 0]
    a = [5, 6]
    st1 = ps.permute([st[0], 5, 6, 1, 0])
    assert ps.hash_elements(a) == st1[1]
    b = [1, 2, 3, 4]
    st2 = ps.permute([st[0], 1, 2, 3, 4])
    st2 = ps.permute([st2[0], (st2[1] + 1) % R, st2[2], st2[3], st2[4]])
    assert ps.hash_elements(b) == st2[1]
    assert ps.hash_elements([]) == ps.permute([st[0], 1, 0, 0, 0])[1]


def test_transcript_absorbs_points_mod_r():
    tr = op.PoseidonWrite()
    big_x = pm.P_MOD - 1                              # a coordinate above r must be reduced (fe_to_fe)
    tr.sponge.update([big_x % R, 2])
    want = tr.sponge.squeeze()
    tr2 = op.PoseidonWrite()
    tr2.common_point((big_x, 2))
    assert tr2.squeeze_challenge() == want
    assert big_x >= R


@pytest.mark.parametrize("shape,k,seed", [("toy", 5, 6), ("qe3", 8, 2)])
def test_poseidon_transcript_proofs(shape, k, seed):
    case = pc.make_case(shape, k, seed)
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(k), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"],
                            transcript=op.PoseidonWrite())
    plain = op.create_proof(orc, pc.params_for(k), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    assert len(proof) == len(plain) and proof != plain
    assert ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], k, case["instances_int"], proof, transcript="poseidon")
    assert native_verify(case, pk, case["instances_limbs"], proof, transcript="poseidon")
    assert not native_verify(case, pk, case["instances_limbs"], proof, transcript="blake2b")
    assert not native_verify(case, pk, case["instances_limbs"], plain, transcript="poseidon")
    assert not ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], k, case["instances_int"], plain, transcript="poseidon")
    rnd = random.Random(seed)
    for _ in range(6 if shape == "toy" else 2):
        bad = bytearray(proof)
        bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        assert not native_verify(case, pk, case["instances_limbs"], bytes(bad), transcript="poseidon")
        assert not ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], k, case["instances_int"], bytes(bad), transcript="poseidon")
    assert not native_verify(case, pk, case["instances_limbs"], proof[:-32], transcript="poseidon")
    assert not native_verify(case, pk, case["instances_limbs"], proof + bytes(32), transcript="poseidon")


def test_poseidon_proof_points_use_the_golden_encoding():
    """the commitments of a Poseidon-transcript proof are written like those of the reference's bin/assets/proof.bin: y's parity
    in bit 6 of the last byte, bit 7 clear"""
    case = pc.make_case("toy", 5, 9)
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(5), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"],
                            transcript=op.PoseidonWrite())
    n_adv = case["cs"].num_advice_columns
    flagged = 0
    for i in range(n_adv):
        chunk = proof[32 * i:32 * i + 32]
        assert not chunk[31] & 0x80
        pt = op.decompress_point_v04(chunk)
        assert (pt[1] * pt[1] - pt[0] ** 3 - 3) % pm.P_MOD == 0
        flagged += (chunk[31] >> 6) & 1
        assert ((chunk[31] >> 6) & 1) == (pt[1] & 1)
    assert n_adv > 0
