"""vk.bin / pk.bin (zk-dcap-verifier_b200/keyfile.py: halo2's VerifyingKey / ProvingKey byte formats in the raw SerdeFormats the
reference uses, crates/p256-ecdsa/src/base.rs:145-153, bin/src/main.rs:205-225).  The reference ships neither file, so these tests
show self-consistency only: write -> read round trips of a real key (the CPU prover's), the checks of the `raw_bytes` format, a
proof verified with the commitments read back from vk bytes; on the GPU, evaluate_h of a key loaded from pk bytes."""
import struct

import numpy as np
import pytest

import zkdcap_b200 as zk
from oracle import oracle as orc
from oracle import prover as op
from oracle import pymodel as pm
from tests import prover_common as pc
from tests.test_verifier_host import g2_bytes

kf = zk.keyfile


def key_parts(case, pk, selectors=()):
    vk = {"k": case["k"], "fixed_commitments": pm.affine_to_limbs(pk["vk"]["fixed_commitments"]),
          "permutation_commitments": pm.affine_to_limbs(pk["vk"]["permutation_commitments"]), "selectors": list(selectors)}
    return {"vk": vk, "l0": pk["l0"], "l_last": pk["l_last"], "l_active_row": pk["l_active_row"],
            "fixed_values": pk["fixed_values"], "fixed_polys": pk["fixed_polys"], "fixed_cosets": pk["fixed_cosets"],
            "permutations": pk["sigma_values"], "permutation_polys": pk["sigma_polys"], "permutation_cosets": pk["sigma_cosets"]}


def same(a, b):
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    return np.array_equal(np.asarray(a), np.asarray(b))


@pytest.fixture(scope="module")
def toy():
    case = pc.make_case("toy", 5, 3)
    return case, pc.oracle_key(case)


@pytest.mark.parametrize("fmt", kf.FORMATS)
def test_vk_round_trip(toy, fmt):
    case, pk = toy
    rng = np.random.default_rng(5)
    sels = [rng.integers(0, 2, 32).astype(bool) for _ in range(3)]
    parts = key_parts(case, pk, sels)
    data = kf.write_vk(parts["vk"], fmt)
    npc = len(case["cs"].permutation_columns)
    assert len(data) == 8 + 64 * (case["cs"].num_fixed_columns + npc) + 3 * 4
    assert data[:4] == struct.pack(">I", 5) and data[4:8] == struct.pack(">I", case["cs"].num_fixed_columns)
    back = kf.read_vk(data, npc, 3, fmt)
    assert same(back, parts["vk"])
    assert kf.write_vk(back, fmt) == data
    with pytest.raises(kf.KeyFileError):
        kf.read_vk(data + b"\0", npc, 3, fmt)
    with pytest.raises(kf.KeyFileError):
        kf.read_vk(data[:-1], npc, 3, fmt)
    with pytest.raises(kf.KeyFileError):
        kf.read_vk(data, npc, 3, "processed")


def test_selector_bits_are_packed_lsb_first():
    sel = np.zeros(16, dtype=bool)
    sel[0] = sel[9] = True
    vk = {"k": 4, "fixed_commitments": np.zeros((0, 8), np.uint64), "permutation_commitments": np.zeros((0, 8), np.uint64), "selectors": [sel]}
    data = kf.write_vk(vk)
    assert data[8:] == bytes([0x01, 0x02])
    assert np.array_equal(kf.read_vk(data, 0, 1)["selectors"][0], sel)


@pytest.mark.parametrize("fmt", kf.FORMATS)
def test_pk_round_trip(toy, fmt):
    case, pk = toy
    parts = key_parts(case, pk)
    data = kf.write_pk(parts, fmt)
    npc = len(case["cs"].permutation_columns)
    ext_k = 5 + int(np.ceil(np.log2(case["cs"].degree() - 1)))
    back = kf.read_pk(data, npc, 0, fmt, extended_k=ext_k)
    assert same(back, parts)
    assert kf.write_pk(back, fmt) == data
    with pytest.raises(kf.KeyFileError):
        kf.read_pk(data, npc, 0, fmt, extended_k=ext_k + 1)           # polynomial lengths are checked against the domain
    with pytest.raises(kf.KeyFileError):
        kf.read_pk(data, npc + 1, 0, fmt)
    with pytest.raises(kf.KeyFileError):
        kf.read_pk(data[:-32], npc, 0, fmt)
    with pytest.raises(kf.KeyFileError):
        kf.read_pk(data + bytes(4), npc, 0, fmt)


def test_checked_format_rejects_what_unchecked_lets_through(toy):
    case, pk = toy
    parts = key_parts(case, pk)
    npc = len(case["cs"].permutation_columns)
    # a scalar that is not below r
    bad = dict(parts)
    l0 = np.array(parts["l0"], dtype=np.uint64, copy=True)
    l0[3] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)
    bad["l0"] = l0
    data = kf.write_pk(bad)
    kf.read_pk(data, npc, 0, "raw_bytes_unchecked")
    with pytest.raises(kf.KeyFileError):
        kf.read_pk(data, npc, 0, "raw_bytes")
    # a commitment that is not on the curve
    vk = dict(parts["vk"])
    fc = np.array(vk["fixed_commitments"], dtype=np.uint64, copy=True)
    fc[0, 4] ^= np.uint64(1)
    vk["fixed_commitments"] = fc
    data = kf.write_vk(vk)
    kf.read_vk(data, npc, 0, "raw_bytes_unchecked")
    with pytest.raises(kf.KeyFileError):
        kf.read_vk(data, npc, 0, "raw_bytes")
    # the identity is a valid commitment (an all-zero fixed column)
    fc[0] = 0
    kf.read_vk(kf.write_vk(vk), npc, 0, "raw_bytes")


def test_proof_verifies_with_the_commitments_read_from_vk_bytes(toy):
    case, pk = toy
    proof = op.create_proof(orc, pc.params_for(5), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    data = kf.write_vk(key_parts(case, pk)["vk"])
    vk = kf.read_vk(data, len(case["cs"].permutation_columns), 0, "raw_bytes")
    vp = pc.verifier_params()
    g1 = pm.affine_to_limbs([(1, 2)])[0]
    assert zk.verify_proof(case["circ"], vk["fixed_commitments"], vk["permutation_commitments"], g1, g2_bytes(vp["g2"]),
                           g2_bytes(vp["s_g2"]), case["instances_limbs"], proof)
    tampered = np.array(vk["permutation_commitments"], copy=True)
    tampered[0] = vk["fixed_commitments"][0]
    assert not zk.verify_proof(case["circ"], vk["fixed_commitments"], tampered, g1, g2_bytes(vp["g2"]), g2_bytes(vp["s_g2"]),
                               case["instances_limbs"], proof)


@pytest.mark.gpu
def test_proving_key_read_from_bytes_evaluates_like_the_original():
    """ProvingKey.read(pk bytes): evaluate_h of the loaded key == evaluate_h of the key built from the arrays == the oracle's"""
    from tests.test_evalh_host import make_case, oracle_h, L
    shapes = __import__("importlib").import_module("zk-dcap-verifier_b200.shapes")
    ctx = zk.Context(0)
    cs, k = shapes.toy_shape(), 6
    case = make_case(cs, k, 4, orc)
    want, desc, keep = oracle_h(orc, cs, k, case)
    n, ext_n = 1 << k, case["l0"].shape[0]
    zeros = [np.zeros((n, 4), np.uint64)] * len(case["fixed_cosets"])
    zs = [np.zeros((n, 4), np.uint64)] * len(case["sigma_cosets"])
    parts = {"vk": {"k": k, "fixed_commitments": np.zeros((len(case["fixed_cosets"]), 8), np.uint64),
                    "permutation_commitments": np.zeros((len(case["sigma_cosets"]), 8), np.uint64), "selectors": []},
             "l0": case["l0"], "l_last": case["l_last"], "l_active_row": case["l_active"],
             "fixed_values": zeros, "fixed_polys": zeros, "fixed_cosets": case["fixed_cosets"],
             "permutations": zs, "permutation_polys": zs, "permutation_cosets": case["sigma_cosets"]}
    data = kf.write_pk(parts)
    key = zk.ProvingKey.read(data, cs, ctx=ctx)
    ch = pm.to_mont_limbs(case["challenges"], pm.R_MOD)
    args = (case["advice"], case["instance"], ch, L(case["y"]), L(case["beta"]), L(case["gamma"]), L(case["theta"]),
            case["lookup_polys"], case["z_polys"])
    got = key.evaluate_h(*args)
    assert np.array_equal(got, want)
    ref = zk.ProvingKey(cs, k, case["fixed_cosets"], case["sigma_cosets"], case["l0"], case["l_last"], case["l_active"], ctx=ctx)
    assert np.array_equal(ref.evaluate_h(*args), got)
    assert key.file["vk"]["k"] == k and ext_n == key.domain.ext_n


def test_pinning_json():
    text = """[ {"k": 18, "num_advice_per_phase": [5], "num_fixed": 1, "num_lookup_advice_per_phase": [1, 0, 0], "lookup_bits": 17,
                "num_instance_columns": 1}, [[262134, 262133, 262135, 262134]] ]"""
    params, breaks = kf.read_pinning(text)
    assert params["k"] == 18 and params["lookup_bits"] == 17 and breaks == [[262134, 262133, 262135, 262134]]
    for bad in ("{}", "[1, 2]", '[{"k": 18}, []]', '[{"k":18,"num_advice_per_phase":[1],"num_fixed":1,"num_lookup_advice_per_phase":[1],'
                '"lookup_bits":8,"num_instance_columns":1}, [["x"]]]'):
        with pytest.raises(kf.KeyFileError):
            kf.read_pinning(bad)
