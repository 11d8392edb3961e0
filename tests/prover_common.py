"""Shared fixtures of the prover tests: synthetic circuits, SRS with a known trapdoor, oracle keys."""
from importlib import import_module

import numpy as np

import zkdcap_b200 as zk
from oracle import oracle as orc
from oracle import pairing
from oracle import prover as op
from oracle import pymodel as pm

shapes = import_module("zk-dcap-verifier_b200.shapes")

# s = Fr::random of ChaCha20Rng::from_seed([0; 32]); any value would do for the tests
TRAPDOOR = 0x1c59a59b6cff4308740943526ade1d8c09f71b337a67269cc89586bcdd6dfcba


class OrcOps:
    """field ops for the synthetic witness generator, backed by the CPU oracle"""
    mul = staticmethod(lambda a, b: orc.field_op("fr", "mul", a, b))
    add = staticmethod(lambda a, b: orc.field_op("fr", "add", a, b))
    to_mont = staticmethod(lambda a: orc.field_op("fr", "from_raw", a))


_cache = {}


def params_for(k):
    if ("params", k) not in _cache:
        _cache[("params", k)] = op.setup_params(orc, k, TRAPDOOR)
    return _cache[("params", k)]


def verifier_params():
    if "vp" not in _cache:
        _cache["vp"] = {"g0": (1, 2), "g2": pairing.G2_GEN, "s_g2": pairing.g2_mul(TRAPDOOR, pairing.G2_GEN)}
    return _cache["vp"]


def make_case(shape, k, seed):
    """-> dict(cs, circ, fixed, advice, mapping, instances_limbs, instances_int, randoms)"""
    if shape == "qe3":
        cs = shapes.qe3_shape()
        fixed, advice, mapping = shapes.synth_qe3_witness(cs, k, seed, OrcOps)
        inst = []
    else:
        cs = shapes.toy_prover_shape()
        fixed, advice, mapping, inst = shapes.synth_toy_witness(cs, k, seed, OrcOps)
    circ = zk.Circuit(cs, k, transcript_repr=0xDCA9000 + seed)
    randoms = orc.gen_scalars(1000 + seed, op.randoms_needed(cs, k))
    return {"cs": cs, "k": k, "circ": circ, "fixed": fixed, "advice": advice, "mapping": mapping,
            "instances_limbs": inst, "instances_int": [pm.from_mont_limbs(i, pm.R_MOD) for i in inst], "randoms": randoms}


def oracle_key(case):
    return op.keygen(orc, params_for(case["k"]), case["circ"], case["fixed"], case["mapping"], case["circ"].transcript_repr)
