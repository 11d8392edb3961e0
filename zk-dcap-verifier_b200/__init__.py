"""zk-dcap-verifier_b200 -- B200-native backend for the halo2 hot path of zk-dcap-verifier.

The directory name carries a hyphen, so import it with importlib (or through the
`zkdcap_b200` shim at the repository root):

    import importlib; zk = importlib.import_module("zk-dcap-verifier_b200")

Only csrc/ (hand-written sm_100a CUDA + the C ABI) computes; this package is the host-side
mirror of the halo2 interface above that ABI.  There is no CPU fallback.
"""
from . import _abi
from . import keyfile
from ._abi import ZkError, LIB_PATH
from .halo2 import (encode_calldata, g2_decode, g2_encode, Circuit, Context, ConstraintSystem, EvaluationDomain, Evaluator, Expression, GraphEvaluator, Group, ParamsKZG,
                    ProverKey, ProvingKey, best_fft, best_multiexp, default_context, fq_to_limbs, fr_to_limbs, g1_compress, g1_decompress, g1_sum, verify_proof)

__all__ = ["Circuit", "Context", "ConstraintSystem", "EvaluationDomain", "Evaluator", "Expression", "GraphEvaluator", "Group",
           "ParamsKZG", "ProverKey", "ProvingKey", "best_fft", "best_multiexp", "default_context", "fq_to_limbs", "fr_to_limbs",
           "g1_compress", "g1_decompress", "g1_sum", "verify_proof", "ZkError", "LIB_PATH", "keyfile", "g2_decode", "g2_encode", "encode_calldata"]
