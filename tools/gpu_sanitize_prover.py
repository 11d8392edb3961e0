"""compute-sanitizer target: one small create_proof (toy shape, k = 5 and the QE3 shape at k = 9) through the C ABI."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkdcap_b200 as zk
from tests import prover_common as pc

ctx = zk.Context(0)
for shape, k in (("toy", 5), ("qe3", 9)):
    case = pc.make_case(shape, k, 1)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    proof = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    print(shape, k, len(proof))
    del key, params
