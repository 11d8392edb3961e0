"""evaluate_h timing on the full-chain shape for the current environment (ZK_EVALH_* toggles)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zkdcap_b200 as zk
from importlib import import_module
import bench
h2 = import_module("zk-dcap-verifier_b200.halo2"); shapes = import_module("zk-dcap-verifier_b200.shapes")
k = int(os.environ.get("TUNE_K", "17"))
dev = torch.device("cuda", 0)
ctx = zk.Context(0, bench.work_stream(torch, dev))
cs = shapes.full_chain_shape(); c = shapes.counts(cs); n = 1 << k
rng = np.random.default_rng(1)
pk = zk.ProvingKey(cs, k, [bench.rand_fr_host(rng, n) for _ in range(c["fixed"])], [bench.rand_fr_host(rng, n) for _ in range(c["perm_columns"])],
                   bench.rand_fr_host(rng, n), bench.rand_fr_host(rng, n), bench.rand_fr_host(rng, n), coeff_form=True, ctx=ctx)
gen = torch.Generator(device=dev); gen.manual_seed(1)
nw = c["advice"] + c["perm_sets"] + 3 * c["lookups"]
w = bench.rand_fr(torch, nw * n, dev, gen).view(nw, n, 4)
out = torch.empty((pk.domain.ext_n, 4), dtype=torch.int64, device=dev)
A, L, P = c["advice"], c["lookups"], c["perm_sets"]
adv = [w[i].data_ptr() for i in range(A)]; zs = [w[A + i].data_ptr() for i in range(P)]
look = [w[A + P + i].data_ptr() for i in range(3 * L)]
ch = bench.rand_fr_host(rng, 8)
f = lambda: h2.evaluate_h_dev(pk, adv, [], [], ch[0], ch[1], ch[2], ch[3], look, zs, out.data_ptr())
for _ in range(2): f()
ctx.timing(True); ctx.timing_reset()
for _ in range(5): f()
st = ctx.timing_read()
E = pk.domain.ext_n; ps = pk.program_stats()
ms = st["evaluate_h"][1] / 5
print(json.dumps({"env": {e: os.environ[e] for e in os.environ if e.startswith("ZK_")}, "k": k, "evalh_ms": ms, "Gmul_s": E * ps["muls_per_row"] / ms / 1e6, **ps}))
