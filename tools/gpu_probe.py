"""First-contact probe on the GPU box: integer-pipe micro-benchmarks + per-kernel timings at bench sizes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkdcap_b200 as zk

ctx = zk.Context(0)
out = {}
for it in (500, 4000):
    out["imad_lo_T_%d" % it] = ctx.bench_imad(0, it)
    out["imad_wide_T_%d" % it] = ctx.bench_imad(1, it)
    out["modmul_fr_G_%d" % it] = ctx.bench_modmul(0, it)
    out["modmul_fq_G_%d" % it] = ctx.bench_modmul(1, it)
print(json.dumps(out, indent=1))
open("gpurun_out/probe.json", "w").write(json.dumps(out))
