#!/bin/bash
# round-2 GPU job 1: full GPU test suite (no -x), A/B of the NTT pass kernels, baseline bench
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=15) > gpurun_out/r2_gputest_a.log 2>&1
tail -8 gpurun_out/r2_gputest_a.log
ZK_NTT_SMEM=1 python bench.py --steps 5 --warmup 3 --no-dropin --no-cpu-baseline > gpurun_out/r2_bench_nttsmem.json 2> gpurun_out/r2_bench_nttsmem.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_nttsmem.json", "gpurun_out/r2_bench_a.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items() if v["ms_per_step"] > 1})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_bench_a.err
