#!/bin/bash
# round-2 GPU job 3: tests with the generated evaluate_h, A/B against the interpreter, ncu of the generated kernel
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/r2_gputest_c.log 2>&1
tail -6 gpurun_out/r2_gputest_c.log
ZK_EVALH=interp python bench.py --steps 5 --warmup 3 --no-dropin --no-cpu-baseline --inflight 1 > gpurun_out/r2_bench_evalh_interp.json 2> gpurun_out/r2_bench_evalh_interp.err
python bench.py --steps 5 --warmup 3 --no-dropin --no-cpu-baseline --inflight 1 > gpurun_out/r2_bench_evalh_gen.json 2> gpurun_out/r2_bench_evalh_gen.err
ZK_NTT_SMEM=1 python bench.py --steps 5 --warmup 3 --no-dropin --no-cpu-baseline --inflight 1 > gpurun_out/r2_bench_c_nttsmem.json 2> gpurun_out/r2_bench_c_nttsmem.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_evalh_interp.json", "gpurun_out/r2_bench_evalh_gen.json", "gpurun_out/r2_bench_c_nttsmem.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), d.get("evaluate_h"), {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items() if v["ms_per_step"] > 3}, d.get("proof_sha256")[:16])
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --set full --clock-control none --import-source on -k regex:zk_evalh_gen -s 1 -c 1 -f -o gpurun_out/r2_evalh_gen python tools/gpu_profile_target.py 2 > gpurun_out/ncu7.log 2>&1
tail -3 gpurun_out/ncu7.log
tail -3 gpurun_out/r2_bench_evalh_gen.err
