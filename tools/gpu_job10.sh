#!/bin/bash
# round-2 GPU job 10 (2 GPUs): all GPU tests (incl. multi-GPU) at the current code, N=2 default bench + proof-sharded table
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/r2_gputest_e.log 2>&1
tail -n 5 gpurun_out/r2_gputest_e.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 5 --warmup 3 --mode proof-sharded --no-cpu-baseline > gpurun_out/r2_bench_proof_sharded_n2.json 2> gpurun_out/r2_bench_proof_sharded_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], json.dumps(d.get("sharded_proof")), json.dumps({k: v for k, v in d.get("sharded_msm").items() if k != "result_point"})[:300])
    d = json.loads(open("gpurun_out/r2_bench_proof_sharded_n2.json").read().strip().splitlines()[-1])
    print("sharded", d["ms_per_step"], {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.5})
except Exception as e:
    print("ERR", e)
PY
tail -n 3 gpurun_out/r2_bench_n2.err gpurun_out/r2_bench_proof_sharded_n2.err
