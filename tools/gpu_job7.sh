#!/bin/bash
# round-2 GPU job 7 (8 GPUs): group-of-eight proof test, the default bench at N=8 (replicas + sharded proof + sharded MSM)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k all_gpus) > gpurun_out/r2_gputest_multi_n8.log 2>&1
tail -n 5 gpurun_out/r2_gputest_multi_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 8 --steps 5 --warmup 3 --mode proof-sharded --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_proof_sharded_n8.json 2> gpurun_out/r2_bench_proof_sharded_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n8.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d.get("sharded_proof")), json.dumps({k: v for k, v in d.get("sharded_msm").items() if k != "result_point"}))
    d = json.loads(open("gpurun_out/r2_bench_proof_sharded_n8.json").read().strip().splitlines()[-1])
    print("sharded", d["ms_per_step"], {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.4})
except Exception as e:
    print("ERR", e)
PY
tail -n 3 gpurun_out/r2_bench_n8.err gpurun_out/r2_bench_proof_sharded_n8.err
