#!/bin/bash
# round-2 GPU job 4 (2 GPUs): group-of-two proofs (P2P exchange and NCCL), the default bench at N=2 with its sharded objects
mkdir -p gpurun_out
nvidia-smi -L
(time python -m pytest tests/test_multi_gpu.py -m gpu -q -x) > gpurun_out/r2_gputest_multi.log 2>&1
tail -15 gpurun_out/r2_gputest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], json.dumps(d.get("sharded_proof")), json.dumps(d.get("sharded_msm"))[:600], d.get("e2e"))
except Exception as e:
    print("ERR", e)
PY
tail -5 gpurun_out/r2_bench_n2.err
