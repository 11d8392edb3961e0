#!/bin/bash
run() { python bench.py --k 19 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --inflight 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$1', round(d['ms_per_step'],2), {x:round(k[x]['ms_per_step'],2) for x in k if x.startswith('msm_acc') or x.startswith('msm_pair')})"; }
run base
for M in 4 8 16 32 128; do ZK_MSM_BA=1 ZK_MSM_BA_M=$M run BA,M=$M; done
