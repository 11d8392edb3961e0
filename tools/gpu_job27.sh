#!/bin/bash
# 4-GPU default torchrun line (the one world size not yet run this round)
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/r2_bench_n4_final.err
grep -v "OMP_NUM\|^\*\*\*\|NCCL version" gpurun_out/r2_bench_n4_final.err | tail -n 6
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n4_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {a:b for a,b in d['batch'].items() if a!='note'}, {k:{a:b for a,b in v.items() if a in ('ms_per_proof','ms','Gpts_s','bytes_equal_single_gpu_proof_on_all_ranks','matches_single_gpu')} for k,v in d.items() if k.startswith('sharded') and v})"
