#!/bin/bash
# 2-GPU confirmation of the default torchrun line at the final code: every leg present, in-flight batch reported
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
tail -n 3 gpurun_out/r2_bench_n2_final.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n2_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('batch'), {k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('ms_per_proof','ms_per_step','speedup_vs_one_gpu','bytes_equal','ms','p2p','proofs_identical')}) for k,v in d.items() if k.startswith('sharded')})"
