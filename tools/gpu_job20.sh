#!/bin/bash
# final 1-GPU job of the round: tests, the default bench line (all legs), k=17, launch list at the final kernels
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=8) > gpurun_out/r2_gputest_final.log 2>&1
tail -n 5 gpurun_out/r2_gputest_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python bench.py --k 17 --steps 5 --warmup 3 > gpurun_out/r2_bench_k17_final.json 2> gpurun_out/r2_bench_k17_final.err
python -c "
import json
for f in ('gpurun_out/r2_bench_final.json','gpurun_out/r2_bench_k17_final.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('e2e_dropin') or {}).get('ms_per_step'), d['roofline']['frac'], d['roofline'].get('int_pipe',{}).get('frac_of_modmul_peak'), d['cpu_baseline'].get('commitments_vs_oracle'), d['proof_sha256'][:12], {k: round(v['ms_per_step'],2) for k,v in d['kernels'].items() if v['ms_per_step']>1.0})"
ncu --metrics gpu__time_duration.sum --clock-control none -s 470 -c 480 --csv --log-file gpurun_out/r2_launches_fullproof_final.csv python tools/gpu_profile_target.py 2 > gpurun_out/ncu6.log 2>&1
tail -n 2 gpurun_out/ncu6.log
python __graft_entry__.py --smoke 2>&1 | tail -n 2
