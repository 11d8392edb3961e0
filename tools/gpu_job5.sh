#!/bin/bash
# round-2 GPU job 5 (1 GPU): ncu of all accumulate launches of one proof, launch list at the final kernels, sweeps with parity
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_msm_accumulate -s 8 -c 8 -f -o gpurun_out/r2_msm_accumulate python tools/gpu_profile_target.py 2 > gpurun_out/ncu1.log 2>&1
tail -n 2 gpurun_out/ncu1.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 470 -c 480 --csv --log-file gpurun_out/r2_launches_fullproof_v2.csv python tools/gpu_profile_target.py 2 > gpurun_out/ncu6.log 2>&1
tail -n 2 gpurun_out/ncu6.log
(time python bench.py --mode sweep --sweep-min 16 --sweep-max 26 --sweep-classes) > gpurun_out/r2_sweep.jsonl 2> gpurun_out/r2_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_sweep.jsonl"):
    try:
        d = json.loads(l)
        print(d["k"], d["window_bits"], round(d["msm_resident_ms"], 3), round(d["msm_resident_Gpts_s"], 3), round(d["ntt_ms"], 3), round(d["ntt_GB_s"], 1), round(d.get("coset_ntt_ms", 0), 3), d["parity"])
    except Exception as e:
        print("ERR", e, l[:200])
PY
tail -n 5 gpurun_out/r2_sweep.err
python bench.py --k 17 --steps 5 --warmup 3 > gpurun_out/r2_bench_k17.json 2> gpurun_out/r2_bench_k17.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_k17.json').read().strip().splitlines()[-1]); print('k17', d['ms_per_step'], d['e2e']['ms_per_step'], d['cpu_baseline'].get('commitments_vs_oracle'))"
