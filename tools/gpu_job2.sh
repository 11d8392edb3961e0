#!/bin/bash
# round-2 GPU job 2: tests, ncu captures of the hot kernels at this round's code, launch list, reference arm
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/r2_gputest_b.log 2>&1
tail -4 gpurun_out/r2_gputest_b.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_msm_accumulate -s 8 -c 2 -f -o gpurun_out/r2_msm_accumulate python tools/gpu_profile_target.py 2 > gpurun_out/ncu1.log 2>&1
ZK_NTT_WARP=1 $NCU -k regex:k_ntt_pass_w -s 150 -c 3 -f -o gpurun_out/r2_ntt_pass_w python tools/gpu_profile_target.py 2 > gpurun_out/ncu2.log 2>&1
ZK_NTT_SMEM=1 $NCU -k regex:k_ntt_pass_t -s 150 -c 3 -f -o gpurun_out/r2_ntt_pass_t python tools/gpu_profile_target.py 2 > gpurun_out/ncu3.log 2>&1
$NCU -k regex:k_evalh -s 1 -c 1 -f -o gpurun_out/r2_evalh python tools/gpu_profile_target.py 2 > gpurun_out/ncu4.log 2>&1
$NCU -k regex:k_msm_digits_smem -s 16 -c 2 -f -o gpurun_out/r2_msm_digits python tools/gpu_profile_target.py 2 > gpurun_out/ncu5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file gpurun_out/r2_launches_fullproof.csv python tools/gpu_profile_target.py 2 > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu6.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
(time python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_b.json", "gpurun_out/r2_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("proof_sha256"), d.get("proof_verified"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_bench_reference.err
