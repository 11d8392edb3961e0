#!/bin/bash
# round-2 GPU job 8 (1 GPU): tests + bench at the current kernels, window-width probe for large single MSMs, compute-sanitizer
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/r2_gputest_d.log 2>&1
tail -n 5 gpurun_out/r2_gputest_d.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v['ms_per_step'],2) for k,v in d['kernels'].items() if v['ms_per_step']>1.2}, d['roofline']['traffic'], d['proof_sha256'][:12])"
for c in 16 18 20; do
  ZK_MSM_C=$c python bench.py --mode sweep --sweep-min 24 --sweep-max 24 --sweep-coset-max 0 2> gpurun_out/r2_sweep_c$c.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('C', d['window_bits'], d['windows'], 'k', d['k'], round(d['msm_resident_ms'],2), 'ms', round(d['msm_resident_Gpts_s'],3), 'Gpts/s', d['msm_resident_parity'][:20])"
done
# compute-sanitizer: the c = 16 digit sort / row-column reduction on a small input, one complete proof, the generated evaluate_h
SAN="compute-sanitizer --error-exitcode 9"
(timeout 900 $SAN --tool memcheck python -m pytest tests/test_gpu_production.py -q -x -k "window16 and 13" ) > gpurun_out/r2_sanitizer_memcheck_msm_c16.log 2>&1; echo "memcheck msm rc=$?"
(timeout 900 $SAN --tool racecheck python -m pytest tests/test_gpu_production.py -q -x -k "window16 and 13" ) > gpurun_out/r2_sanitizer_racecheck_msm_c16.log 2>&1; echo "racecheck msm rc=$?"
(timeout 900 $SAN --tool memcheck python tools/gpu_sanitize_prover.py ) > gpurun_out/r2_sanitizer_memcheck_prover.log 2>&1; echo "memcheck prover rc=$?"
(timeout 900 $SAN --tool racecheck python tools/gpu_sanitize_prover.py ) > gpurun_out/r2_sanitizer_racecheck_prover.log 2>&1; echo "racecheck prover rc=$?"
tail -n 4 gpurun_out/r2_sanitizer_*.log
