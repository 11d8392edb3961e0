#!/bin/bash
# A/B of NTT pass variants (unroll / resident CTAs) on the complete proof, same box
mkdir -p gpurun_out
: > gpurun_out/r2_ntt_variants.txt
for v in default u2c2 u2c3 u4c2 u1c6 default; do
  if [ $v = default ]; then unset ZKDCAP_B200_LIB; else export ZKDCAP_B200_LIB=$PWD/zk-dcap-verifier_b200/build/variants/libzk_ntt_$v.so; fi
  python bench.py --steps 5 --warmup 3 --no-dropin --no-cpu-baseline --inflight 1 --no-e2e 2>> gpurun_out/r2_ntt_variants.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],2), round(d['ms_per_step_one_stream'],2), 'ntt', round(d['kernels']['ntt_pass']['ms_per_step'],2), 'acc', round(d['kernels']['msm_accumulate']['ms_per_step'],2), 'evalh', round(d['kernels']['evaluate_h']['ms_per_step'],2), 'lincomb', round(d['kernels']['poly_lincomb']['ms_per_step'],2), d['proof_sha256'][:10])" | tee -a gpurun_out/r2_ntt_variants.txt
done
tail -n 3 gpurun_out/r2_ntt_variants.err
