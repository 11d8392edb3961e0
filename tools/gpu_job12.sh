#!/bin/bash
# A/B on one box: dedicated squaring on/off, batch size of the coset transforms on the second stream, NTT pass variants
mkdir -p gpurun_out
: > gpurun_out/r2_ab.txt
run() {  # label, env assignments...
  label=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --no-dropin --no-cpu-baseline --inflight 1 --no-e2e 2>> gpurun_out/r2_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('$label', round(d['ms_per_step'],2), round(d['ms_per_step_one_stream'],2), 'ntt', round(k['ntt_pass']['ms_per_step'],2), 'acc', round(k['msm_accumulate']['ms_per_step'],2), 'part', round(k['msm_acc_partials']['ms_per_step'],2), 'evalh', round(k['evaluate_h']['ms_per_step'],2), 'binv', round(k['batch_invert']['ms_per_step'],2), 'modmul', round(d['int_pipe']['modmul_peak_G_per_s'],1), d['proof_sha256'][:10])" | tee -a gpurun_out/r2_ab.txt
}
V=$PWD/zk-dcap-verifier_b200/build/variants
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2_gputest_f.log 2>&1
tail -n 4 gpurun_out/r2_gputest_f.log
run default X=1
run nosqr ZKDCAP_B200_LIB=$V/libzk_nosqr.so
run extbatch1 ZK_PROVER_EXT_BATCH=1
run extbatch2 ZK_PROVER_EXT_BATCH=2
run extbatch16 ZK_PROVER_EXT_BATCH=16
run ntt_u2c2 ZKDCAP_B200_LIB=$V/libzk_ntt_u2c2.so
run ntt_u2c3 ZKDCAP_B200_LIB=$V/libzk_ntt_u2c3.so
run ntt_u1c6 ZKDCAP_B200_LIB=$V/libzk_ntt_u1c6.so
run default2 X=1
tail -n 3 gpurun_out/r2_ab.err
