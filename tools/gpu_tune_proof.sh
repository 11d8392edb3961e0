#!/bin/bash
# A/B runs of alternative kernel builds (ZKDCAP_B200_LIB) / env tunables on the complete-proof bench
run() { python bench.py --k 19 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --inflight 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), 'ntt', round(d['kernels']['ntt_pass']['ms_per_step'],2), 'verified', d['proof_verified'])"; }
run base
for v in u2c4 u2c3 u4c3; do ZKDCAP_B200_LIB=$PWD/zk-dcap-verifier_b200/libzkdcap_b200_$v.so run $v; done
