#!/bin/bash
run() { python bench.py --k 19 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --inflight 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), 'evalh', round(d['kernels']['evaluate_h']['ms_per_step'],2))"; }
run base
ZKDCAP_B200_LIB=$PWD/zk-dcap-verifier_b200/libzkdcap_b200_e5.so run evalh_5ctas
ZKDCAP_B200_LIB=$PWD/zk-dcap-verifier_b200/libzkdcap_b200_e6.so run evalh_6ctas
