#!/bin/bash
run() { python bench.py --k 19 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), 'ntt', round(d['kernels']['ntt_pass']['ms_per_step'],2), d['int_pipe']['ntt_frac_of_modmul_peak'])"; }
run radix4_3ctas
ZK_NTT_RADIX4=0 run radix2
