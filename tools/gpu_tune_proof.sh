#!/bin/bash
run() { python bench.py --k $K --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('k=$K $1', round(d['ms_per_step'],2))"; }
for K in 16 17 18; do
  for C in $((K-3)) $((K-2)) $((K-1)); do
    ZK_MSM_C=$C run C=$C
  done
done
