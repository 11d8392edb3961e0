#!/bin/bash
# window-width probe for large single resident MSMs (parity checked by the sweep itself)
mkdir -p gpurun_out
: > gpurun_out/r2_msm_window_probe.txt
for k in 21 22 23 24 25 26; do
  for c in 16 20 22 24; do
    ZK_MSM_C=$c timeout 300 python bench.py --mode sweep --sweep-min $k --sweep-max $k --sweep-coset-max 0 2>> gpurun_out/r2_probe.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('k', d['k'], 'c', d['window_bits'], 'W', d['windows'], round(d['msm_resident_ms'],2), 'ms', round(d['msm_resident_Gpts_s'],3), 'Gpts/s', d['msm_resident_parity'][:5])" | tee -a gpurun_out/r2_msm_window_probe.txt
  done
done
tail -n 3 gpurun_out/r2_probe.err
