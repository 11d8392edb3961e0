#!/bin/bash
# in-flight provers under host CPU contention (what an 8-rank bench run looks like to each rank)
mkdir -p gpurun_out
for i in $(seq 1 24); do (python -c "
import time
t=time.time()
x=0
while time.time()-t < 170: x+=1
" &) ; done
python tools/gpu_inflight_stress.py 250 3 19 > gpurun_out/r2_inflight_stress_loaded.log 2>&1
tail -n 12 gpurun_out/r2_inflight_stress_loaded.log
