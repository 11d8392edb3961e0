#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_inflight_stress.py 150 2 19 > gpurun_out/r2_inflight_stress.log 2>&1
tail -n 12 gpurun_out/r2_inflight_stress.log
ZK_EVALH=interp python tools/gpu_inflight_stress.py 100 2 19 > gpurun_out/r2_inflight_stress_interp.log 2>&1
tail -n 5 gpurun_out/r2_inflight_stress_interp.log
