#!/bin/bash
# fresh provers next to an established one: does the first proof of a new prover ever differ?
mkdir -p gpurun_out
timeout 240 python tools/gpu_fresh_prover_stress.py 110 2 19 0 > gpurun_out/r2_fresh_prover_stress.log 2>&1
tail -n 12 gpurun_out/r2_fresh_prover_stress.log
