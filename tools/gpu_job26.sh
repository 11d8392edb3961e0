#!/bin/bash
# targeted GPU tests of the late-round additions (Poseidon transcript through the device prover, key read from pk bytes)
mkdir -p gpurun_out
(time python -m pytest tests/test_prover_gpu.py tests/test_keyfile.py -m gpu -q -x -k "poseidon or evm or proving_key_read or proof_bytes_match") > gpurun_out/r2_gputest_poseidon_keyfile.log 2>&1
tail -n 6 gpurun_out/r2_gputest_poseidon_keyfile.log
