#!/bin/bash
# A/B builds of the library (zk-dcap-verifier_b200/build/variants/*.so, selected with ZKDCAP_B200_LIB): NTT pass unroll / resident
# CTAs, and the multiplier without the dedicated squaring.  Objects are linked against the CURRENT sources.
set -e
P=/root/repo/zk-dcap-verifier_b200
V=$P/build/variants
mkdir -p $V
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
L="-gencode arch=compute_100a,code=sm_100a -cudart static -ldl"
BASE="$P/build/api.o $P/build/msm.o $P/build/evalh.o $P/build/prover.o $P/build/group.o $P/build/params.o $P/build/verifier.o"
for v in "2 2" "2 3" "1 6"; do
  set -- $v
  nvcc $F -DZK_NTT_UNROLL=$1 -DZK_NTT_MIN_CTAS=$2 -c $P/csrc/ntt.cu -o $V/ntt_u$1c$2.o
  nvcc -shared -o $V/libzk_ntt_u$1c$2.so $BASE $V/ntt_u$1c$2.o $L
done
for f in api ntt msm evalh prover group params verifier; do nvcc $F -DZK_NO_DEDICATED_SQR -c $P/csrc/$f.cu -o $V/nosqr_$f.o & done
wait
nvcc -shared -o $V/libzk_nosqr.so $V/nosqr_api.o $V/nosqr_ntt.o $V/nosqr_msm.o $V/nosqr_evalh.o $V/nosqr_prover.o $V/nosqr_group.o $V/nosqr_params.o $V/nosqr_verifier.o $L
ls -la $V/*.so | awk '{print $5, $9}'
