#!/bin/bash
OUT=gpurun_out
run() { echo "== TUN=$1 OPT=$2 batch=$3" | tee -a $OUT/decode_ab_r1p.log
  FTCF_TUNABLES=$1 FTCF_OPTIONS=$2 timeout 300 python tools/profile_decode.py --batch $3 --out-len 65 --requests 2 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1p.log; }
for b in 32 16; do
run pdl=1 gemm_impl=0 $b
run pdl=1 gemm_impl=2 $b
run pdl=1 gemm_impl=0,two_branch=0 $b
run pdl=1 gemm_impl=2,two_branch=0 $b
done
