#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
{
for tun in "" "decode_max_stages=6" "decode_max_stages=10" "decode_lean=1" "decode_lean=1,decode_max_stages=6"; do
  echo "### tun=$tun"
  timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 1 --tun "$tun" 2>&1 | tail -5
done
echo "### skinny"
timeout 300 python tools/bench_gemm_chain.py --impl 1 --m 1 2>&1 | tail -4
} | tee $OUT/r2s_chain.txt
