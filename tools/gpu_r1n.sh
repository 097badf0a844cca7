#!/bin/bash
OUT=gpurun_out
run() { echo "== TUN=$1 OPT=$2 batch=$3" | tee -a $OUT/decode_ab_r1n.log
  FTCF_TUNABLES=$1 FTCF_OPTIONS=$2 timeout 300 python tools/profile_decode.py --batch $3 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1n.log; }
for b in 1 2; do
run pdl=1 pro_ctas=148 $b
run pdl=1 pro_ctas=0 $b
run pdl=1 pro_ctas=120 $b
run pdl=1 pro_ctas=200 $b
done
run pdl=1 pro_ctas=148 4
run pdl=1 pro_ctas=0 4
run pdl=1 pro_ctas=0,fused_ln=0 4
FTCF_OPTIONS=pro_ctas=148 timeout 300 python tools/trace_step.py > $OUT/trace_r1n.log 2>&1; tail -30 $OUT/trace_r1n.log
