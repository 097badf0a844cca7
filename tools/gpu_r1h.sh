#!/bin/bash
OUT=gpurun_out
run() { echo "== PRIO=$1 TUN=$2 OPT=$3 batch=$4" | tee -a $OUT/decode_ab_r1h.log
  FTCF_STREAM_PRIO=$1 FTCF_TUNABLES=$2 FTCF_OPTIONS=$3 timeout 300 python tools/profile_decode.py --batch $4 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1h.log; }
run 1 mmha_pdl=1 fused_ln=1 1
run 0 mmha_pdl=1 fused_ln=1 1
run -1 mmha_pdl=1 fused_ln=1 1
run 1 mmha_pdl=0 fused_ln=1 1
run 0 mmha_pdl=0 fused_ln=1 1
run -1 mmha_pdl=0 fused_ln=1 1
run 0 mmha_pdl=0 fused_ln=0 1
run 0 mmha_pdl=1 fused_ln=0 1
run 0 mmha_pdl=0 fused_ln=1 2
run 0 mmha_pdl=0 fused_ln=0 2
run 0 mmha_pdl=0 fused_ln=0 8
run 0 mmha_pdl=1 fused_ln=0 8
