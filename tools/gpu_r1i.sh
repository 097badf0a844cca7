#!/bin/bash
OUT=gpurun_out
run() { echo "== TUN=$1 OPT=$2 batch=$3" | tee -a $OUT/decode_ab_r1i.log
  FTCF_STREAM_PRIO=0 FTCF_TUNABLES=$1 FTCF_OPTIONS=$2 timeout 300 python tools/profile_decode.py --batch $3 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1i.log; }
for b in 1 8; do
run mmha_pdl=0,mmha_prefetch=1,skinny_carveout=1 fused_ln=1 $b
run mmha_pdl=0,mmha_prefetch=0,skinny_carveout=1 fused_ln=1 $b
run mmha_pdl=0,mmha_prefetch=1,skinny_carveout=0 fused_ln=1 $b
run mmha_pdl=0,mmha_prefetch=0,skinny_carveout=0 fused_ln=1 $b
run mmha_pdl=1,mmha_prefetch=1,skinny_carveout=0 fused_ln=1 $b
run mmha_pdl=1,mmha_prefetch=0,skinny_carveout=0 fused_ln=1 $b
done
run mmha_pdl=0,mmha_prefetch=0,skinny_carveout=0 fused_ln=0 1
