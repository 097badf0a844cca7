#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
printf '%s\n' "||1" "decode_lean=1||1" "decode_lean=1|ffn2_ctas=280|1" "decode_lean=1|ffn2_ctas=148|1" "decode_lean=1,decode_max_stages=3||1" "decode_impl=1||1" "||8" "decode_lean=1||8" | bash tools/decode_ab.sh | tee $OUT/r2r_ab.txt
FTCF_TUNABLES="decode_lean=1" timeout 600 python tools/trace_step.py --show 1 --detail 1 > $OUT/r2r_timeline_lean.txt 2>&1
grep -v "end deciles" $OUT/r2r_timeline_lean.txt | sed -n 8,24p | cut -c1-150
