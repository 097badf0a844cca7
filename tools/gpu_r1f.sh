#!/bin/bash
OUT=gpurun_out
for opts in "two_branch=2" "two_branch=1" "two_branch=0"; do
    echo "== FTCF_OPTIONS=$opts batch=1" | tee -a $OUT/decode_ab_r1f.log
    FTCF_OPTIONS=$opts timeout 300 python tools/profile_decode.py --batch 1 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1f.log
done
FTCF_OPTIONS=two_branch=2 timeout 300 python tools/trace_step.py > $OUT/trace_r1f.log 2>&1; tail -32 $OUT/trace_r1f.log
