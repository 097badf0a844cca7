#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1g.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1g.log
tail -5 $OUT/pytest_gpu_r1g.log
for opts in "two_branch=1" "two_branch=1,fused_ln=0" "two_branch=2"; do
  for b in 1 4 8; do
    echo "== FTCF_OPTIONS=$opts batch=$b" | tee -a $OUT/decode_ab_r1g.log
    FTCF_OPTIONS=$opts timeout 300 python tools/profile_decode.py --batch $b --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1g.log
  done
done
FTCF_OPTIONS=two_branch=1 timeout 300 python tools/trace_step.py > $OUT/trace_r1g.log 2>&1; tail -32 $OUT/trace_r1g.log
