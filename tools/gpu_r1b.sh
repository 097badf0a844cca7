#!/bin/bash
# GPU session: parity tests, then decode-step timing of the persistent kernel vs the multi-kernel graph.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_r1b.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1b.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1b.log
tail -3 $OUT/pytest_gpu_r1b.log
for opts in "mega=1" "mega=0" "mega=0,two_branch=0"; do
  for b in 1 8; do
    echo "== FTCF_OPTIONS=$opts batch=$b" | tee -a $OUT/decode_ab_r1b.log
    FTCF_OPTIONS=$opts timeout 300 python tools/profile_decode.py --batch $b --out-len 129 --requests 3 --graph 1 2>&1 | tail -2 | tee -a $OUT/decode_ab_r1b.log
  done
done
timeout 300 python tools/mega_barrier_trace.py > $OUT/mega_trace_r1b.log 2>&1; tail -40 $OUT/mega_trace_r1b.log
