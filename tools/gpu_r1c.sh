#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1c.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1c.log
tail -15 $OUT/pytest_gpu_r1c.log
for opts in "fused_ln=1" "fused_ln=0"; do
  for b in 1 4; do
    echo "== FTCF_OPTIONS=$opts batch=$b" | tee -a $OUT/decode_ab_r1c.log
    FTCF_OPTIONS=$opts timeout 300 python tools/profile_decode.py --batch $b --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1c.log
  done
done
for t in "skinny_pf_ahead=8" "skinny_pf_ahead=32" "skinny_target_ctas=444"; do
    echo "== FTCF_TUNABLES=$t batch=1" | tee -a $OUT/decode_ab_r1c.log
    FTCF_TUNABLES=$t timeout 300 python tools/profile_decode.py --batch 1 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1c.log
done
