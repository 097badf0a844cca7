#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1m.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1m.log
tail -5 $OUT/pytest_gpu_r1m.log
run() { echo "== TUN=$1 OPT=$2 batch=$3" | tee -a $OUT/decode_ab_r1m.log
  FTCF_TUNABLES=$1 FTCF_OPTIONS=$2 timeout 300 python tools/profile_decode.py --batch $3 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1m.log; }
for b in 1 2; do
run skinny_evict_first=1 kv_prefetch=1 $b
run skinny_evict_first=0 kv_prefetch=1 $b
run skinny_evict_first=1 kv_prefetch=0 $b
run skinny_evict_first=0 kv_prefetch=0 $b
done
FTCF_OPTIONS=kv_prefetch=1 timeout 300 python tools/trace_step.py > $OUT/trace_r1m.log 2>&1; tail -30 $OUT/trace_r1m.log
