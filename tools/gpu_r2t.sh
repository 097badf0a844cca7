#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for L in 256 512 1024; do
  timeout 600 python tools/trace_step.py --show 1 --detail 1 --in-len $L > $OUT/r2t_timeline_$L.txt 2>&1
  echo "=== in-len $L"; grep -A1 "^mmha\|^gemm_w8     5120   5120" $OUT/r2t_timeline_$L.txt | sed -n 5,12p | cut -c1-170
done
