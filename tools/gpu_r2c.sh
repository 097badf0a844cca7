#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/trace_step.py --show 2 > $OUT/r2c_timeline_tc.txt 2>&1; tail -40 $OUT/r2c_timeline_tc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_decode -c 12 -f -o $OUT/r2c_decode_gemm \
    python tools/run_decode_gemms.py 1 3 3 > $OUT/r2c_ncu.log 2>&1; tail -3 $OUT/r2c_ncu.log
