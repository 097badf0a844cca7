#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_ref_kernels_gpu.py tests/test_model_gpu.py tests/test_baseline_shapes_gpu.py -q -x 2>&1 | tail -4
FTCF_TUNABLES=mmha_lite=0 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_ref_kernels_gpu.py -q -x -k "mmha or attention" 2>&1 | tail -2
printf "%s\n" "||1" "mmha_lite=0||1" "||1" "mmha_lite=0||1" "||2" "mmha_lite=0||2" | bash tools/decode_ab.sh | tee $OUT/r2v_ab.txt
timeout 600 python tools/trace_step.py --show 1 --detail 1 > $OUT/r2v_timeline.txt 2>&1
grep -A1 "^mmha" $OUT/r2v_timeline.txt | sed -n 4,6p | cut -c1-170
