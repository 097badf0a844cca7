#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py tests/test_beam_search_gpu.py tests/test_sampling_gpu.py -q -x 2>&1 | tail -4
printf "%s\n" "||1" "|o_by_head=0|1" "||1" "|o_by_head=0|1" "||4" "|o_by_head=0|4" "|o_ctas=120|1" "|o_ctas=280|1" | bash tools/decode_ab.sh | tee $OUT/r2w_ab.txt
timeout 600 python tools/trace_step.py --show 1 --detail 1 > $OUT/r2w_timeline.txt 2>&1
grep -v "end deciles" $OUT/r2w_timeline.txt | sed -n 11,16p | cut -c1-150
