#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
FTCF_TUNABLES=mmha_lite=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_ref_kernels_gpu.py tests/test_model_gpu.py -q -x -k "mmha or greedy or ragged or model or attention" 2>&1 | tail -4
printf "%s\n" "mmha_lite=1||1" "mmha_lite=0||1" "mmha_lite=1|ffn2_ctas=280|1" "mmha_lite=1|ffn2_ctas=200|1" "mmha_lite=1|ffn2_ctas=148|1" "mmha_lite=1,decode_max_stages=5||1" "mmha_lite=1||4" "mmha_lite=0||4" | bash tools/decode_ab.sh | tee $OUT/r2u_ab.txt
FTCF_TUNABLES=mmha_lite=1 timeout 600 python tools/trace_step.py --show 1 --detail 1 > $OUT/r2u_timeline_lite.txt 2>&1
grep -A1 "^mmha\|^gemm_w8     5120   5120" $OUT/r2u_timeline_lite.txt | sed -n 5,10p | cut -c1-170
