#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gemm_decode_gpu.py tests/test_model_gpu.py tests/test_ref_kernels_gpu.py -q -x 2>&1 | tail -3
FTCF_OPTIONS="ffn2_ctas=160" timeout 600 python tools/trace_step.py --show 1 > $OUT/r2n_timeline_a.txt 2>&1; tail -30 $OUT/r2n_timeline_a.txt | head -12
bash tools/decode_ab.sh > $OUT/r2n_ab.txt 2>&1 <<'EOT'
|ffn2_ctas=160|1
decode_max_stages=4|ffn2_ctas=160|1
|ffn2_after_attn=1|1
decode_impl=1||1
||8
||32
EOT
cat $OUT/r2n_ab.txt
