#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_ref_kernels_gpu.py tests/test_baseline_shapes_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -3
FTCF_OPTIONS="ffn2_ctas=160" timeout 600 python tools/trace_step.py --show 2 > $OUT/r2m_timeline_a.txt 2>&1; tail -34 $OUT/r2m_timeline_a.txt | head -18
FTCF_OPTIONS="ffn2_after_attn=1" timeout 600 python tools/trace_step.py --show 2 > $OUT/r2m_timeline_b.txt 2>&1; tail -34 $OUT/r2m_timeline_b.txt | head -18
bash tools/decode_ab.sh > $OUT/r2m_ab.txt 2>&1 <<'EOT'
|ffn2_ctas=160|1
decode_max_stages=4|ffn2_ctas=160|1
|ffn2_after_attn=1|1
|ffn2_after_attn=1,ffn2_ctas=160|1
|ffn2_after_attn=1,ffn2_ctas=160,o_ctas=120|1
decode_max_stages=4|ffn2_after_attn=1|1
decode_impl=1||1
EOT
cat $OUT/r2m_ab.txt
