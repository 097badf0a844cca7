#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "launch_hints" 2>&1 | tail -3
FTCF_OPTIONS="ffn2_ctas=148,ffn2_stages=6" timeout 600 python tools/trace_step.py --show 2 > $OUT/r2l_timeline_a.txt 2>&1; tail -34 $OUT/r2l_timeline_a.txt | head -20
bash tools/decode_ab.sh > $OUT/r2l_ab.txt 2>&1 <<'EOT'
|ffn2_ctas=148,ffn2_stages=6|1
|ffn2_ctas=148,ffn2_stages=8|1
|ffn2_ctas=148,ffn2_stages=10|1
|ffn2_ctas=120,ffn2_stages=8|1
|ffn2_ctas=148,ffn2_stages=8,o_ctas=148|1
|ffn2_ctas=148,ffn2_stages=8,o_ctas=148,o_stages=3|1
|ffn2_ctas=148,ffn2_stages=6,o_ctas=200,o_stages=3|1
decode_impl=1||1
EOT
cat $OUT/r2l_ab.txt
