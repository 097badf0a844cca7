#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
FTCF_OPTIONS="ffn2_ctas=148,o_ctas=148" timeout 600 python tools/trace_step.py --show 2 > $OUT/r2k_timeline_a.txt 2>&1; tail -34 $OUT/r2k_timeline_a.txt | head -20
bash tools/decode_ab.sh > $OUT/r2k_ab.txt 2>&1 <<'EOT'
|ffn2_ctas=148|1
|ffn2_ctas=120|1
|ffn2_ctas=148,o_ctas=148|1
|ffn2_ctas=148,o_ctas=120|1
|ffn2_ctas=148,o_ctas=80|1
|ffn2_ctas=148,pro_ctas=0|1
|ffn2_ctas=148,qkv_ctas=120,ffn1_ctas=160|1
|ffn2_ctas=148,ffn2_no_pdl=1|1
decode_min_kb=4|ffn2_ctas=148|1
mmha_pdl=1|ffn2_ctas=148|1
|ffn2_ctas=148,kv_prefetch=1|1
decode_impl=1||1
EOT
cat $OUT/r2k_ab.txt
