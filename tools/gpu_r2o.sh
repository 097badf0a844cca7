#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
FTCF_OPTIONS="qkv_first=1,qkv_ctas=296,ffn1_ctas=160,ffn2_ctas=160,ffn2_no_pdl=1" timeout 600 python tools/trace_step.py --show 1 > $OUT/r2o_timeline_a.txt 2>&1; tail -30 $OUT/r2o_timeline_a.txt | head -12
bash tools/decode_ab.sh > $OUT/r2o_ab.txt 2>&1 <<'EOT'
|qkv_first=1,qkv_ctas=296,ffn1_ctas=160,ffn2_ctas=160,ffn2_no_pdl=1|1
|qkv_first=1,qkv_ctas=296,ffn1_ctas=160,ffn2_ctas=160|1
|qkv_first=1,qkv_ctas=296,ffn1_ctas=148,ffn2_ctas=148,ffn2_no_pdl=1|1
|qkv_first=1,qkv_ctas=240,ffn1_ctas=160,ffn2_ctas=160,ffn2_no_pdl=1,o_ctas=120|1
|qkv_first=1,qkv_ctas=444,ffn1_ctas=160,ffn2_ctas=160,ffn2_no_pdl=1|1
|qkv_first=1,qkv_ctas=296,ffn1_ctas=296,ffn2_ctas=296|1
|ffn2_ctas=160|1
EOT
cat $OUT/r2o_ab.txt
