#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gemm_decode_gpu.py tests/test_kernels_gpu.py -q -x > $OUT/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/r2j_pytest.log
python tools/decode_gemm_probe.py 20480 5120 1 2>&1 | head -14 | tee $OUT/r2j_probe.txt
for tun in "" "decode_max_stages=3"; do
  echo "### tun=$tun"; timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 1 --tun "$tun" 2>&1 | grep tcgen05
done | tee $OUT/r2j_chain.log
bash tools/decode_ab.sh > $OUT/r2j_ab.txt 2>&1 <<'EOT'
||1
decode_max_stages=3||1
decode_max_stages=3|ffn2_no_pdl=1|1
decode_max_stages=3|ffn2_ctas=160|1
decode_max_stages=3|ffn2_ctas=160,o_ctas=160|1
decode_max_stages=3|qkv_ctas=240,ffn1_ctas=160|1
decode_max_stages=3|qkv_ctas=240,ffn1_ctas=160,ffn2_ctas=160|1
decode_max_stages=4|ffn2_ctas=160|1
|ffn2_ctas=160|1
decode_impl=1||1
||8
||32
EOT
cat $OUT/r2j_ab.txt
timeout 600 python tools/trace_step.py --show 2 > $OUT/r2j_timeline.txt 2>&1; tail -34 $OUT/r2j_timeline.txt | head -22
