#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gemm_decode_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 1 2>&1 | grep "tcgen05" | head -3
timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 1 --tun decode_cluster=0 2>&1 | grep "tcgen05" | head -2
bash tools/decode_ab.sh > $OUT/r2p_ab.txt 2>&1 <<'EOT'
||1
decode_cluster=0||1
|o_ctas=120|1
|ffn2_ctas=280|1
decode_impl=1||1
EOT
cat $OUT/r2p_ab.txt
timeout 600 python tools/trace_step.py --show 1 > $OUT/r2p_timeline.txt 2>&1; tail -30 $OUT/r2p_timeline.txt | head -12
