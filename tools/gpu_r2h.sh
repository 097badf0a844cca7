#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
(cd tools && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../fastertransformer4codefuse_b200/csrc -I../include umma_probe.cu -o /tmp/umma_probe && timeout 60 /tmp/umma_probe) | tee $OUT/r2h_umma_probe.txt
timeout 900 python -m pytest tests/test_gemm_decode_gpu.py tests/test_kernels_gpu.py tests/test_baseline_shapes_gpu.py -q -x > $OUT/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/r2h_pytest.log
for tun in "" "decode_max_stages=3"; do
  echo "### tun=$tun"; timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 1 --tun "$tun" 2>&1 | grep tcgen05
done | tee $OUT/r2h_chain.log
timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 8 32 2>&1 | grep "target_ctas=296 min_kb=8" | tee -a $OUT/r2h_chain.log
timeout 300 python tools/bench_gemm_chain.py --impl 1 --m 1 2>&1 | grep "pdl=1" | tee -a $OUT/r2h_chain.log
timeout 600 python tools/trace_step.py --show 2 > $OUT/r2h_timeline.txt 2>&1; tail -34 $OUT/r2h_timeline.txt
bash tools/decode_ab.sh > $OUT/r2h_ab.txt 2>&1 <<'EOT'
||1
decode_max_stages=3||1
decode_impl=1||1
|pro_ctas=0|1
||8
decode_impl=1||8
||32
decode_impl=1||32
EOT
cat $OUT/r2h_ab.txt
