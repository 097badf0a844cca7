#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gemm_decode_gpu.py tests/test_kernels_gpu.py tests/test_baseline_shapes_gpu.py -q -x > $OUT/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/r2e_pytest.log
timeout 600 python tools/bench_gemm_chain.py --impl 3 --m 1 2>&1 | grep "pdl=1" > $OUT/r2e_chain.log; cat $OUT/r2e_chain.log
timeout 600 python tools/trace_step.py --show 2 > $OUT/r2e_timeline.txt 2>&1; tail -34 $OUT/r2e_timeline.txt
bash tools/decode_ab.sh > $OUT/r2e_ab.txt 2>&1 <<'EOT'
||1
mmha_bulk=0||1
decode_impl=1||1
decode_max_stages=3||1
|kv_prefetch=1|1
decode_max_stages=3|kv_prefetch=1|1
|pro_ctas=0|1
||8
decode_impl=1||8
||32
decode_impl=1||32
EOT
cat $OUT/r2e_ab.txt
