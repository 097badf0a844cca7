#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1o.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1o.log
tail -12 $OUT/pytest_gpu_r1o.log
run() { echo "== TUN=$1 OPT=$2 batch=$3" | tee -a $OUT/decode_ab_r1o.log
  FTCF_TUNABLES=$1 FTCF_OPTIONS=$2 timeout 300 python tools/profile_decode.py --batch $3 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1o.log; }
for b in 1 8; do
run skinny_even_rows=1 pro_ctas=148 $b
run skinny_even_rows=0 pro_ctas=148 $b
run skinny_even_rows=1 pro_ctas=0 $b
done
run skinny_even_rows=1 pro_ctas=148 32
run skinny_even_rows=0 pro_ctas=148 32
FTCF_OPTIONS=pro_ctas=148 timeout 300 python tools/trace_step.py > $OUT/trace_r1o.log 2>&1; tail -30 $OUT/trace_r1o.log
LD_LIBRARY_PATH=fastertransformer4codefuse_b200/lib timeout 300 tools/gcb.bin 1 2>&1 | head -4
