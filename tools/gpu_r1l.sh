#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1l.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1l.log
tail -12 $OUT/pytest_gpu_r1l.log
run() { echo "== TUN=$1 OPT=$2 batch=$3" | tee -a $OUT/decode_ab_r1l.log
  FTCF_TUNABLES=$1 FTCF_OPTIONS=$2 timeout 300 python tools/profile_decode.py --batch $3 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1l.log; }
for b in 1 8; do
run skinny_ksplit=1 fused_ln=1 $b
run skinny_ksplit=0 fused_ln=1 $b
run skinny_ksplit=1,skinny_target_ctas=296 fused_ln=1 $b
run skinny_ksplit=0,skinny_target_ctas=296 fused_ln=1 $b
run skinny_ksplit=1 fused_ln=0 $b
done
LD_LIBRARY_PATH=fastertransformer4codefuse_b200/lib timeout 300 tools/gcb.bin 1 > $OUT/gcb_r1l.log 2>&1; cat $OUT/gcb_r1l.log
