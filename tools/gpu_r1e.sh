#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1e.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1e.log
tail -15 $OUT/pytest_gpu_r1e.log
for prio in 1 0 -1; do
  for opts in "fused_ln=1" "fused_ln=0"; do
    echo "== FTCF_STREAM_PRIO=$prio FTCF_OPTIONS=$opts batch=1" | tee -a $OUT/decode_ab_r1e.log
    FTCF_STREAM_PRIO=$prio FTCF_OPTIONS=$opts timeout 300 python tools/profile_decode.py --batch 1 --out-len 129 --requests 3 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1e.log
  done
done
echo "== batch 8 / 32" | tee -a $OUT/decode_ab_r1e.log
timeout 300 python tools/profile_decode.py --batch 8 --out-len 129 --requests 2 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1e.log
timeout 300 python tools/profile_decode.py --batch 32 --out-len 65 --requests 2 --graph 1 2>&1 | tail -1 | tee -a $OUT/decode_ab_r1e.log
timeout 300 python tools/trace_step.py > $OUT/trace_r1e.log 2>&1; tail -40 $OUT/trace_r1e.log
LD_LIBRARY_PATH=fastertransformer4codefuse_b200/lib timeout 300 tools/gcb.bin 1 > $OUT/gcb_r1e.log 2>&1; cat $OUT/gcb_r1e.log
