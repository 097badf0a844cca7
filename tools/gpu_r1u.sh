#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r1u.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r1u.log
tail -12 $OUT/pytest_gpu_r1u.log
for b in 1 8 32; do for t in 1 0; do echo "== batch=$b mmha_onepass=$t"; FTCF_TUNABLES=mmha_onepass=$t timeout 300 python tools/profile_decode.py --batch $b --out-len 65 --requests 2 --graph 1 2>&1 | tail -1; done; done
