#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list + full captures of the two dominant decode kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_ncu]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
if [ -z "$2" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
      python tools/profile_decode.py --out-len 3 > $OUT/prof_list_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 40 -c 4 -f -o $OUT/skinny_$TAG \
      python tools/profile_decode.py --out-len 3 --layers 4 > $OUT/prof_skinny_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:mmha_decode -s 4 -c 2 -f -o $OUT/mmha_$TAG \
      python tools/profile_decode.py --out-len 3 --layers 4 > $OUT/prof_mmha_$TAG.log 2>&1
fi
ls -la $OUT | tail -12
