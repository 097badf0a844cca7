#!/bin/bash
# Final evidence run of a round: parity tests, the bench line, ncu launch list of decode steps, ncu --set full of the decode GEMMs + attention.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python tools/profile_decode.py --in-len 1024 --out-len 4 > $OUT/prof_list_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 1 -c 4 -f -o $OUT/skinny_$TAG \
    python tools/profile_decode.py --out-len 3 --layers 4 > $OUT/prof_skinny_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmha_decode -s 2 -c 2 -f -o $OUT/mmha_$TAG \
    python tools/profile_decode.py --out-len 3 --layers 4 > $OUT/prof_mmha_$TAG.log 2>&1
ls -la $OUT | tail -8
