#!/bin/bash
# Final evidence run of a round on ONE GPU: parity tests, the bench line (both arms), ncu launch list of decode steps,
# ncu --set full of the four decode GEMMs (tcgen05 kernel), the decode attention and the prefill GEMM.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python tools/profile_decode.py --in-len 1024 --out-len 4 > $OUT/prof_list_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_decode -c 4 -f -o $OUT/${TAG}_gemm_decode \
    python tools/run_decode_gemms.py 1 3 1 > $OUT/prof_gemm_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmha_decode -s 2 -c 2 -f -o $OUT/${TAG}_mmha \
    python tools/profile_decode.py --out-len 3 --layers 4 > $OUT/prof_mmha_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 2 -f -o $OUT/${TAG}_gemm_prefill \
    python tools/profile_decode.py --out-len 2 --layers 2 > $OUT/prof_prefill_$TAG.log 2>&1
ls -la $OUT | tail -12
