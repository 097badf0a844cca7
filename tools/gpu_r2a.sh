#!/bin/bash
# round-2 first GPU session: all gpu tests (no -x: every failure listed), bench, ncu of the prefill tcgen05 GEMM.
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_tp_gpu.py > $OUT/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2a_pytest.log
tail -30 $OUT/r2a_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r2a_bench.json 2> $OUT/r2a_bench.err; echo "bench rc=$?"
cat $OUT/r2a_bench.json; tail -5 $OUT/r2a_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 2 -f -o $OUT/r2a_tc_prefill \
    python tools/profile_decode.py --out-len 2 --layers 3 --in-len 1024 > $OUT/r2a_prof_tc.log 2>&1
tail -3 $OUT/r2a_prof_tc.log
ls -la $OUT | tail -8
