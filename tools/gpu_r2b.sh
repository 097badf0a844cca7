#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gemm_decode_gpu.py -x -q > $OUT/r2b_pytest_decode.log 2>&1; rc=$?; echo "decode pytest rc=$rc"
tail -25 $OUT/r2b_pytest_decode.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python tools/bench_gemm_chain.py --impl 3 --m 1 > $OUT/r2b_chain_tc.log 2>&1; cat $OUT/r2b_chain_tc.log
timeout 300 python tools/bench_gemm_chain.py --impl 1 --m 1 > $OUT/r2b_chain_sk.log 2>&1; cat $OUT/r2b_chain_sk.log
timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 8 32 > $OUT/r2b_chain_tc_m.log 2>&1; grep "pdl=1" $OUT/r2b_chain_tc_m.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_tp_gpu.py > $OUT/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/r2b_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --skip-cpu > $OUT/r2b_bench.json 2> $OUT/r2b_bench.err; echo "bench rc=$?"
cat $OUT/r2b_bench.json; tail -5 $OUT/r2b_bench.err
