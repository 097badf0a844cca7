#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
FTCF_OPTIONS="tp_gather_kernel=1" timeout 600 python -m pytest tests/test_tp_gpu.py -q -x -k "parity" 2>&1 | tail -3
for opt in "" "tp_gather_kernel=1,tp_gather_both=0" "tp_gather_kernel=1,tp_gather_both=1"; do
  echo "### FTCF_OPTIONS=$opt"
  FTCF_OPTIONS="$opt" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --skip-extra --skip-cpu 2> $OUT/tp2_bench.err | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['decode']['p50_token_ms'], d['decode']['prefill_ms'])"
done | tee $OUT/tp2_bench_e.txt
FTCF_OPTIONS="tp_gather_kernel=1,tp_gather_both=0" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/trace_step.py --show 1 --detail 1 > $OUT/tp2_timeline_f0.txt 2>&1
FTCF_OPTIONS="tp_gather_kernel=1,tp_gather_both=1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 tools/trace_step.py --show 1 --detail 1 > $OUT/tp2_timeline_f1.txt 2>&1
grep -v "end deciles" $OUT/tp2_timeline_f1.txt | sed -n '8,26p' | cut -c1-150
