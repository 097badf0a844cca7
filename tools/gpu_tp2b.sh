#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/trace_step.py --show 1 > $OUT/tp2_timeline_b.txt 2>&1; tail -40 $OUT/tp2_timeline_b.txt | head -16
timeout 600 python -m pytest tests/test_tp_gpu.py -q -x -k "parity" 2>&1 | tail -2
for opt in "" "o_ctas=40" "ffn2_ctas=120,o_ctas=40" "tp_fused=0"; do
  echo "### FTCF_OPTIONS=$opt"
  FTCF_OPTIONS="$opt" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --skip-extra --skip-cpu 2> $OUT/tp2_bench.err | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['decode']['p50_token_ms'], d['decode']['prefill_ms'])"
done | tee $OUT/tp2_bench_b.txt
