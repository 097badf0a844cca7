#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tp_gpu.py -q -x -k "parity" 2>&1 | tail -2
for tun in "" "decode_cluster=0" "decode_max_stages=5"; do
  echo "### FTCF_TUNABLES=$tun"
  FTCF_TUNABLES="$tun" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --skip-extra --skip-cpu 2> $OUT/tp2_bench.err | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['decode']['p50_token_ms'], d['decode']['prefill_ms'])"
done | tee $OUT/tp2_bench_c.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/trace_step.py --show 1 > $OUT/tp2_timeline_c.txt 2>&1; tail -40 $OUT/tp2_timeline_c.txt | head -14
