#!/bin/bash
# usage: gpu_tpn.sh N  -- TP parity + bench at N GPUs with both gather variants + one detailed timeline
N=${1:-4}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_tp_gpu.py -q -x -k "parity" 2>&1 | tail -3
for opt in "" "tp_gather_kernel=0" "tp_fused=0"; do
  echo "### N=$N FTCF_OPTIONS=$opt"
  FTCF_OPTIONS="$opt" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --skip-extra --skip-cpu 2> $OUT/tp${N}_bench.err | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['decode']['p50_token_ms'], d['decode']['prefill_ms'])"
done | tee $OUT/tp${N}_bench.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/trace_step.py --show 1 --detail 1 > $OUT/tp${N}_timeline.txt 2>&1; tail -40 $OUT/tp${N}_timeline.txt | head -30
