#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tp_gpu.py -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --skip-cpu 2> $OUT/tp2_bench.err | tail -1 > $OUT/r2_tp2_bench.json
python -c "
import json; d = json.loads(open('$OUT/r2_tp2_bench.json').read())
print('N=2', d['value'], d['decode']['p50_token_ms'], d['decode']['prefill_ms'], 'config5', d.get('config5_batch32_2048_512', {}).get('tokens_per_s'), 'config4', d.get('config4_batch8', {}).get('tokens_per_s'))"
