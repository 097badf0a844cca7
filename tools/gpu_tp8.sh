#!/bin/bash
# 8-GPU box: TP parity at 8 ranks, bench at N = 8 and 4 (+ the NCCL path at 8 for comparison), one detailed timeline at 8
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_tp_gpu.py -q -x -k "8" 2>&1 | tail -3
run() {  # N opts
  FTCF_OPTIONS="$2" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $1 --steps 2 --warmup 3 --skip-extra --skip-cpu 2> $OUT/tp$1_bench.err | tail -1 > $OUT/tp$1_bench_line.json
  python -c "import sys, json; d = json.loads(open('$OUT/tp$1_bench_line.json').read()); print('N=$1 [$2]', d['value'], d['decode']['p50_token_ms'], d['decode']['prefill_ms'])"
}
{ run 8 ""; cp $OUT/tp8_bench_line.json $OUT/tp8_bench_default.json; run 4 ""; run 8 "tp_fused=0"; } | tee $OUT/tp8_bench.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/trace_step.py --show 1 --detail 1 > $OUT/tp8_timeline.txt 2>&1
grep -v "end deciles" $OUT/tp8_timeline.txt | sed -n '8,30p' | cut -c1-150
