#!/bin/bash
# two GPUs: tensor-parallel parity (fused exchange on) and the bench line with / without the fused exchange
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_tp_gpu.py -q -x -k "parity" > $OUT/tp2_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $OUT/tp2_pytest.log
for opt in "" "tp_fused=0"; do
  echo "### FTCF_OPTIONS=$opt"
  FTCF_OPTIONS="$opt" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --skip-extra --skip-cpu 2> $OUT/tp2_bench.err | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(json.dumps({k: d[k] for k in ('value', 'n_gpus', 'decode', 'e2e')}))"
done | tee $OUT/tp2_bench.txt
tail -5 $OUT/tp2_bench.err
