#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_r2.log
tail -12 $OUT/pytest_gpu_r2.log | cut -c1-300
timeout 900 python bench.py --steps 2 --warmup 3 --skip-cpu > $OUT/bench_r2q.json 2> $OUT/bench_r2q.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2q.json').read().strip().splitlines()[-1])
print("value", d["value"], "p50", d["decode"]["p50_token_ms"], "step_frac", d["decode"]["step_roofline_frac"], "prefill", d["decode"]["prefill_ms"])
print("roofline", d["roofline"]["frac"], d["roofline"]["avg_launch_us"], "batch32", d["roofline_batch32"]["frac"], d["roofline_batch32"]["avg_launch_us"])
print("config5", d["config5_batch32_2048_512"]["tokens_per_s"], d["config5_batch32_2048_512"]["p50_token_ms"], "config2", d["config2_fp16"]["tokens_per_s"])
PY
