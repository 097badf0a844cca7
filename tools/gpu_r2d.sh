#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_tp_gpu.py > $OUT/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/r2d_pytest.log
timeout 600 python tools/trace_step.py --show 2 > $OUT/r2d_timeline.txt 2>&1; tail -34 $OUT/r2d_timeline.txt
bash tools/decode_ab.sh > $OUT/r2d_ab.txt 2>&1 <<'EOT'
||1
mmha_bulk=0||1
decode_impl=1||1
decode_impl=1,mmha_bulk=0||1
decode_max_stages=3||1
decode_max_stages=4||1
decode_target_ctas=240||1
|pro_ctas=0|1
|pro_ctas=240|1
decode_max_stages=3|pro_ctas=0|1
||8
decode_impl=1||8
||32
decode_impl=1||32
EOT
cat $OUT/r2d_ab.txt
