#!/bin/bash
# one-off A/B driver: lines "TUNABLES|OPTIONS|batch" from $1 (a file) through tools/decode_ab.sh, result into gpurun_out/$2
OUT=gpurun_out
mkdir -p $OUT
bash tools/decode_ab.sh < "$1" | tee $OUT/$2
