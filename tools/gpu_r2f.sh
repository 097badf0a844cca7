#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for tun in "" "decode_fake_tiled=1" "decode_fake_tiled=1,decode_max_stages=3" "decode_max_stages=3"; do
  echo "### tun=$tun"; timeout 300 python tools/bench_gemm_chain.py --impl 3 --m 1 --tun "$tun" 2>&1 | grep tcgen05
done | tee $OUT/r2f_chain.log
bash tools/decode_ab.sh > $OUT/r2f_ab.txt 2>&1 <<'EOT'
||1
decode_fake_tiled=1||1
decode_fake_tiled=1,decode_max_stages=3||1
decode_impl=1||1
EOT
cat $OUT/r2f_ab.txt
