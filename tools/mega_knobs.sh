#!/bin/bash
# timing experiments on the persistent decode kernel (results are wrong for dbg != 0; only decode_ms matters)
for cfg in "$@"; do
  echo "== $cfg"
  FTCF_TUNABLES=$cfg timeout 120 python tools/profile_decode.py --out-len 33 --requests 2 --graph 1 2>&1 | tail -1
done
