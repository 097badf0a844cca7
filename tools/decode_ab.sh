#!/bin/bash
# A/B of decode-step settings: each line of stdin is "FTCF_TUNABLES|FTCF_OPTIONS|batch"; prints decode ms for 128 steps.
while IFS='|' read -r tun opt batch; do
  [ -z "$batch" ] && continue
  r=$(FTCF_TUNABLES="$tun" FTCF_OPTIONS="$opt" timeout 300 python tools/profile_decode.py --out-len 129 --graph 1 --batch $batch --requests 2 2>&1 | tail -1)
  echo "== TUN=$tun OPT=$opt batch=$batch	$r"
done
