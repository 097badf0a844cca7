#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by warp-stall samples, with totals per opcode and per stall reason.
Usage: python tools/ncu_hot_sass.py file.ncu-rep [N]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def iv(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    col = {h: i for i, h in enumerate(hdr)}
    samp = col["# Samples"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(iv(r[samp]) for r in data)
    print(f"{len(data)} SASS instructions, {tot} samples")
    agg = {h: sum(iv(r[col[h]]) for r in data) for h in stalls}
    print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
    ops = Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
        ops[m.group(2) if m else "?"] += iv(r[samp])
    print("by opcode:", ", ".join(f"{k} {100 * v / max(tot, 1):.1f}%" for k, v in ops.most_common(14)))
    for i, r in sorted(enumerate(data), key=lambda ir: -iv(ir[1][samp]))[:n]:
        s = sorted(((h[6:], iv(r[col[h]])) for h in stalls), key=lambda kv: -kv[1])[:2]
        print(f"{iv(r[samp]):7d} {100 * iv(r[samp]) / max(tot, 1):5.1f}%  #{i:5d} {r[col['Source']].strip()[:80]:80s} {s}")


if __name__ == "__main__":
    main()
