#!/usr/bin/env python
"""Top source lines of an .ncu-rep by warp-stall samples (needs -lineinfo and --import-source on).
Usage: python tools/ncu_hot_lines.py file.ncu-rep [N]"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, hdr, data = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 3 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is not None and len(r) == len(hdr):
            data.append((fname, r))
    col = {h: i for i, h in enumerate(hdr)}
    samp = col["# Samples"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def iv(x):
        try:
            return int(x)
        except ValueError:
            return 0
    tot = sum(iv(r[samp]) for _, r in data)
    print(f"total samples {tot}")
    agg = {h: sum(iv(r[col[h]]) for _, r in data) for h in stalls}
    print("stall totals:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for f, r in sorted(data, key=lambda fr: -iv(fr[1][samp]))[:n]:
        s = sorted(((h[6:], iv(r[col[h]])) for h in stalls), key=lambda kv: -kv[1])[:3]
        print(f"{iv(r[samp]):7d} {100 * iv(r[samp]) / max(tot, 1):5.1f}%  {f}:{r[0]:>4}  {r[1].strip()[:95]:95s} {s}")


if __name__ == "__main__":
    main()
