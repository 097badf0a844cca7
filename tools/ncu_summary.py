#!/usr/bin/env python
"""Key per-launch numbers of an .ncu-rep (read with `ncu -i ... --page raw --csv`): duration, DRAM bytes, achieved GB/s,
DRAM %, occupancy, registers, shared memory.  Usage: python tools/ncu_summary.py file.ncu-rep [peak_gbs]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    path = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6541.1
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[col["Kernel Name"]][:90]
        print(f"== {name}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        vals = {}
        for k in KEYS:
            if k in col:
                vals[k] = (r[col[k]], units[col[k]])
                print(f"   {k:70s} {r[col[k]]:>16s} {units[col[k]]}")
        try:
            def num(k):
                v, u = vals[k]
                v = float(v.replace(",", ""))
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3}.get(u, 1)
                return v * scale
            t = num("gpu__time_duration.sum")
            b = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            print(f"   -> DRAM traffic {b / 1e6:.2f} MB in {t * 1e6:.2f} us = {b / t / 1e9:.0f} GB/s = {b / t / 1e9 / peak:.3f} of measured peak {peak:.0f} GB/s")
        except Exception as exc:  # noqa: BLE001
            print("   (no derived numbers:", exc, ")")


if __name__ == "__main__":
    main()
