#!/usr/bin/env python
"""Per-launch timeline of ONE decode step (graph replay, headline shape) from the per-CTA trace records
(ftcf_debug_trace_start / _stop): for every kernel launch of the step -- first CTA start, last CTA start, when the
dependency wait ended, when the first weight stage landed, last CTA end -- plus how long HBM had no streaming kernel.
    python tools/trace_step.py [--batch 1] [--layers 40] [--show 2]  (FTCF_OPTIONS / FTCF_TUNABLES apply)"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fastertransformer4codefuse_b200 import capi, weights as W
from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--in-len", type=int, default=1024)
ap.add_argument("--layers", type=int, default=40)
ap.add_argument("--show", type=int, default=2, help="layers to print launch by launch")
ap.add_argument("--detail", type=int, default=0, help="1: per-launch distribution of CTA end times and busy times (min 10% 25% 50% 75% 90% max)")
a = ap.parse_args()
# under torchrun (WORLD_SIZE > 1) the model is tensor-parallel and rank 0 prints ITS timeline
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
comm = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    comm = dist.group.WORLD
cfg = W.NeoXConfig(head_num=40, size_per_head=128, inter_size=20480, layer_num=a.layers, vocab_size=100864, rotary_embedding_dim=128,
                   start_id=100000, end_id=100863)
rw = W.make_synthetic_fast(cfg, world, rank, 1, dev)
rw.w[12 * cfg.layer_num + 3][cfg.end_id].zero_()
w, q, s = rw.lists()
op = GptNeoXOp(comm, rank, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim, cfg.start_id,
               cfg.end_id, world, 1, 1, 2048, True, w, q, s)
ids = torch.from_numpy(np.random.default_rng(1234).integers(0, cfg.vocab_size - 2, size=(a.batch, a.in_len)).astype(np.int32)).to(dev)
lens = torch.full((a.batch,), a.in_len, dtype=torch.int32, device=dev)
op.forward(ids, lens, 12)          # warm-up: captures the graph
lib = capi.load()
CAP = 3_000_000 if a.batch <= 4 else 16_000_000
capi.check(lib.ftcf_debug_trace_start(CAP))
op.forward(ids, lens, 12)
rec_t = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("t2", "<u8"), ("t3", "<u8"), ("kind", "<i4"), ("cta", "<i4"), ("ncta", "<i4"),
                  ("a", "<i4"), ("b", "<i4"), ("pad", "<i4")])
buf = np.zeros(CAP, dtype=rec_t)
n = C.c_uint(0)
capi.check(lib.ftcf_debug_trace_stop(buf.ctypes.data, CAP, C.byref(n)))
r = buf[:n.value]
if rank != 0:
    if world > 1:
        dist.barrier()
    sys.exit(0)
print("records:", len(r), op.last_stats)
ends = np.sort(r[(r["kind"] == 30) & (r["b"] == 3)]["t3"])     # step_finalize of every step
assert len(ends) >= 4, len(ends)
lo, hi = ends[-3], ends[-2]                                    # one step in the middle of the replayed ones
st = r[(r["t0"] > lo) & (r["t3"] <= hi + 1)]
st = st[np.argsort(st["t0"])]
T0 = lo
print(f"step window {(hi - lo) / 1e3:.1f} us, {len(st)} CTA records")
# group into launches: same (kind, a, b, ncta), consecutive in time, ncta records each
names = {1: "gemm_w8", 2: "gemm_f16", 10: "mmha", 20: "ln", 21: "residual", 22: "embed", 30: "sampling"}
launches = []
keys = {}
for x in st:
    k = (int(x["kind"]), int(x["a"]) if x["kind"] != 10 else 0, int(x["b"]) if x["kind"] in (1, 2) else 0, int(x["ncta"]))
    g = keys.get(k)
    if g is None or len(g["recs"]) >= k[3]:
        g = {"key": k, "recs": []}
        keys[k] = g
        launches.append(g)
    g["recs"].append(x)
rows = []
for g in launches:
    rr = np.array(g["recs"], dtype=rec_t)
    k = g["key"]
    rows.append(dict(name=names.get(k[0], str(k[0])), n=k[1], k=k[2], ncta=k[3], got=len(rr), first=(rr["t0"].min() - T0) / 1e3,
                     last_start=(rr["t0"].max() - T0) / 1e3, wait_end=(rr["t1"].min() - T0) / 1e3,
                     first_data=(rr["t2"][rr["t2"] > 0].min() - T0) / 1e3 if (rr["t2"] > 0).any() else float("nan"),
                     first_end=(rr["t3"].min() - T0) / 1e3, end=(rr["t3"].max() - T0) / 1e3,
                     mb=(k[1] * k[2] * (1 if k[0] == 1 else 2) / 1e6) if k[0] in (1, 2) else 0.0,
                     pro=float(np.median(rr["pad"])) / 1e3, cta_med=float(np.median(rr["t3"].astype(np.int64) - rr["t0"].astype(np.int64))) / 1e3,
                     data_med=float(np.median((rr["t2"].astype(np.int64) - rr["t1"].astype(np.int64))[rr["t2"] > 0])) / 1e3 if (rr["t2"] > 0).any() else float("nan")))
    r_ = rows[-1]
    t3 = np.sort((rr["t3"].astype(np.int64) - int(T0)) / 1e3)
    dur = np.sort((rr["t3"].astype(np.int64) - np.maximum(rr["t2"], rr["t1"]).astype(np.int64)) / 1e3)
    pick = lambda a_: " ".join(f"{a_[min(len(a_) - 1, int(qq * (len(a_) - 1)))]:.1f}" for qq in (0, .1, .25, .5, .75, .9, 1))
    r_["detail"] = f"end deciles [{pick(t3)}]  busy us per CTA [{pick(dur)}]"
rows.sort(key=lambda d: d["first"])
per_layer = max(1, (len(rows) - 4) // max(a.layers, 1))
print(f"{len(rows)} launches in the step (~{per_layer} per layer)")
print(f"{'kernel':9s} {'n':>6s} {'k':>6s} {'CTAs':>5s} | {'start':>8s} {'lastCTA':>8s} {'deps ok':>8s} {'1st data':>8s} {'1st end':>8s} {'end':>8s} | {'dur':>6s} {'GB/s':>7s} | {'prolog':>6s} {'->data':>6s} {'CTA us':>6s}  (medians per CTA)")
mid = len(rows) // 2
sel = rows[:per_layer + 2] + rows[mid - (mid % per_layer if per_layer else 0):][:per_layer * a.show] + rows[-4:]
for d in sel:
    dur = d["end"] - d["wait_end"] if d["name"].startswith("gemm") else d["end"] - d["first"]
    gbs = d["mb"] / dur * 1e3 / 1e3 if d["mb"] and dur > 0 else 0
    print(f"{d['name']:9s} {d['n']:6d} {d['k']:6d} {d['ncta']:5d} | {d['first']:8.1f} {d['last_start']:8.1f} {d['wait_end']:8.1f} {d['first_data']:8.1f} "
          f"{d['first_end']:8.1f} {d['end']:8.1f} | {dur:6.1f} {gbs * 1e3:7.0f} | {d['pro']:6.1f} {d['data_med']:6.1f} {d['cta_med']:6.1f}")
    if a.detail:
        print("          " + d["detail"])
# aggregate per kernel type
agg = {}
for d in rows:
    key = (d["name"], d["n"], d["k"])
    e = agg.setdefault(key, dict(cnt=0, dur=0.0, active=0.0, mb=0.0, wait=0.0))
    e["cnt"] += 1
    e["dur"] += d["end"] - d["first"]
    e["active"] += d["end"] - (d["wait_end"] if d["name"].startswith("gemm") else d["first"])
    e["wait"] += d["wait_end"] - d["first"]
    e["mb"] += d["mb"]
print("\nper kernel type over the step: count, mean resident us, mean dependency-wait us, mean active us, GB/s while active")
for key, e in sorted(agg.items(), key=lambda kv: -kv[1]["active"]):
    c = e["cnt"]
    print(f"  {key[0]:9s} n={key[1]:6d} k={key[2]:6d}  x{c:3d}  resident {e['dur'] / c:7.1f}  wait {e['wait'] / c:6.1f}  active {e['active'] / c:6.1f}  "
          f"{(e['mb'] / e['active'] * 1e3) if e['active'] > 0 and e['mb'] else 0:7.0f}")
# time with no GEMM / attention CTA past its dependency wait
ev = []
for d in rows:
    if d["name"] in ("gemm_w8", "gemm_f16", "mmha"):
        ev.append((d["wait_end"] if d["name"] != "mmha" else d["first"], 1))
        ev.append((d["end"], -1))
ev.sort()
idle, depth, last = 0.0, 0, 0.0
for t, dlt in ev:
    if depth == 0:
        idle += t - last
    depth += dlt
    last = t
print(f"\ntime with NO streaming kernel active (past its dependency wait): {idle:.1f} us of {(hi - lo) / 1e3:.1f} us")
if world > 1:
    dist.barrier()
