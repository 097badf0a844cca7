#!/usr/bin/env python
"""Per-barrier timing of the persistent decode kernel (mega_dbg=64): arrival spread across CTAs vs release latency."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FTCF_TUNABLES"] = os.environ.get("FTCF_TUNABLES", "mega_dbg=64")
import numpy as np
import torch

from fastertransformer4codefuse_b200 import capi, weights as W
from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp

dev = torch.device("cuda:0")
cfg = W.NeoXConfig(head_num=40, size_per_head=128, inter_size=20480, layer_num=40, vocab_size=100864, rotary_embedding_dim=128,
                   start_id=100000, end_id=100863)
rw = W.make_synthetic_fast(cfg, 1, 0, 1, dev)
w, q, s = rw.lists()
op = GptNeoXOp(None, 0, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim, cfg.start_id,
               cfg.end_id, 1, 1, 1, 2048, True, w, q, s)
ids = torch.from_numpy(np.random.default_rng(1234).integers(0, cfg.vocab_size - 2, size=(1, 1024)).astype(np.int32)).to(dev)
lens = torch.full((1,), 1024, dtype=torch.int32, device=dev)
op.forward(ids, lens, 8)
lib = capi.load()
NB, NC = 130, 160
buf = np.zeros((3, NB, NC), dtype=np.uint64)
fn = lib.ftcf_debug_mega_timestamps
fn.argtypes = [C.c_void_p, C.c_size_t]
assert fn(buf.ctypes.data, buf.nbytes) == 0
t = buf[:, :120, :148].astype(np.int64)
t0 = t[0].min()
enter, arrive, release = t[0] - t0, t[1] - t0, t[2] - t0
print("kernel span (first enter of barrier 0 .. last release of barrier 119): %.1f us" % ((release[119].max()) / 1e3))
print("bar  phase  first_enter  last_enter  spread   fence+cbar(avg)  release-after-last-arrive(avg/max)   next-phase-length")
for i in range(0, 120):
    if i < 12 or i % 12 < 3:
        spread = (enter[i].max() - enter[i].min()) / 1e3
        fence = (arrive[i] - enter[i]).mean() / 1e3
        rel = release[i] - arrive[i].max()
        nxt = (enter[i + 1].min() - release[i].max()) / 1e3 if i + 1 < 120 else 0
        print(f"{i:3d}  {'ABC'[i % 3]}  {enter[i].min() / 1e3:10.1f} {enter[i].max() / 1e3:10.1f} {spread:8.1f} {fence:12.2f} {rel.mean() / 1e3:14.2f} {rel.max() / 1e3:8.2f} {nxt:12.1f}")
tot_spread = sum((enter[i].max() - enter[i].min()) for i in range(120)) / 1e3
tot_rel = sum((release[i].max() - arrive[i].max()) for i in range(120)) / 1e3
tot_fence = sum((arrive[i] - enter[i]).max() for i in range(120)) / 1e3
print(f"sum over 120 barriers: arrival spread {tot_spread:.0f} us, fence+cbar {tot_fence:.0f} us, release latency {tot_rel:.0f} us")
# which CTAs are late?
late = np.zeros(148)
for i in range(120):
    late += (enter[i] - enter[i].min()) / 1e3
print("mean lateness per CTA (us per barrier): min %.2f  median %.2f  max %.2f ; worst CTAs %s" % (late.min() / 120, np.median(late) / 120, late.max() / 120, np.argsort(-late)[:8]))

cyc = np.zeros((8, NC), dtype=np.int64)
fn2 = lib.ftcf_debug_mega_cycles
fn2.argtypes = [C.c_void_p, C.c_size_t]
assert fn2(cyc.ctypes.data, cyc.nbytes) == 0
c = cyc[:, :148].astype(np.float64)
print("per CTA, last step (mean over CTAs, Mcycles): producer total %.2f  empty-wait %.2f (%.0f%%)  throttle-wait %.2f (%.0f%%) | consumer total %.2f  warp0 full-wait %.2f (%.0f%%)  warp7 full-wait %.2f (%.0f%%)" % (
    c[2].mean() / 1e6, c[0].mean() / 1e6, 100 * c[0].mean() / c[2].mean(), c[1].mean() / 1e6, 100 * c[1].mean() / c[2].mean(),
    c[4].mean() / 1e6, c[3].mean() / 1e6, 100 * c[3].mean() / c[4].mean(), c[5].mean() / 1e6, 100 * c[5].mean() / c[4].mean()))
