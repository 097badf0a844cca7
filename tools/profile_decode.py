#!/usr/bin/env python
"""One short request of the headline shape, for ncu: `ncu ... python tools/profile_decode.py --out-len 3`.
Graph replay is switched off so that every kernel is an ordinary launch in the profiler's list."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fastertransformer4codefuse_b200 import weights as W
from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--in-len", type=int, default=1024)
ap.add_argument("--out-len", type=int, default=3)
ap.add_argument("--layers", type=int, default=40)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--gemm-impl", type=int, default=0)
ap.add_argument("--requests", type=int, default=1)
a = ap.parse_args()

dev = torch.device("cuda:0")
cfg = W.NeoXConfig(head_num=40, size_per_head=128, inter_size=20480, layer_num=a.layers, vocab_size=100864, rotary_embedding_dim=128,
                   start_id=100000, end_id=100863)
rw = W.make_synthetic_fast(cfg, 1, 0, 1, dev)
rw.w[12 * cfg.layer_num + 3][cfg.end_id].zero_()
w, q, s = rw.lists()
op = GptNeoXOp(None, 0, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim, cfg.start_id,
               cfg.end_id, 1, 1, 1, 2048, True, w, q, s)
op.set_option("cuda_graph", a.graph)
op.set_option("gemm_impl", a.gemm_impl)
ids = torch.from_numpy(np.random.default_rng(1234).integers(0, cfg.vocab_size - 2, size=(a.batch, a.in_len)).astype(np.int32)).to(dev)
lens = torch.full((a.batch,), a.in_len, dtype=torch.int32, device=dev)
for _ in range(a.requests):
    op.forward(ids, lens, a.out_len)
    print(op.last_stats)
