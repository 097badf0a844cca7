#!/usr/bin/env python
"""Per-K-step clock stamps of one CTA of the tcgen05 decode GEMM (ftcf_debug_decode_probe): where a K step's ~cycles go.
    python tools/decode_gemm_probe.py [n] [k] [m]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fastertransformer4codefuse_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20480
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5120
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib = capi.load()
dev = torch.device("cuda:0")
ws = [torch.randint(0, 255, (n, k), dtype=torch.uint8, device=dev) for _ in range(3)]
sc = torch.full((n,), 0.01, dtype=torch.float16, device=dev)
x = torch.randn(m, k, device=dev).half()
y = torch.empty(m, n, dtype=torch.float16, device=dev)
st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
for i in range(3):
    capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), ws[i % 3].data_ptr(), sc.data_ptr(), None, y.data_ptr(), m, n, k, 0, 3, st))
torch.cuda.synchronize()
capi.check(lib.ftcf_debug_decode_probe(buf.data_ptr()))
capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), ws[0].data_ptr(), sc.data_ptr(), None, y.data_ptr(), m, n, k, 0, 3, st))
torch.cuda.synchronize()
capi.check(lib.ftcf_debug_decode_probe(None))
t = buf.cpu().numpy().reshape(64, 8)
t0 = t[0, 0]
print(f"n={n} k={k} m={m}: cycles relative to the converter's first stamp")
print(" kb | conv: wait_w  +got_w  +converted  +got_tmem  +stored | mma: wait_a  +got_a  +issued | step")
prev = t0
for kb in range(64):
    if t[kb, 0] == 0:
        break
    r = t[kb] - t0
    print(f"{kb:3d} | {r[0]:8d} {r[1]-r[0]:7d} {r[2]-r[1]:10d} {r[3]-r[2]:9d} {r[4]-r[3]:8d} | {r[5]:8d} {r[6]-r[5]:7d} {r[7]-r[6]:8d} | {t[kb,4]-prev:6d}")
    prev = t[kb, 4]
