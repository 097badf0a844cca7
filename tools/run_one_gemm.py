#!/usr/bin/env python
"""A few launches of ONE skinny INT8 GEMM shape (for ncu): python tools/run_one_gemm.py n k [m] [launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fastertransformer4codefuse_b200 import capi

n, k = int(sys.argv[1]), int(sys.argv[2])
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 24
lib = capi.load()
dev = torch.device("cuda:0")
ws = [torch.randint(0, 255, (n, k), dtype=torch.uint8, device=dev) for _ in range(8)]
sc = torch.full((n,), 0.01, dtype=torch.float16, device=dev)
x = torch.randn(m, k, device=dev).half()
y = torch.empty(m, n, dtype=torch.float16, device=dev)
st = torch.cuda.current_stream().cuda_stream
for i in range(reps):
    capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), ws[i % 8].data_ptr(), sc.data_ptr(), None, y.data_ptr(), m, n, k, 0, 1, st))
torch.cuda.synchronize()
print("done")
