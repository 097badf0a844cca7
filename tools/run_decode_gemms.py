#!/usr/bin/env python
"""The four INT8 GEMM shapes of a 13B decode layer, a few launches each on rotating weight buffers (cold L2), for ncu:
    ncu --set full -k regex:gemm_decode ... python tools/run_decode_gemms.py [m] [impl] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fastertransformer4codefuse_b200 import capi

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = capi.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
for (n, k) in [(20480, 5120), (5120, 20480), (15360, 5120), (5120, 5120)]:
    ws = [torch.randint(0, 255, (n, k), dtype=torch.uint8, device=dev) for _ in range(4)]
    sc = torch.full((n,), 0.01, dtype=torch.float16, device=dev)
    x = torch.randn(m, k, device=dev).half()
    y = torch.empty(m, n, dtype=torch.float16, device=dev)
    for i in range(reps):
        capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), ws[i % 4].data_ptr(), sc.data_ptr(), None, y.data_ptr(), m, n, k, 0, impl, st))
    torch.cuda.synchronize()
print("done")
