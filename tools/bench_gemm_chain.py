#!/usr/bin/env python
"""Micro-benchmark: the 160 weight-only-INT8 GEMMs of one decode token (13B shape) back to back through the C ABI, for a few
tunable settings.  Prints GB/s of weight bytes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fastertransformer4codefuse_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, nargs="+", default=[1])
ap.add_argument("--layers", type=int, default=40)
ap.add_argument("--impl", type=int, default=1)
ap.add_argument("--tun", type=str, default="", help="extra tunables name=value,... applied to every tcgen05 run")
a = ap.parse_args()
lib = capi.load()
dev = torch.device("cuda:0")
h, inter, L = 5120, 20480, a.layers
shapes = [(h, 3 * h), (h, h), (h, inter), (inter, h)]
ws = [[torch.randint(0, 256, (n, k), dtype=torch.uint8, device=dev) for (k, n) in shapes] for _ in range(L)]
sc = [torch.rand(n, device=dev).half() * 0.01 for (k, n) in shapes]
st = torch.cuda.current_stream().cuda_stream


def run(m, reps=10):
    x = torch.randn(m, inter, device=dev).half()
    y = torch.empty(m, inter, dtype=torch.float16, device=dev)

    def one(stream):
        for layer in range(L):
            for i, (k, n) in enumerate(shapes):
                capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), ws[layer][i].data_ptr(), sc[i].data_ptr(), None, y.data_ptr(), m, n, k, 0, a.impl, stream))
    # captured into a CUDA graph so that the host launch rate (ctypes + driver) is not what is measured
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        one(s.cuda_stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            one(s.cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, L * sum(k * n for k, n in shapes) / (ms * 1e-3) / 1e9


for m in a.m:
    if a.impl == 3:
        for item in filter(None, a.tun.split(",")):
            k_, _, v_ = item.partition("=")
            capi.check(lib.ftcf_set_tunable(k_.encode(), int(v_)))
        for pdl in (1,):
            for ctas, min_kb in ((296, 8), (148, 8), (160, 40), (240, 8), (296, 4)):
                lib.ftcf_set_tunable(b"pdl", pdl)
                lib.ftcf_set_tunable(b"decode_target_ctas", ctas)
                lib.ftcf_set_tunable(b"decode_min_kb", min_kb)
                ms, gbs = run(m)
                print(f"tcgen05 decode m={m:3d} pdl={pdl} target_ctas={ctas} min_kb={min_kb}: {ms:7.3f} ms/token-pass  {gbs:7.1f} GB/s", flush=True)
        continue
    for pdl in (0, 1):
        for ctas in (148, 296):
            lib.ftcf_set_tunable(b"pdl", pdl)
            lib.ftcf_set_tunable(b"skinny_target_ctas", ctas)
            ms, gbs = run(m)
            print(f"skinny m={m:3d} pdl={pdl} ctas={ctas}: {ms:7.3f} ms/token-pass  {gbs:7.1f} GB/s", flush=True)
