#!/usr/bin/env python
"""Benchmark of the CodeFuse-13B weight-only-INT8 request path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: HF-transformers GPT-NeoX on the host cores

A "step" is one whole request of the headline workload: CodeFuse-13B shape (h 5120, 40 heads x 128, 40 layers, inter 20480,
vocab 100864), int8_mode = 1, batch 1, 1024 prompt tokens, 512 generated tokens, greedy, end_id never sampled.  Random-init
weights quantised with the reference's symmetric per-column rule, synthetic prompt ids.  `value` is generated tokens per
second over the whole request (prefill + decode) -- the quantity the reference publishes as "Tokens Per Sec"
(/root/reference/README.md:95-99, 512 / latency) -- with the inputs already resident in HBM; `e2e` is the same through
GptNeoXOp.forward with HOST buffers (pinned ids -> H2D, outputs -> D2H inside the timed region).  With N > 1 the model is
tensor-parallel over the N GPUs (same request, strong scaling) with NCCL all-reduces where the reference has them.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PUBLISHED_TOKENS_PER_S = {1: 75.0, 2: 98.0}      # int8, 1xA100 / 2xA100 TP ("Tokens Per Sec")
HBM_FALLBACK_GBS = 6650.0                        # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the decode INT8 GEMMs, from the `ncu --set full` capture in
# profiles/r1s_gemm_skinny_ncu_full.txt: FFN1 108.55 MB, FFN2 108.77 MB, QKV 80.66 MB, O 26.26 MB -> mean of the four shapes
# (algorithmic mean: 78.64 MB; the extra 3 % is the activations / partial sector fetches)
GEMM_TRAFFIC_BYTES_PER_LAUNCH_TP1 = (108.55e6 + 108.77e6 + 80.66e6 + 26.26e6) / 4

MODEL = dict(head_num=40, size_per_head=128, inter_size=20480, layer_num=40, vocab_size=100864, rotary_embedding_dim=128)
B1 = dict(batch=1, in_len=1024, out_len=512)
B32 = dict(batch=32, in_len=1024, out_len=512)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # only the samples taken under load say anything about a clock lock
        load = [c for c in sm if mx and c >= 0.5 * max(mx)] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_baseline(sample_layers=(2, 4), in_len=16, out_len=8):
    """HF-transformers GPTNeoXForCausalLM on the host cores (fp32, tanh-GELU, parallel residual), CodeFuse-13B shape truncated
    to `sample_layers` layers; per-token decode time is separated into a per-layer and a fixed (embedding + LM head) part from
    the two depths and extrapolated to 40 layers.  Bounded on purpose: the full model is 52 GB in fp32."""
    import torch
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    h = MODEL["head_num"] * MODEL["size_per_head"]
    per_tok = {}
    t_all0 = time.perf_counter()
    nl_max = max(sample_layers)
    cfg = GPTNeoXConfig(hidden_size=h, num_hidden_layers=nl_max, num_attention_heads=MODEL["head_num"],
                        intermediate_size=MODEL["inter_size"], vocab_size=MODEL["vocab_size"], hidden_act="gelu_new",
                        use_parallel_residual=True, max_position_embeddings=2048, tie_word_embeddings=False)
    with torch.device("meta"):
        model = GPTNeoXForCausalLM(cfg)
    model = model.to_empty(device="cpu")
    with torch.no_grad():
        for p in model.parameters():
            p.normal_(0.0, 0.02)
        for name, b in model.named_buffers():
            if "inv_freq" in name:
                dim = b.shape[0] * 2
                b.copy_(1.0 / (10000 ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim)))
    model.eval()
    all_layers = list(model.gpt_neox.layers)
    ids = torch.randint(0, MODEL["vocab_size"], (1, in_len), generator=torch.Generator().manual_seed(1234))
    # the SAME weights at both depths (the deeper model first, then its first layers only), three untimed tokens each
    for nl in sorted(sample_layers, reverse=True):
        model.gpt_neox.layers = torch.nn.ModuleList(all_layers[:nl])
        model.config.num_hidden_layers = nl
        with torch.no_grad():
            out = model(ids, use_cache=True)                    # prefill
            past, nxt = out.past_key_values, out.logits[:, -1:].argmax(-1)
            for _ in range(3):
                out = model(nxt, past_key_values=past, use_cache=True)
                past, nxt = out.past_key_values, out.logits[:, -1:].argmax(-1)
            t0 = time.perf_counter()
            for _ in range(out_len):
                out = model(nxt, past_key_values=past, use_cache=True)
                past, nxt = out.past_key_values, out.logits[:, -1:].argmax(-1)
            per_tok[nl] = (time.perf_counter() - t0) / out_len
        del out, past
    del model, all_layers
    a, b = sample_layers
    t_layer = max((per_tok[b] - per_tok[a]) / (b - a), 1e-9)
    t_fixed = max(per_tok[a] - a * t_layer, 0.0)
    t40 = t_fixed + MODEL["layer_num"] * t_layer
    return {"value": 1.0 / t40, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": (f"HF transformers GPTNeoXForCausalLM fp32 on CPU ({cores} threads), CodeFuse-13B shape at {a} and {b} layers, batch 1, "
                       f"{in_len} in / {out_len} out, decode ms/token {per_tok[a] * 1e3:.1f} / {per_tok[b] * 1e3:.1f} -> per-layer "
                       f"{t_layer * 1e3:.2f} ms + fixed {t_fixed * 1e3:.2f} ms, extrapolated to 40 layers; "
                       f"sample wall time {time.perf_counter() - t_all0:.0f} s"),
            "ms_per_token": t40 * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline()
    ms_step = cb["ms_per_token"] * B1["out_len"]
    line = {"impl": "reference", "metric": "tokens/s (generated tokens / request latency), CodeFuse-13B, batch 1, 1024 in / 512 out",
            "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32 (CPU)",
            "data": "synthetic", "config": {"workload": "CodeFuse-13B shape, batch 1, decode tokens/s on host CPU (bounded sample, see cpu_baseline.sample)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def algorithmic_bytes_per_step(t, batch, ctx):
    """layer weights (int8) + fp16 LM head rows of this rank + KV read + KV write, per GPU."""
    h, L, V = MODEL["head_num"] * MODEL["size_per_head"], MODEL["layer_num"], MODEL["vocab_size"]
    return (L * 12 * h * h * 1 + 2 * V * h) / t + batch * ctx * (2 * L * h * 2) / t + batch * (2 * L * h * 2) / t


def time_gemm_kernel(rw, t, torch, capi):
    """The dominant decode kernel alone: all 4 x L weight-only-INT8 GEMMs of one token (m = 1), back to back on the current
    stream, CUDA events around them.  Every launch reads a different weight matrix (12.6 GB/t in total >> 126 MB of L2)."""
    lib = capi.load()
    L = MODEL["layer_num"]
    h = MODEL["head_num"] * MODEL["size_per_head"]
    hl, il = h // t, MODEL["inter_size"] // t
    shapes = [(h, 3 * hl), (hl, h), (h, il), (il, h)]       # (k, n) of qkv, o, ffn1, ffn2
    dev = rw.int8_w[0].device
    x = torch.randn(1, max(h, il), device=dev).half()
    y = torch.empty(1, max(3 * hl, il, h), dtype=torch.float16, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def one_pass():
        for layer in range(L):
            for kind, (k, n) in enumerate(shapes):
                capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), rw.int8_w[kind * L + layer].data_ptr(), rw.scale[kind * L + layer].data_ptr(),
                                               None, y.data_ptr(), 1, n, k, 0, 0, st))
    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    total_bytes = L * sum(k * n for k, n in shapes)
    launches = 4 * L
    return {"avg_launch_us": ms * 1e3 / launches, "bytes_per_launch": total_bytes / launches, "gbs": total_bytes / (ms * 1e-3) / 1e9,
            "launches_per_pass": launches}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fastertransformer4codefuse_b200 import capi, weights as W
    from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    capi.check(capi.load().ftcf_device_check())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t = world

    cfg = W.NeoXConfig(start_id=100000, end_id=MODEL["vocab_size"] - 1, use_gptj_residual=True, **MODEL)
    t_w0 = time.perf_counter()
    rw = W.make_synthetic_fast(cfg, t, rank, 1, dev, seed=0)
    # the end_id row of the LM head is zeroed and its logit can never win: exactly out_len tokens are generated
    rw.w[12 * cfg.layer_num + 3][cfg.end_id].zero_()
    w, q, s = rw.lists()
    comm = dist.group.WORLD if world > 1 else None
    op = GptNeoXOp(comm, rank, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim,
                   cfg.start_id, cfg.end_id, t, 1, 1, 2048, True, w, q, s)
    weight_s = time.perf_counter() - t_w0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def workload(wl, steps, warmup, timing):
        B, S, out = wl["batch"], wl["in_len"], wl["out_len"]
        g = np.random.default_rng(1234)
        ids_host = torch.from_numpy(g.integers(0, cfg.vocab_size - 2, size=(B, S)).astype(np.int32)).pin_memory()
        lens_host = torch.full((B,), S, dtype=torch.int32).pin_memory()
        ids_dev, lens_dev = ids_host.to(dev), lens_host.to(dev)
        op.set_option("step_timing", 1 if timing else 0)
        for _ in range(warmup):
            op.forward(ids_dev, lens_dev, out)
        # ---- device-resident inputs
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches, prefill_ms, decode_ms, step_ms = 0, [], [], []
        for _ in range(steps):
            res = op.forward(ids_dev, lens_dev, out)
            launches += op.last_stats["kernel_launches"]
            prefill_ms.append(op.last_stats["prefill_ms"])
            decode_ms.append(op.last_stats["decode_ms"])
            if timing:
                step_ms += op.last_step_ms()[1:]          # [0] is the first token (LM head + sampling only)
        e1.record()
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        assert int(res[1].min()) == S + out, "a sequence stopped early: the benchmark would be timing less work"
        # ---- end to end through the public call with HOST buffers
        op.set_option("step_timing", 0)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(steps):
            a = ids_host.to(dev, non_blocking=True)
            b = lens_host.to(dev, non_blocking=True)
            r = op.forward(a, b, out)
            out_host = r[0].cpu()
            len_host = r[1].cpu()
        e3.record()
        barrier()
        e2e_ms = max_over_ranks(e2.elapsed_time(e3))
        toks = B * out * steps
        return {"tokens_per_s": toks / (total_ms * 1e-3), "ms_per_request": total_ms / steps, "prefill_ms": statistics.median(prefill_ms),
                "decode_ms": statistics.median(decode_ms), "decode_tokens_per_s": B * (out - 1) / (statistics.median(decode_ms) * 1e-3),
                "p50_token_ms": statistics.median(step_ms) if step_ms else None, "launches": launches,
                "e2e_tokens_per_s": toks / (e2e_ms * 1e-3), "h2d": ids_host.numel() * 4 + lens_host.numel() * 4,
                "d2h": out_host.numel() * 4 + len_host.numel() * 4}

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main = workload(B1, args.steps, args.warmup, timing=True)
    clocks = sampler.stop() if rank == 0 else None
    extra = None
    if not args.skip_batch32:
        extra = workload(B32, 1, 1, timing=True)
    gemm = time_gemm_kernel(rw, t, torch, capi) if rank == 0 else None
    cb = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cb = cpu_baseline()
    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    peak, peak_src = hbm_peak()
    ctx_mean = B1["in_len"] + B1["out_len"] / 2
    step_bytes = algorithmic_bytes_per_step(t, 1, ctx_mean)
    p50 = main["p50_token_ms"]
    line = {
        "metric": "tokens/s (generated tokens / request latency: prefill + decode), CodeFuse-13B int8, batch 1, 1024 in / 512 out",
        "value": main["tokens_per_s"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_request"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": (main["tokens_per_s"] / PUBLISHED_TOKENS_PER_S[world]) if world in PUBLISHED_TOKENS_PER_S else None,
        "dtype": "int8 weights x fp16 activations, fp32 accumulate", "data": "synthetic",
        "config": {"workload": f"CodeFuse-13B weight-only int8 (int8_mode=1), batch 1, 1024 in / 512 out, greedy, tensor_para={world}",
                   "l2": "inputs larger than L2 (12.6 GB of weights streamed per token, no flush needed)",
                   "parallelism": f"tp{world}", "weights_init_s": round(weight_s, 1)},
        "decode": {"tokens_per_s": main["decode_tokens_per_s"], "p50_token_ms": p50, "prefill_ms": main["prefill_ms"],
                   "decode_ms": main["decode_ms"],
                   "step_roofline_frac": (step_bytes / (p50 * 1e-3) / 1e9 / peak) if p50 else None,
                   "algorithmic_bytes_per_step_per_gpu": step_bytes},
        "e2e": {"value": main["e2e_tokens_per_s"], "unit": "tokens/s", "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"]},
        "gpu_launches": main["launches"],
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": gemm["gbs"], "peak": peak, "unit": "GB/s", "frac": gemm["gbs"] / peak,
                     "traffic": GEMM_TRAFFIC_BYTES_PER_LAUNCH_TP1 if world == 1 else None,
                     "kernel": "gemm_skinny_kernel<uint8_t,...> (weight-only INT8 GEMM, m = 1): all 160 layer GEMMs of one token",
                     "avg_launch_us": gemm["avg_launch_us"], "bytes_per_launch": gemm["bytes_per_launch"], "peak_source": peak_src},
    }
    if extra is not None:
        ctx32 = B32["in_len"] + B32["out_len"] / 2
        b32_bytes = algorithmic_bytes_per_step(t, 32, ctx32)
        line["batch32"] = {"workload": "batch 32, 1024 in / 512 out", "tokens_per_s": extra["tokens_per_s"],
                           "decode_tokens_per_s": extra["decode_tokens_per_s"], "p50_token_ms": extra["p50_token_ms"],
                           "prefill_ms": extra["prefill_ms"], "e2e_tokens_per_s": extra["e2e_tokens_per_s"],
                           "step_roofline_frac": (b32_bytes / (extra["p50_token_ms"] * 1e-3) / 1e9 / peak) if extra["p50_token_ms"] else None}
    if cb is not None:
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-batch32", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
