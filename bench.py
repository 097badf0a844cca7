#!/usr/bin/env python
"""Benchmark of the CodeFuse-13B weight-only-INT8 request path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: HF-transformers GPT-NeoX on the host cores

A "step" is one whole request of the headline workload (BASELINE configs[2]): CodeFuse-13B shape (h 5120, 40 heads x 128,
40 layers, inter 20480, vocab 100864), int8_mode = 1, batch 1, 1024 prompt tokens, 512 generated tokens, greedy, end_id never
sampled.  Random-init weights quantised with the reference's symmetric per-column rule, synthetic prompt ids.  `value` is
generated tokens per second over the whole request (prefill + decode) -- the quantity the reference publishes as "Tokens Per
Sec" (/root/reference/README.md:95-99, 512 / latency) -- with the inputs already resident in HBM; `e2e` is the same through
the drop-in itself, `libth_gptneox.GptNeoXOp.forward` (the compiled pybind11 module a user of codefuse_example.py loads), with
HOST buffers (pinned ids -> H2D, outputs -> D2H inside the timed region).  With N > 1 the model is tensor-parallel over the N
GPUs (same request, strong scaling).  The other BASELINE configurations are reported as sub-objects of the same line:
configs[1] (fp16, 256 in / 128 out) at N = 1, configs[3] (batch 8, 1024 / 512) at N = 2 and configs[4] (batch 32,
2048 in / 512 out) at every N -- its tensor_para = 8 point is the 8-GPU headline, the other N are its throughput sweep.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PUBLISHED_TOKENS_PER_S = {1: 75.0, 2: 98.0}      # int8, 1xA100 / 2xA100 TP ("Tokens Per Sec")
HBM_FALLBACK_GBS = 6650.0                        # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# one metric string for both arms: the driver divides the two lines only when metric, unit and direction agree
METRIC = "tokens/s (generated tokens / request latency), CodeFuse-13B int8, batch 1, 1024 in / 512 out"

MODEL = dict(head_num=40, size_per_head=128, inter_size=20480, layer_num=40, vocab_size=100864, rotary_embedding_dim=128)
C3 = dict(name="configs[2]: int8, batch 1, 1024 in / 512 out", int8=1, batch=1, in_len=1024, out_len=512)
C2 = dict(name="configs[1]: fp16, batch 1, 256 in / 128 out", int8=0, batch=1, in_len=256, out_len=128)
C4 = dict(name="configs[3]: int8, batch 8, 1024 in / 512 out, tensor_para 2", int8=1, batch=8, in_len=1024, out_len=512)
C5 = dict(name="configs[4]: int8, batch 32, 2048 in / 512 out", int8=1, batch=32, in_len=2048, out_len=512)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """Mean dram__bytes_read.sum + dram__bytes_write.sum per launch of the four decode INT8 GEMMs, parsed from the newest
    committed `ncu --set full` summary (profiles/r*_gemm_*ncu_full.txt).  Returns (bytes or None, where it came from)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_*ncu_full.txt")), key=os.path.getmtime)
    unit = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    for path in reversed(files):
        rd, wr = [], []
        for line in open(path):
            m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+([KMG]?byte)", line)
            if m:
                (rd if m.group(1) == "read" else wr).append(float(m.group(2)) * unit[m.group(3).lower()])
        if len(rd) >= 4 and len(rd) == len(wr):
            n = 4 * (len(rd) // 4)
            return sum(rd[:n] + wr[:n]) / n, os.path.relpath(path, ROOT)
    return None, "no profiles/r*_gemm_*ncu_full.txt with four GEMM captures"


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
def cpu_baseline(depths=(1, 2, 4), in_len=C3["in_len"], out_len=C3["out_len"], timed_tokens=32):
    """HF-transformers GPTNeoXForCausalLM on the host cores (fp32, tanh-GELU, parallel residual, all threads) on the headline
    request: CodeFuse-13B shape truncated to `depths` layers (the full model is 52 GB in fp32), the real 1024-token prompt, then
    `timed_tokens` decode tokens at context 1024+ after 3 untimed ones.  Prefill time and the MEDIAN per-token decode time are
    fitted as fixed + per-layer over the three depths (least squares) and extrapolated to 40 layers;
    value = out_len / (prefill_40 + out_len * token_40), the metric of the GPU arm."""
    import numpy as np
    import torch
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    h = MODEL["head_num"] * MODEL["size_per_head"]
    t_all0 = time.perf_counter()
    nl_max = max(depths)
    cfg = GPTNeoXConfig(hidden_size=h, num_hidden_layers=nl_max, num_attention_heads=MODEL["head_num"],
                        intermediate_size=MODEL["inter_size"], vocab_size=MODEL["vocab_size"], hidden_act="gelu_new",
                        use_parallel_residual=True, max_position_embeddings=4096, tie_word_embeddings=False)
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

    def fwd(x, past=None):
        try:
            return model(x, past_key_values=past, use_cache=True, logits_to_keep=1)
        except TypeError:
            return model(x, past_key_values=past, use_cache=True)

    prefill, token = {}, {}
    with torch.no_grad():
        fwd(ids[:, :8])                                          # pages every weight in once (untimed)
        # the SAME weights at every depth (the deepest model first, then its first layers only)
        for nl in sorted(depths, reverse=True):
            model.gpt_neox.layers = torch.nn.ModuleList(all_layers[:nl])
            model.config.num_hidden_layers = nl
            t0 = time.perf_counter()
            out = fwd(ids)
            prefill[nl] = time.perf_counter() - t0
            past, nxt = out.past_key_values, out.logits[:, -1:].argmax(-1)
            ts = []
            for i in range(3 + timed_tokens):
                t0 = time.perf_counter()
                out = fwd(nxt, past)
                past, nxt = out.past_key_values, out.logits[:, -1:].argmax(-1)
                if i >= 3:
                    ts.append(time.perf_counter() - t0)
            token[nl] = statistics.median(ts)
            del out, past
    del model, all_layers
    xs = np.asarray(sorted(depths), dtype=np.float64)
    fit = lambda d: np.polyfit(xs, np.asarray([d[int(x)] for x in xs]), 1)      # [per-layer, fixed]
    (tl, tf), (pl, pf) = fit(token), fit(prefill)
    tl, pl, tf, pf = max(tl, 1e-9), max(pl, 1e-9), max(tf, 0.0), max(pf, 0.0)
    L = MODEL["layer_num"]
    tok40, pre40 = tf + L * tl, pf + L * pl
    latency = pre40 + out_len * tok40
    d = sorted(depths)
    return {"value": out_len / latency, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": (f"HF transformers GPTNeoXForCausalLM fp32 on CPU ({cores} threads), CodeFuse-13B shape at {d} layers, batch 1, "
                       f"{in_len}-token prompt, median of {timed_tokens} decode tokens at context {in_len}+: "
                       f"ms/token {[round(token[x] * 1e3, 1) for x in d]}, prefill s {[round(prefill[x], 2) for x in d]} -> per token "
                       f"{tl * 1e3:.2f} ms/layer + {tf * 1e3:.2f} ms fixed, prefill {pl:.3f} s/layer + {pf:.2f} s fixed; extrapolated "
                       f"to 40 layers: {tok40 * 1e3:.0f} ms/token, prefill {pre40:.1f} s, request {latency:.0f} s; "
                       f"sample wall time {time.perf_counter() - t_all0:.0f} s"),
            "ms_per_token": tok40 * 1e3, "ms_per_request": latency * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline()
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_request"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "fp32 (CPU)", "data": "synthetic",
            "config": {"workload": "CodeFuse-13B shape, batch 1, 1024 in / 512 out on the host CPU (bounded sample extrapolated to 40 "
                                   "layers, see cpu_baseline.sample)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def algorithmic_bytes_per_step(t, batch, ctx, wbytes=1):
    """layer weights + fp16 LM head rows of this rank + KV read + KV write, per GPU."""
    h, L, V = MODEL["head_num"] * MODEL["size_per_head"], MODEL["layer_num"], MODEL["vocab_size"]
    return (L * 12 * h * h * wbytes + 2 * V * h) / t + batch * ctx * (2 * L * h * 2) / t + batch * (2 * L * h * 2) / t


def prefill_flops(t, batch, s):
    """2 T (12 h^2 L) / t + 4 sum_b S_b^2 h L / (2 t)  (causal)."""
    h, L = MODEL["head_num"] * MODEL["size_per_head"], MODEL["layer_num"]
    return 2.0 * batch * s * (12 * h * h * L) / t + 4.0 * batch * s * s * h * L / (2 * t)


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


def time_mmha_kernel(t, batch, ctx, torch, capi):
    """The dominant kernel of the batch-32 decode step alone: the decode attention of one layer (heads / t, KV rows of `batch`
    sequences at context `ctx`), CUDA events around repeated launches.  One layer's K + V is >> L2 at batch 32."""
    import math
    lib = capi.load()
    H, dh = MODEL["head_num"] // t, MODEL["size_per_head"]
    max_len = ctx + 8
    dev = torch.device("cuda", torch.cuda.current_device())
    kc = torch.randn(batch, H, max_len, dh, device=dev, dtype=torch.float16)
    vc = torch.randn(batch, H, max_len, dh, device=dev, dtype=torch.float16)
    qkv = torch.randn(batch, 3 * H * dh, device=dev, dtype=torch.float16)
    ctxo = torch.empty(batch, H * dh, device=dev, dtype=torch.float16)
    seq = torch.full((batch,), ctx - 1, dtype=torch.int32, device=dev)
    inl = torch.full((batch,), 16, dtype=torch.int32, device=dev)
    pad = torch.zeros(batch, dtype=torch.int32, device=dev)
    fin = torch.zeros(batch, dtype=torch.uint8, device=dev)
    step = torch.tensor([ctx], dtype=torch.int32, device=dev)
    splits = lib.ftcf_mmha_choose_splits(batch, H, max_len)
    part = torch.zeros(batch * H * splits * (dh + 2) + 64, dtype=torch.float32, device=dev)
    cnt = torch.zeros(batch * H, dtype=torch.int32, device=dev)
    p = capi.MmhaParams(qkv.data_ptr(), None, kc.data_ptr(), vc.data_ptr(), ctxo.data_ptr(), seq.data_ptr(), inl.data_ptr(),
                        pad.data_ptr(), fin.data_ptr(), step.data_ptr(), part.data_ptr(), cnt.data_ptr(), batch, H, dh,
                        MODEL["rotary_embedding_dim"], max_len, 16, splits, 1.0 / math.sqrt(dh))
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        capi.check(lib.ftcf_mmha_decode(p, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        capi.check(lib.ftcf_mmha_decode(p, st))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    nbytes = batch * H * ctx * dh * 2 * 2
    return {"avg_launch_us": us, "bytes_per_launch": nbytes, "gbs": nbytes / (us * 1e-6) / 1e9, "splits": splits}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fastertransformer4codefuse_b200 import capi, weights as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    capi.check(capi.load().ftcf_device_check())
    # the drop-in itself: the compiled pybind11 module the unchanged driver imports from --lib_path (codefuse_example.py:468-470)
    if capi.LIB_DIR not in sys.path:
        sys.path.append(capi.LIB_DIR)
    import libth_gptneox
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t = world
    comm = dist.group.WORLD if world > 1 else None
    peak, peak_src = hbm_peak()
    tensor_peak = None
    try:
        tensor_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:  # noqa: BLE001
        pass

    def make_op(int8_mode):
        cfg = W.NeoXConfig(start_id=100000, end_id=MODEL["vocab_size"] - 1, use_gptj_residual=True, **MODEL)
        rw = W.make_synthetic_fast(cfg, t, rank, int8_mode, dev, seed=0)
        # the end_id row of the LM head is zeroed and its logit can never win: exactly out_len tokens are generated
        rw.w[12 * cfg.layer_num + 3][cfg.end_id].zero_()
        w, q, s = rw.lists()
        op = libth_gptneox.GptNeoXOp(comm, rank, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size,
                                     cfg.rotary_embedding_dim, cfg.start_id, cfg.end_id, t, 1, int8_mode, 4096, True, w, q, s)
        return cfg, rw, op

    def fwd(op, ids, lens, out):   # positional, as codefuse_example.py:575-589
        return op.forward(ids, lens, out, 1, None, None, None, None, None, None, None, None, None, 0, None)

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

    def workload(op, cfg, wl, steps, warmup):
        B, S, out = wl["batch"], wl["in_len"], wl["out_len"]
        g = np.random.default_rng(1234)
        ids_host = torch.from_numpy(g.integers(0, cfg.vocab_size - 2, size=(B, S)).astype(np.int32)).pin_memory()
        lens_host = torch.full((B,), S, dtype=torch.int32).pin_memory()
        ids_dev, lens_dev = ids_host.to(dev), lens_host.to(dev)
        op.set_option("step_timing", 1)
        for _ in range(warmup):
            fwd(op, ids_dev, lens_dev, out)
        # ---- device-resident inputs
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches, prefill_ms, decode_ms, step_ms = 0, [], [], []
        for _ in range(steps):
            res = fwd(op, ids_dev, lens_dev, out)
            st = op.last_stats()
            launches += st["kernel_launches"]
            prefill_ms.append(st["prefill_ms"])
            decode_ms.append(st["decode_ms"])
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
            r = fwd(op, a, b, out)
            out_host = r[0].cpu()
            len_host = r[1].cpu()
        e3.record()
        barrier()
        e2e_ms = max_over_ranks(e2.elapsed_time(e3))
        toks = B * out * steps
        p50 = statistics.median(step_ms) if step_ms else None
        wbytes = 1 if wl["int8"] else 2
        step_bytes = algorithmic_bytes_per_step(t, B, S + out / 2, wbytes)
        pre = statistics.median(prefill_ms)
        res = {"workload": wl["name"] + f", tensor_para {t}", "tokens_per_s": toks / (total_ms * 1e-3), "ms_per_request": total_ms / steps,
               "prefill_ms": pre, "decode_ms": statistics.median(decode_ms),
               "decode_tokens_per_s": B * (out - 1) / (statistics.median(decode_ms) * 1e-3), "p50_token_ms": p50,
               "step_roofline_frac": (step_bytes / (p50 * 1e-3) / 1e9 / peak) if p50 else None,
               "algorithmic_bytes_per_step_per_gpu": step_bytes, "launches": launches,
               "e2e_tokens_per_s": toks / (e2e_ms * 1e-3), "h2d": ids_host.numel() * 4 + lens_host.numel() * 4,
               "d2h": out_host.numel() * 4 + len_host.numel() * 4}
        if tensor_peak and pre > 0:
            res["prefill_tensor_frac"] = prefill_flops(t, B, S) / (pre * 1e-3) / 1e12 / tensor_peak
        return res

    def sub(r):
        keys = ("workload", "tokens_per_s", "e2e_tokens_per_s", "decode_tokens_per_s", "p50_token_ms", "prefill_ms", "step_roofline_frac",
                "prefill_tensor_frac")
        return {k: r[k] for k in keys if k in r}

    t_w0 = time.perf_counter()
    cfg, rw, op = make_op(1)
    weight_s = time.perf_counter() - t_w0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main = workload(op, cfg, C3, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    gemm = time_gemm_kernel(rw, t, torch, capi) if rank == 0 else None
    mmha = time_mmha_kernel(t, 32, C5["in_len"] + C5["out_len"] // 2, torch, capi) if rank == 0 else None
    # The other BASELINE configurations are sub-objects of the line; a failure in one of them must not cost the headline
    # (argument / capacity errors are raised on every rank before any collective, so the ranks stay in step).
    subs = {}

    def extra(name, fn):
        try:
            subs[name] = sub(fn())
        except Exception as exc:  # noqa: BLE001
            subs[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if not args.skip_extra:
        if world == 2:
            extra("config4_batch8", lambda: workload(op, cfg, C4, 1, 1))
        extra("config5_batch32_2048_512", lambda: workload(op, cfg, C5, 1, 1))
    if world == 1 and not args.skip_extra:
        del op, rw
        torch.cuda.empty_cache()

        def fp16_config():
            cfg2, rw2, op2 = make_op(0)
            return workload(op2, cfg2, C2, 2, 3)
        extra("config2_fp16", fp16_config)
        torch.cuda.empty_cache()
    cb = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cb = cpu_baseline()
    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    traffic, traffic_src = ncu_traffic_per_launch() if world == 1 else (None, "captured at tensor_para 1 only")
    line = {
        "metric": METRIC, "value": main["tokens_per_s"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_request"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": (main["tokens_per_s"] / PUBLISHED_TOKENS_PER_S[world]) if world in PUBLISHED_TOKENS_PER_S else None,
        "dtype": "int8 weights x fp16 activations, fp32 accumulate", "data": "synthetic",
        "config": {"workload": f"CodeFuse-13B weight-only int8 (int8_mode=1), batch 1, 1024 in / 512 out, greedy, tensor_para={world}",
                   "l2": "inputs larger than L2 (12.6 GB of weights streamed per token, no flush needed)",
                   "parallelism": f"tp{world}", "weights_init_s": round(weight_s, 1),
                   "api": "libth_gptneox.GptNeoXOp (compiled pybind11 drop-in) for every request of this line"},
        "decode": {"tokens_per_s": main["decode_tokens_per_s"], "p50_token_ms": main["p50_token_ms"], "prefill_ms": main["prefill_ms"],
                   "decode_ms": main["decode_ms"], "step_roofline_frac": main["step_roofline_frac"],
                   "prefill_tensor_frac": main.get("prefill_tensor_frac"),
                   "algorithmic_bytes_per_step_per_gpu": main["algorithmic_bytes_per_step_per_gpu"]},
        "e2e": {"value": main["e2e_tokens_per_s"], "unit": "tokens/s", "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"]},
        "gpu_launches": main["launches"],
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": gemm["gbs"], "peak": peak, "unit": "GB/s", "frac": gemm["gbs"] / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "weight-only INT8 decode GEMM (m = 1): all 160 layer GEMMs of one token, timed alone",
                     "avg_launch_us": gemm["avg_launch_us"], "bytes_per_launch": gemm["bytes_per_launch"], "peak_source": peak_src},
        "roofline_batch32": {"bound": "hbm", "achieved": mmha["gbs"], "peak": peak, "unit": "GB/s", "frac": mmha["gbs"] / peak,
                             "traffic": None,
                             "kernel": f"decode attention, batch 32, context {C5['in_len'] + C5['out_len'] // 2}, {MODEL['head_num'] // t} heads "
                                       f"(dominant kernel of the batch-32 step), {mmha['splits']} KV splits, timed alone",
                             "avg_launch_us": mmha["avg_launch_us"], "bytes_per_launch": mmha["bytes_per_launch"]},
    }
    line.update(subs)
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
    ap.add_argument("--skip-extra", action="store_true", help="only the headline workload (no configs[1] / [3] / [4] sub-objects)")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
