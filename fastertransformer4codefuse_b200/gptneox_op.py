"""Python mirror of the reference operator `libth_gptneox.GptNeoXOp` over the C ABI.

Same constructor and `forward` arguments, same outputs and error behaviour as the pybind11 class
th_op/gptneox/GptNeoXOp.cc:190-212 (constructor :25-106, forward :113-185; call site
examples/pytorch/codefuse/codefuse_example.py:533-536,575-589).  The compiled shim in csrc/binding/ does exactly
this in C++; this module is what the tests and bench.py drive, so they read like the reference's usage.
torch is used for device memory and the process group only.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional

import torch

from . import capi


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


class GptNeoXOp:
    def __init__(self, comm, rank: int, head_num: int, size_per_head: int, inter_size: int, layer_num: int, vocab_size: int,
                 rotary_embedding_dim: int, start_id: int, end_id: int, tensor_para_size: int, pipeline_para_size: int,
                 int8_mode: int, max_seq_len: int, use_gptj_residual: bool, weights: List[torch.Tensor],
                 int8_weights: List[torch.Tensor], scale: List[torch.Tensor], int8_layout: int = 0):
        lib = capi.load()
        if pipeline_para_size != 1:
            raise RuntimeError("pipeline_para_size must be 1 (the CodeFuse driver fixes it, codefuse_example.py:647)")
        if len(weights) != 12 * layer_num + 4:
            raise RuntimeError(f"expected {12 * layer_num + 4} weight tensors, got {len(weights)}")
        st = weights[0].dtype
        if st not in (torch.float16, torch.float32):                 # dispatch on weights[0], GptNeoXOp.cc:46,56-105
            raise RuntimeError("Wrong tensor type: weights must be fp16 or fp32")
        for t in weights:                                            # CHECK_INPUT, GptNeoXOp.cc:52-54
            if t.numel() and (not t.is_cuda or not t.is_contiguous() or t.dtype != st):
                raise RuntimeError("weights must be contiguous CUDA tensors of one dtype")
        if st == torch.float32:
            # fp32 checkpoints are rounded to fp16 once at load: the engine computes in fp16 with fp32 accumulation only
            weights = [t.half().contiguous() if t.numel() else t for t in weights]
            scale = [t.half().contiguous() if t.numel() and t.dtype == torch.float32 else t for t in scale]
        self.weights, self.int8_weights, self.scale = list(weights), list(int8_weights), list(scale)   # keep alive
        self.tensor_para_size = tensor_para_size
        self.end_id = end_id
        self.vocab_size = vocab_size
        cfg = capi.GptNeoXConfig(head_num, size_per_head, inter_size, layer_num, vocab_size, rotary_embedding_dim, start_id,
                                 end_id, tensor_para_size, rank % tensor_para_size, int8_mode, 1 if use_gptj_residual else 0,
                                 1e-5, int8_layout)
        warr = (C.c_void_p * len(weights))(*[_ptr(t) for t in weights])
        n8 = len(int8_weights) if int8_mode == 1 else 0
        qarr = (C.c_void_p * max(n8, 1))(*[_ptr(t) for t in int8_weights[:n8]])
        sarr = (C.c_void_p * max(n8, 1))(*[_ptr(t) for t in scale[:n8]])
        uid = None
        if tensor_para_size > 1:
            uid = self._exchange_nccl_id(lib, comm, rank)
        self.stream = torch.cuda.current_stream().cuda_stream       # captured at construction, GptNeoXOp.h:180
        handle = C.c_void_p()
        capi.check(lib.ftcf_gptneox_create(C.byref(handle), C.byref(cfg), warr, len(weights), qarr if n8 else None,
                                           sarr if n8 else None, n8, uid, self.stream))
        self._h = handle
        self._lib = lib
        self.last_stats = None
        # experiment hook: FTCF_OPTIONS="two_branch=0,cuda_graph=1"
        import os
        for item in filter(None, os.environ.get("FTCF_OPTIONS", "").split(",")):
            key, _, val = item.partition("=")
            self.set_option(key.strip(), int(val))

    @staticmethod
    def _exchange_nccl_id(lib, comm, rank):
        """Rank 0 makes the ncclUniqueId, everyone gets it through the torch process group
        (replaces nccl_inherit::ftNcclInitialize, th_op/gptneox/utils/nccl_inherit_utils.cc:8-68)."""
        import torch.distributed as dist
        buf = (C.c_char * 128)()
        group_rank = dist.get_rank(comm) if comm is not None else rank
        if group_rank == 0:
            capi.check(lib.ftcf_nccl_unique_id(buf))
        obj = [bytes(buf) if group_rank == 0 else None]
        src = dist.get_global_rank(comm, 0) if comm is not None else 0
        dist.broadcast_object_list(obj, src=src, group=comm)
        return C.create_string_buffer(obj[0], 128)

    def set_option(self, name: str, value: int) -> None:
        capi.check(self._lib.ftcf_gptneox_set_option(self._h, name.encode(), int(value)))

    def last_step_ms(self, n: int = 8192):
        arr = (C.c_float * n)()
        c = self._lib.ftcf_gptneox_last_step_ms(self._h, arr, n)
        return [arr[i] for i in range(c)]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.ftcf_gptneox_destroy(h)
            self._h = None

    def forward(self, input_ids: torch.Tensor, input_lengths: torch.Tensor, output_len: int, beam_width: Optional[int] = None,
                top_k: Optional[torch.Tensor] = None, top_p: Optional[torch.Tensor] = None,
                beam_search_diversity_rate: Optional[torch.Tensor] = None, temperature: Optional[torch.Tensor] = None,
                len_penalty: Optional[torch.Tensor] = None, repetition_penalty: Optional[torch.Tensor] = None,
                random_seed: Optional[torch.Tensor] = None, stop_words_list: Optional[torch.Tensor] = None,
                optional_last_tokens: Optional[torch.Tensor] = None, return_cum_log_probs: Optional[int] = None,
                callback: Optional[Callable[[dict], None]] = None, logits_trace: Optional[torch.Tensor] = None):
        # argument checks as GptNeoXOp.cc:133-145
        for name, t in (("input_ids", input_ids), ("input_lengths", input_lengths)):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.int32):
                raise RuntimeError(f"{name} must be a contiguous CUDA int32 tensor")
        if input_ids.dim() != 2:
            raise RuntimeError("input_ids must be [batch, max_input_length]")
        bw = 1 if beam_width is None else max(1, int(beam_width))
        rcl = 0 if return_cum_log_probs is None else int(return_cum_log_probs)
        if rcl not in (0, 1):                                        # GptNeoXOp.cc:143-145
            raise RuntimeError("return_cum_log_probs should be 0 (no return cum_log_probs), "
                               "1 (the cumulative log probs of generated sequences)")
        B, S = input_ids.shape
        total = S + int(output_len)
        dev = input_ids.device
        out_ids = torch.empty((B, bw, total), dtype=torch.int32, device=dev)
        seq_lens = torch.empty((B, bw), dtype=torch.int32, device=dev)
        cum = torch.empty((B, bw), dtype=torch.float32, device=dev) if rcl > 0 else None

        keep = []

        def host(t, dt):
            if t is None:
                return None, 0
            tt = t.detach().to("cpu", dt).contiguous().reshape(-1)
            keep.append(tt)
            return tt.data_ptr(), tt.numel()

        rq = capi.GptNeoXRequest()
        rq.input_ids, rq.input_lengths = input_ids.data_ptr(), input_lengths.data_ptr()
        rq.batch, rq.max_input_len, rq.output_len, rq.beam_width = B, S, int(output_len), bw
        rq.top_k_host, rq.n_top_k = host(top_k, torch.int32)
        rq.top_p_host, rq.n_top_p = host(top_p, torch.float32)
        rq.temperature_host, rq.n_temperature = host(temperature, torch.float32)
        rq.repetition_penalty_host, rq.n_repetition_penalty = host(repetition_penalty, torch.float32)
        rq.random_seed_host, rq.n_random_seed = host(random_seed, torch.int64)
        rq.beam_search_diversity_rate_host, rq.n_beam_search_diversity_rate = host(beam_search_diversity_rate, torch.float32)
        rq.len_penalty_host, rq.n_len_penalty = host(len_penalty, torch.float32)
        if stop_words_list is not None:
            if not (stop_words_list.is_cuda and stop_words_list.dtype == torch.int32 and stop_words_list.dim() == 3):
                raise RuntimeError("stop_words_list must be a CUDA int32 tensor [batch, 2, n]")
            sw = stop_words_list.contiguous()
            keep.append(sw)
            rq.stop_words, rq.n_stop = sw.data_ptr(), sw.shape[2]
        if optional_last_tokens is not None:
            if not (optional_last_tokens.is_cuda and optional_last_tokens.dtype == torch.int32 and optional_last_tokens.dim() == 2):
                raise RuntimeError("optional_last_tokens must be a CUDA int32 tensor [batch, n]")
            ol = optional_last_tokens.contiguous()
            keep.append(ol)
            rq.optional_last_tokens, rq.n_last = ol.data_ptr(), ol.shape[1]
        rq.return_cum_log_probs = rcl
        rq.output_ids, rq.sequence_lengths = out_ids.data_ptr(), seq_lens.data_ptr()
        rq.cum_log_probs = cum.data_ptr() if cum is not None else None
        if logits_trace is not None:
            rq.logits_trace, rq.logits_trace_steps = logits_trace.data_ptr(), logits_trace.shape[0]

        cb_err = []
        if callback is not None:
            def _cb(_user, _step, toks, idxs, n):
                try:    # same message shape as th_op/gptneox/utils/pybind_callback_utils.cc:59-103
                    callback({"last_tokens": [[int(toks[b * bw + j]) for j in range(bw)] for b in range(n // bw)],
                              "idxs": [[int(idxs[b * bw + j]) for j in range(bw)] for b in range(n // bw)]})
                except BaseException as exc:   # noqa: BLE001 -- must not unwind through C
                    cb_err.append(exc)
            cfn = capi.TOKEN_CALLBACK(_cb)
            keep.append(cfn)
            rq.callback = cfn
        stats = capi.GptNeoXStats()
        capi.check(self._lib.ftcf_gptneox_forward(self._h, C.byref(rq), C.byref(stats)))
        if cb_err:
            raise cb_err[0]
        self.last_stats = {"steps": stats.steps, "prefill_ms": stats.prefill_ms, "decode_ms": stats.decode_ms,
                           "kernel_launches": stats.kernel_launches}
        res = [out_ids, seq_lens]
        if cum is not None:
            res.append(cum)
        return res
