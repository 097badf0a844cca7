"""`libth_common.symmetric_quantize_last_axis_of_batched_matrix_int8` over the C ABI.

Mirrors the reference binding th_op/common/WeightOnlyQuantOps.cc:140-233 (call sites
examples/pytorch/codefuse/codefuse_example.py:392-399, quant_and_save.py:46-47): CPU tensor [k, n] or [e, k, n] in
fp32 / fp16 / bf16 -> [int8 tensor of the same shape holding the *processed* bytes, scales [n] / [e, n] in the
weight dtype].  "Processed" here is the B200 layout (W^T, k contiguous, q + 128 as uint8), not the reference's sm80
interleave; see INTEGRATION.md.
"""
from __future__ import annotations

import torch

from . import capi

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def symmetric_quantize_last_axis_of_batched_matrix_int8(weight: torch.Tensor, return_unprocessed: bool = False):
    if weight.device.type != "cpu":
        raise RuntimeError("weight must be a CPU tensor")            # CHECK_CPU, WeightOnlyQuantOps.cc:146
    if weight.dtype not in _DTYPES:
        raise RuntimeError("Invalid datatype. Weight must be FP16, BF16 or FP32")   # :216-220
    if weight.dim() not in (2, 3):
        raise RuntimeError("Invalid dim. The dim of weight should be 2 or 3")       # :149
    w = weight.contiguous()
    e = 1 if w.dim() == 2 else w.shape[0]
    k, n = w.shape[-2], w.shape[-1]
    processed = torch.empty(w.shape, dtype=torch.int8)
    scales = torch.empty((n,) if w.dim() == 2 else (e, n), dtype=w.dtype)
    unprocessed = torch.empty(w.shape, dtype=torch.int8) if return_unprocessed else None
    lib = capi.load()
    capi.check(lib.ftcf_symmetric_quantize_int8_host(w.data_ptr(), _DTYPES[w.dtype], e, k, n, processed.data_ptr(),
                                                     unprocessed.data_ptr() if unprocessed is not None else None,
                                                     scales.data_ptr()))
    if return_unprocessed:
        return [processed, scales, unprocessed]
    return [processed, scales]


def ampere_layout_to_b200(processed: torch.Tensor, k: int, n: int) -> torch.Tensor:
    """Bytes of a reference-made `*.q.bin` (sm80 layout) -> B200 layout."""
    src = processed.contiguous().view(torch.int8).reshape(-1)
    out = torch.empty(k * n, dtype=torch.int8)
    capi.check(capi.load().ftcf_int8_ampere_to_b200_host(src.data_ptr(), k, n, out.data_ptr()))
    return out.reshape(k, n)
