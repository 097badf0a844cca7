"""CPU tests of the pybind11 shims that carry the reference's Python surface (libth_gptneox / libth_common):
they import under the reference's names from the --lib_path directory, expose the reference's callables and
reject bad arguments with RuntimeError, as th_op/gptneox/GptNeoXOp.cc:190-212 and
th_op/common/WeightOnlyQuantOps.cc:344-349 do.  No GPU work happens here."""
import os
import sys

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi, quant

LIB_DIR = capi.LIB_DIR


@pytest.fixture(scope="module")
def shims():
    if LIB_DIR not in sys.path:
        sys.path.append(LIB_DIR)            # what codefuse_example.py:468 does with --lib_path
    for name in ("libth_gptneox.so", "libth_common.so"):
        assert os.path.exists(os.path.join(LIB_DIR, name)), f"{name} missing: run __graft_entry__.build()"
    import libth_common
    import libth_gptneox
    return libth_gptneox, libth_common


def test_shim_names(shims):
    g, c = shims
    assert hasattr(g, "GptNeoXOp") and callable(g.GptNeoXOp.forward)
    assert callable(c.symmetric_quantize_last_axis_of_batched_matrix_int8)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_shim_quantiser_equals_c_abi(shims, dtype):
    _, c = shims
    w = (torch.randn(96, 48, generator=torch.Generator().manual_seed(3)) * 0.05).to(dtype)
    q, s = c.symmetric_quantize_last_axis_of_batched_matrix_int8(w)
    q2, s2 = quant.symmetric_quantize_last_axis_of_batched_matrix_int8(w)
    assert q.dtype == torch.int8 and q.shape == w.shape and s.dtype == dtype and tuple(s.shape) == (48,)
    assert torch.equal(q, q2) and torch.equal(s.view(torch.int16) if dtype != torch.float32 else s, s2.view(torch.int16) if dtype != torch.float32 else s2)
    w3 = torch.stack([w, w * 0.5])
    q3, s3 = c.symmetric_quantize_last_axis_of_batched_matrix_int8(w3)
    assert tuple(q3.shape) == (2, 96, 48) and tuple(s3.shape) == (2, 48)
    assert torch.equal(q3[0], q)


def test_shim_argument_errors(shims):
    g, c = shims
    with pytest.raises(RuntimeError):
        c.symmetric_quantize_last_axis_of_batched_matrix_int8(torch.zeros(4, dtype=torch.float16))            # 1-D
    with pytest.raises(RuntimeError):
        c.symmetric_quantize_last_axis_of_batched_matrix_int8(torch.zeros(4, 4, dtype=torch.int32))           # dtype
    with pytest.raises(RuntimeError, match="weight tensors"):
        g.GptNeoXOp(None, 0, 2, 64, 512, 1, 100, 32, 0, 1, 1, 1, 0, 64, True, [torch.zeros(1).half()], [], [])
    with pytest.raises(RuntimeError, match="pipeline_para_size"):
        g.GptNeoXOp(None, 0, 2, 64, 512, 1, 100, 32, 0, 1, 1, 2, 0, 64, True, [torch.zeros(1).half()] * 16, [], [])
    with pytest.raises(RuntimeError):   # CPU weights are rejected (CHECK_INPUT)
        g.GptNeoXOp(None, 0, 2, 64, 512, 1, 100, 32, 0, 1, 1, 1, 0, 64, True, [torch.zeros(8).half()] * 16, [], [])
