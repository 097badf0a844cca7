"""GPU parity of the tcgen05 decode GEMM (csrc/gemm_decode.cu, impl = 3: INT8 weights x fp16 activations at m <= 32 rows) through
the C ABI.  Reference = dequantise then fp32 matmul (torch), tolerance = the reference's own for this GEMM (rtol 1e-3 /
atol 2e-3, tests/gemm_dequantize/th_gemm_dequantize.py:111-116); the exact dequant round trip is checked with zero tolerance as
there (:22-51).  Also: k-split launches are deterministic and their ticket counters reset themselves, a chain of dependent
launches under programmatic dependent launch gives the results of the same chain run one launch at a time, and the fused
residual + LayerNorm prologue matches the unfused kernels."""
import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi
from fastertransformer4codefuse_b200 import weights as W
from oracle import gptneox_ref as R
from helpers import assert_close, stream

pytestmark = pytest.mark.gpu
DECODE = 3


def _gemm(lib, x, p, s, bias, m, n, k, act, impl=DECODE):
    y = torch.empty(m, n, dtype=torch.float16, device=x.device)
    capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), p.data_ptr(), s.data_ptr(), bias.data_ptr() if bias is not None else None,
                                   y.data_ptr(), m, n, k, act, impl, stream()))
    torch.cuda.synchronize()
    return y


def test_decode_exact_dequant_round_trip(lib, cuda):
    torch.manual_seed(734876213)
    k, n = 256, 512
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = W.quantize_on_device(w)
    x = torch.eye(k, dtype=torch.float16, device=cuda)
    for m0 in range(0, k, 32):
        y = _gemm(lib, x[m0:m0 + 32].contiguous(), p, s, None, 32, n, k, 0)
        ref = (q[m0:m0 + 32].float() * s.float()[None, :]).half()
        assert torch.equal(y, ref)


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 8, 9, 16, 17, 31, 32])
@pytest.mark.parametrize("n,k", [(1024, 4096), (5120, 640), (768, 3072), (1000, 128), (136, 256), (15360, 5120), (5120, 20480)])
def test_decode_grid(lib, cuda, m, n, k):
    torch.manual_seed(734876213 + m)
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = W.quantize_on_device(w)
    x = torch.randn(m, k, device=cuda).half()
    y = _gemm(lib, x, p, s, None, m, n, k, 0)
    ref = x.float() @ (q.float() * s.float()[None, :])
    assert_close(f"decode w8a16 m={m} n={n} k={k}", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=2e-3)
    y2 = _gemm(lib, x, p, s, None, m, n, k, 0)
    assert torch.equal(y, y2), "k-split reduction must be deterministic and its counters self-resetting"


@pytest.mark.parametrize("m", [1, 5, 32])
def test_decode_bias_gelu(lib, cuda, m):
    torch.manual_seed(11 + m)
    n, k = 2048, 1024
    w = (torch.randn(k, n, device=cuda) * 0.02).half()
    p, s, q = W.quantize_on_device(w)
    x = torch.randn(m, k, device=cuda).half()
    bias = (torch.randn(n, device=cuda) * 0.1).half()
    y = _gemm(lib, x, p, s, bias, m, n, k, 1)
    acc = x.float() @ (q.float() * s.float()[None, :]) + bias.float()
    ref = torch.nn.functional.gelu(acc, approximate="tanh")          # th_gemm_dequantize.py:61
    assert_close(f"decode gelu m={m}", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("target,min_kb", [(1, 8), (148, 8), (296, 2), (592, 1)])
def test_decode_split_choices_agree(lib, cuda, target, min_kb):
    """Whatever k-split the launch heuristic picks, the result stays inside the tolerance."""
    torch.manual_seed(5)
    m, n, k = 3, 640, 5120
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = W.quantize_on_device(w)
    x = torch.randn(m, k, device=cuda).half()
    try:
        capi.check(lib.ftcf_set_tunable(b"decode_target_ctas", target))
        capi.check(lib.ftcf_set_tunable(b"decode_min_kb", min_kb))
        y = _gemm(lib, x, p, s, None, m, n, k, 0)
    finally:
        lib.ftcf_set_tunable(b"decode_target_ctas", 296)
        lib.ftcf_set_tunable(b"decode_min_kb", 8)
    ref = x.float() @ (q.float() * s.float()[None, :])
    assert_close("decode split choice", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=2e-3)


def test_decode_chain_under_pdl(lib, cuda):
    """Ten dependent GEMMs back to back on one stream (each reads what the previous one wrote, every kernel launched with the
    programmatic-dependent-launch attribute): identical to the chain with a synchronize after every launch."""
    torch.manual_seed(3)
    m, h = 2, 1024
    ws = [W.quantize_on_device((torch.randn(h, h, device=cuda) * 0.03).half()) for _ in range(10)]
    x0 = torch.randn(m, h, device=cuda).half()

    def chain(sync):
        bufs = [x0.clone(), torch.empty_like(x0)]
        for i, (p, s, _) in enumerate(ws):
            capi.check(lib.ftcf_gemm_w8a16(bufs[i & 1].data_ptr(), p.data_ptr(), s.data_ptr(), None, bufs[(i + 1) & 1].data_ptr(), m, h, h, 0,
                                           DECODE, stream()))
            if sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        return bufs[len(ws) & 1].clone()

    a, b = chain(True), chain(False)
    assert torch.isfinite(a.float()).all()
    assert torch.equal(a, b)


def _ln_prologue(x, ffn, attn, bias, g, b, x_out):
    pro = capi.LnPrologue()
    pro.x, pro.gamma, pro.beta, pro.eps = x.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-5
    if ffn is not None:
        pro.add_ffn, pro.add_attn = ffn.data_ptr(), attn.data_ptr()
        pro.add_bias = bias.data_ptr() if bias is not None else None
    pro.x_out = x_out.data_ptr() if x_out is not None else None
    return pro


@pytest.mark.parametrize("m", [1, 2, 3, 4])
@pytest.mark.parametrize("with_res", [False, True])
@pytest.mark.parametrize("n,k,hint", [(1024, 5120, 0), (300, 256, 0), (15360, 5120, 148), (2048, 20480, 0)])
def test_decode_fused_residual_layernorm_prologue(lib, cuda, m, with_res, n, k, hint):
    """ftcf_gemm_w8a16_ln on the tcgen05 decode kernel == residual add (exact fp16 adds) -> LayerNorm -> INT8 GEMM."""
    torch.manual_seed(100 * m + n + int(with_res))
    w = (torch.randn(k, n, device=cuda) * 0.02).half()
    p, s, q = W.quantize_on_device(w)
    x, ffn, attn = [torch.randn(m, k, device=cuda).half() for _ in range(3)]
    bias = (0.1 * torch.randn(k, device=cuda)).half()
    g = (1 + 0.1 * torch.randn(k, device=cuda)).half()
    b = (0.1 * torch.randn(k, device=cuda)).half()
    obias = (0.1 * torch.randn(n, device=cuda)).half()
    x_out = torch.zeros_like(x)
    y = torch.empty(m, n, dtype=torch.float16, device=cuda)
    pro = _ln_prologue(x, ffn if with_res else None, attn, bias, g, b, x_out if with_res else None)
    pro.cta_hint = hint
    capi.check(lib.ftcf_gemm_w8a16_ln(pro, p.data_ptr(), s.data_ptr(), obias.data_ptr(), y.data_ptr(), m, n, k, 1, stream()))
    torch.cuda.synchronize()
    r = x.float().cpu()
    if with_res:
        r = R.h(R.h(R.h(ffn.float().cpu() + attn.float().cpu()) + bias.float().cpu()) + r)
        assert torch.equal(x_out.float().cpu(), r)
    a = R.layernorm_ref(r, g.cpu(), b.cpu(), 1e-5)
    a_dev = a.half().to(cuda)
    y2 = _gemm(lib, a_dev, p, s, obias, m, n, k, 1)
    ref = R.gelu_f32(a.float() @ (q.float().cpu() * s.float().cpu()[None, :]) + obias.float().cpu())
    assert_close("fused-prologue decode gemm vs unfused", y.float().cpu(), y2.float().cpu(), rtol=4e-3, atol=4e-3)
    assert_close("fused-prologue decode gemm vs oracle", y.float().cpu(), ref, rtol=4e-3, atol=6e-3)


def test_decode_rejects_unsupported_k(lib, cuda):
    x = torch.zeros(1, 192, dtype=torch.float16, device=cuda)
    w = torch.zeros(64, 192, dtype=torch.uint8, device=cuda)
    sc = torch.ones(64, dtype=torch.float16, device=cuda)
    y = torch.empty(1, 64, dtype=torch.float16, device=cuda)
    assert lib.ftcf_gemm_w8a16(x.data_ptr(), w.data_ptr(), sc.data_ptr(), None, y.data_ptr(), 1, 64, 192, 0, DECODE, stream()) == 4
    # (the streaming kernel has the same 128-byte K granularity: auto mode reports the same error instead of computing garbage)
    assert lib.ftcf_gemm_w8a16(x.data_ptr(), w.data_ptr(), sc.data_ptr(), None, y.data_ptr(), 1, 64, 192, 0, 0, stream()) == 4
    torch.cuda.synchronize()
