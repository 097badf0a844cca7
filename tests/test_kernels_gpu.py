"""GPU parity tests of the individual kernels, called through the C ABI, against the oracle's primitives.

Tolerances: the INT8 GEMM uses the reference's own test tolerance rtol 1e-3 / atol 2e-3 on fp16 outputs with
weights N(0, 0.002)-scale quantised and activations N(0, 1) (tests/gemm_dequantize/th_gemm_dequantize.py:111-116);
the exact dequant round trip (identity activations) is checked with zero tolerance as there (:22-51).
"""
import math

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi
from fastertransformer4codefuse_b200 import weights as W
from oracle import gptneox_ref as R
from helpers import assert_close, stream

pytestmark = pytest.mark.gpu


def _quant(w_kn):
    p, s, q = W.quantize_on_device(w_kn)
    return p, s, q


def _gemm_w8(lib, x, p, s, bias, m, n, k, act, impl=0):
    y = torch.empty(m, n, dtype=torch.float16, device=x.device)
    capi.check(lib.ftcf_gemm_w8a16(x.data_ptr(), p.data_ptr(), s.data_ptr(), bias.data_ptr() if bias is not None else None,
                                   y.data_ptr(), m, n, k, act, impl, stream()))
    torch.cuda.synchronize()
    return y


def test_w8a16_exact_dequant_round_trip(lib, cuda):
    # identity activations reproduce the dequantised weights exactly (th_gemm_dequantize.py:22-51, atol = rtol = 0)
    torch.manual_seed(734876213)
    k, n = 256, 512
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = _quant(w)
    x = torch.eye(k, dtype=torch.float16, device=cuda)
    for m0 in range(0, k, 32):
        y = _gemm_w8(lib, x[m0:m0 + 32].contiguous(), p, s, None, 32, n, k, 0, impl=1)
        ref = (q[m0:m0 + 32].float() * s.float()[None, :]).half()
        assert torch.equal(y, ref)


@pytest.mark.parametrize("m", [1, 2, 3, 8, 9, 17, 32, 33, 66, 125])
@pytest.mark.parametrize("n,k", [(1024, 4096), (5120, 640), (768, 3072)])
def test_w8a16_skinny_grid(lib, cuda, m, n, k):
    torch.manual_seed(734876213 + m)
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = _quant(w)
    x = torch.randn(m, k, device=cuda).half()
    y = _gemm_w8(lib, x, p, s, None, m, n, k, 0, impl=1)
    ref = x.float() @ (q.float() * s.float()[None, :])
    assert_close(f"w8a16 m={m} n={n} k={k}", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("m", [1, 5, 32])
def test_w8a16_bias_gelu(lib, cuda, m):
    torch.manual_seed(11 + m)
    n, k = 2048, 1024
    w = (torch.randn(k, n, device=cuda) * 0.02).half()
    p, s, q = _quant(w)
    x = torch.randn(m, k, device=cuda).half()
    bias = (torch.randn(n, device=cuda) * 0.1).half()
    y = _gemm_w8(lib, x, p, s, bias, m, n, k, 1, impl=1)
    acc = x.float() @ (q.float() * s.float()[None, :]) + bias.float()
    ref = torch.nn.functional.gelu(acc, approximate="tanh")          # th_gemm_dequantize.py:61
    assert_close(f"w8a16 gelu m={m}", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("m", [1, 4, 32, 40])
@pytest.mark.parametrize("out_f32", [0, 1])
def test_f16_gemm(lib, cuda, m, out_f32):
    torch.manual_seed(5 + m)
    n, k = 1008, 768          # n not a multiple of 32: exercises the row guard
    w_nk = (torch.randn(n, k, device=cuda) * 0.02).half()
    x = torch.randn(m, k, device=cuda).half()
    y = torch.empty(m, n, dtype=torch.float32 if out_f32 else torch.float16, device=cuda)
    capi.check(lib.ftcf_gemm_f16(x.data_ptr(), w_nk.data_ptr(), None, y.data_ptr(), m, n, k, n, 0, out_f32, 1, stream()))
    torch.cuda.synchronize()
    ref = x.float() @ w_nk.float().t()
    if out_f32:
        assert_close("f16 gemm fp32 out", y.cpu(), ref.cpu(), rtol=1e-4, atol=1e-4)
    else:
        assert_close("f16 gemm fp16 out", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=1e-3)


def test_f16_gemm_bias_gelu_matches_reference_rounding(lib, cuda):
    torch.manual_seed(3)
    m, n, k = 7, 512, 256
    w_nk = (torch.randn(n, k, device=cuda) * 0.05).half()
    x = torch.randn(m, k, device=cuda).half()
    bias = (torch.randn(n, device=cuda) * 0.1).half()
    y = torch.empty(m, n, dtype=torch.float16, device=cuda)
    capi.check(lib.ftcf_gemm_f16(x.data_ptr(), w_nk.data_ptr(), bias.data_ptr(), y.data_ptr(), m, n, k, n, 1, 0, 1, stream()))
    torch.cuda.synchronize()
    acc = R.h((x.float() @ w_nk.float().t()).cpu())
    ref = R.gelu_half2(R.h(acc + bias.float().cpu()))
    assert_close("f16 gemm bias gelu", y.float().cpu(), ref, rtol=2e-3, atol=1e-3)


def test_transpose_f16(lib, cuda):
    torch.manual_seed(0)
    k, n = 130, 71
    a = torch.randn(k, n, device=cuda).half()
    out = torch.empty(n, k, dtype=torch.float16, device=cuda)
    capi.check(lib.ftcf_transpose_f16(a.data_ptr(), out.data_ptr(), k, n, stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, a.t().contiguous())


@pytest.mark.parametrize("m,n", [(1, 768), (5, 5120), (3, 256), (2, 8192)])
def test_layernorm(lib, cuda, m, n):
    torch.manual_seed(m * 7 + n)
    x = (torch.randn(m, n, device=cuda) * 2 + 0.3).half()
    g = (1 + 0.1 * torch.randn(n, device=cuda)).half()
    b = (0.1 * torch.randn(n, device=cuda)).half()
    y = torch.empty_like(x)
    capi.check(lib.ftcf_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), m, n, 1e-5, stream()))
    torch.cuda.synchronize()
    ref = R.layernorm_ref(x.float().cpu(), g.cpu(), b.cpu(), 1e-5)
    # identical arithmetic except the fp32 reduction order: allow one fp16 ulp on a tiny fraction of elements
    assert_close("layernorm", y.float().cpu(), ref, rtol=2e-3, atol=2e-3)
    frac_exact = (y.float().cpu() == ref).float().mean().item()
    assert frac_exact > 0.97, frac_exact


def test_add_bias_residual_layernorm(lib, cuda):
    torch.manual_seed(1)
    m, n = 4, 1024
    x = torch.randn(m, n, device=cuda).half()
    a = torch.randn(m, n, device=cuda).half()
    bias = (0.1 * torch.randn(n, device=cuda)).half()
    g = (1 + 0.1 * torch.randn(n, device=cuda)).half()
    b = (0.1 * torch.randn(n, device=cuda)).half()
    r = torch.empty_like(x)
    y = torch.empty_like(x)
    capi.check(lib.ftcf_add_bias_residual_layernorm(x.data_ptr(), a.data_ptr(), bias.data_ptr(), r.data_ptr(), g.data_ptr(),
                                                    b.data_ptr(), y.data_ptr(), m, n, 1e-5, stream()))
    torch.cuda.synchronize()
    r_ref = R.h((bias.float().cpu() + x.float().cpu()) + a.float().cpu())
    assert torch.equal(r.float().cpu(), r_ref)
    y_ref = R.layernorm_ref(r_ref, g.cpu(), b.cpu(), 1e-5)
    assert_close("pre-LN", y.float().cpu(), y_ref, rtol=4e-3, atol=4e-3)


@pytest.mark.parametrize("tp", [1, 2, 8])
def test_residual_parallel(lib, cuda, tp):
    torch.manual_seed(tp)
    m, n = 3, 768
    x, ffn, attn = [torch.randn(m, n, device=cuda).half() for _ in range(3)]
    bias = (0.1 * torch.randn(n, device=cuda)).half()
    out = torch.empty_like(x)
    capi.check(lib.ftcf_add_bias_attn_ffn_residual(out.data_ptr(), ffn.data_ptr(), attn.data_ptr(), x.data_ptr(), bias.data_ptr(),
                                                   m, n, tp, stream()))
    torch.cuda.synchronize()
    xs = R.h(x.float().cpu() / tp) if tp > 1 else x.float().cpu()
    ref = R.h(R.h(R.h(ffn.float().cpu() + attn.float().cpu()) + bias.float().cpu()) + xs)
    assert torch.equal(out.float().cpu(), ref)


def test_residual_sequential(lib, cuda):
    torch.manual_seed(9)
    m, n = 3, 768
    x, y = [torch.randn(m, n, device=cuda).half() for _ in range(2)]
    bias = (0.1 * torch.randn(n, device=cuda)).half()
    out = torch.empty_like(x)
    capi.check(lib.ftcf_add_bias_residual(out.data_ptr(), y.data_ptr(), x.data_ptr(), bias.data_ptr(), m, n, stream()))
    torch.cuda.synchronize()
    ref = R.h(R.h(y.float().cpu() + x.float().cpu()) + bias.float().cpu())
    assert torch.equal(out.float().cpu(), ref)


def test_embedding(lib, cuda):
    torch.manual_seed(2)
    v, n, m = 100, 256, 9
    table = torch.randn(v, n, device=cuda).half()
    ids = torch.randint(0, v, (m,), device=cuda, dtype=torch.int32)
    out = torch.empty(m, n, dtype=torch.float16, device=cuda)
    capi.check(lib.ftcf_embedding_lookup(out.data_ptr(), table.data_ptr(), ids.data_ptr(), m, n, v, stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, table[ids.long()])


def _mmha_oracle(qkv, bias, kc, vc, tl, in_len, max_in, pad, step, H, Dh, rot):
    """One decode-attention call restated with the oracle primitives (gptneox_ref.GptNeoXRef.forward, dec_attn)."""
    B = qkv.shape[0]
    hl = H * Dh
    x = R.h(qkv + bias)
    q, k, v = [z.reshape(B, H, Dh) for z in x.split(hl, dim=-1)]
    pos = torch.tensor([(step - 1) - int(pad[b]) for b in range(B)])
    cos, sin = R.rotary_coef(pos, rot)
    q = R.apply_rotary_neox(q, cos[:, None, :], sin[:, None, :], rot)
    k = R.apply_rotary_neox(k, cos[:, None, :], sin[:, None, :], rot)
    ctx = torch.zeros(B, H, Dh)
    for b in range(B):
        t = int(tl[b])
        kc[b, :, t] = k[b]
        vc[b, :, t] = v[b]
        keys, vals = kc[b, :, :t + 1], vc[b, :, :t + 1]
        sc = (keys @ q[b][:, :, None]).squeeze(-1) / math.sqrt(Dh)
        mk = torch.zeros(t + 1, dtype=torch.bool)
        mk[int(in_len[b]):max_in] = True
        mk = mk[:t + 1]
        scm = sc.masked_fill(mk[None, :], float("-inf"))
        mx = scm.max(dim=-1, keepdim=True).values
        e = torch.exp(sc - mx).masked_fill(mk[None, :], 0.0)
        p = R.h(e * (1.0 / (e.sum(-1, keepdim=True) + 1e-6)))
        ctx[b] = R.h((p[:, None, :] @ vals).squeeze(1))
    return ctx.reshape(B, hl), kc, vc


@pytest.mark.parametrize("Dh,rot", [(64, 16), (128, 128), (128, 32)])
@pytest.mark.parametrize("splits", [1, 3, 8])
def test_mmha_decode(lib, cuda, Dh, rot, splits):
    torch.manual_seed(Dh + splits)
    B, H, max_in, max_len = 3, 4, 40, 96
    in_len = torch.tensor([40, 17, 33], dtype=torch.int32)
    step = 61                                       # 21 tokens already generated
    tl = torch.full((B,), step - 1, dtype=torch.int32)
    pad = max_in - in_len
    qkv = torch.randn(B, 3 * H * Dh).half().float()
    bias = (0.1 * torch.randn(3 * H * Dh)).half().float()
    kc = torch.randn(B, H, max_len, Dh).half().float()
    vc = torch.randn(B, H, max_len, Dh).half().float()
    ref, kc_ref, vc_ref = _mmha_oracle(qkv, bias, kc.clone(), vc.clone(), tl, in_len, max_in, pad, step, H, Dh, rot)

    d = lambda t, dt: t.to(cuda, dt).contiguous()
    qkv_d, bias_d, kc_d, vc_d = d(qkv, torch.float16), d(bias, torch.float16), d(kc, torch.float16), d(vc, torch.float16)
    ctx = torch.zeros(B, H * Dh, dtype=torch.float16, device=cuda)
    tl_d, in_d, pad_d = d(tl, torch.int32), d(in_len, torch.int32), d(pad, torch.int32)
    fin = torch.zeros(B, dtype=torch.uint8, device=cuda)
    step_d = torch.tensor([step], dtype=torch.int32, device=cuda)
    part = torch.zeros(B * H * splits * (Dh + 2), dtype=torch.float32, device=cuda)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    p = capi.MmhaParams(qkv_d.data_ptr(), bias_d.data_ptr(), kc_d.data_ptr(), vc_d.data_ptr(), ctx.data_ptr(), tl_d.data_ptr(),
                        in_d.data_ptr(), pad_d.data_ptr(), fin.data_ptr(), step_d.data_ptr(), part.data_ptr(), cnt.data_ptr(),
                        B, H, Dh, rot, max_len, max_in, splits, 1.0 / math.sqrt(Dh))
    for _ in range(2):                              # twice: the split counters must reset themselves
        capi.check(lib.ftcf_mmha_decode(p, stream()))
    torch.cuda.synchronize()
    assert_close("mmha ctx", ctx.float().cpu(), ref, rtol=5e-3, atol=2e-3)
    assert_close("k append", kc_d.float().cpu()[:, :, step - 1], kc_ref[:, :, step - 1], rtol=2e-3, atol=1e-3)
    assert torch.equal(vc_d.float().cpu()[:, :, step - 1], vc_ref[:, :, step - 1])
    assert int(cnt.abs().sum()) == 0


def test_mmha_finished_rows_untouched(lib, cuda):
    torch.manual_seed(1)
    B, H, Dh, max_len = 2, 2, 64, 32
    qkv = torch.randn(B, 3 * H * Dh, device=cuda).half()
    kc = torch.randn(B, H, max_len, Dh, device=cuda).half()
    vc = torch.randn(B, H, max_len, Dh, device=cuda).half()
    ctx = torch.full((B, H * Dh), 7.0, dtype=torch.float16, device=cuda)
    tl = torch.tensor([9, 9], dtype=torch.int32, device=cuda)
    inl = torch.tensor([8, 8], dtype=torch.int32, device=cuda)
    pad = torch.zeros(B, dtype=torch.int32, device=cuda)
    fin = torch.tensor([1, 0], dtype=torch.uint8, device=cuda)
    step_d = torch.tensor([10], dtype=torch.int32, device=cuda)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    part = torch.zeros(16, dtype=torch.float32, device=cuda)
    kc0 = kc.clone()
    p = capi.MmhaParams(qkv.data_ptr(), None, kc.data_ptr(), vc.data_ptr(), ctx.data_ptr(), tl.data_ptr(), inl.data_ptr(),
                        pad.data_ptr(), fin.data_ptr(), step_d.data_ptr(), part.data_ptr(), cnt.data_ptr(), B, H, Dh, 16,
                        max_len, 8, 1, 0.125)
    capi.check(lib.ftcf_mmha_decode(p, stream()))
    torch.cuda.synchronize()
    assert torch.all(ctx[0] == 7.0) and not torch.all(ctx[1] == 7.0)
    assert torch.equal(kc[0], kc0[0])


@pytest.mark.parametrize("Dh,rot", [(64, 16), (128, 64)])
def test_prefill_scatter_and_attention(lib, cuda, Dh, rot):
    torch.manual_seed(Dh)
    B, H, max_len = 3, 2, 80
    lens = [37, 5, 64]
    T = sum(lens)
    hl = H * Dh
    qkv = torch.randn(T, 3 * hl).half().float()
    bias = (0.1 * torch.randn(3 * hl)).half().float()
    tok_b = torch.tensor(sum([[b] * n for b, n in enumerate(lens)], []), dtype=torch.int32)
    tok_p = torch.tensor(sum([list(range(n)) for n in lens], []), dtype=torch.int32)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    # oracle (gptneox_ref.GptNeoXRef.forward, bias_rotary + ctx_attn)
    x = R.h(qkv + bias)
    q, k, v = [z.reshape(T, H, Dh) for z in x.split(hl, dim=-1)]
    cos, sin = R.rotary_coef(tok_p, rot)
    q = R.apply_rotary_neox(q, cos[:, None, :], sin[:, None, :], rot)
    k = R.apply_rotary_neox(k, cos[:, None, :], sin[:, None, :], rot)
    scale = float(torch.tensor(1.0 / math.sqrt(Dh)).half())
    ctx_ref = torch.zeros(T, H, Dh)
    for b in range(B):
        s, e = offs[b], offs[b + 1]
        n = e - s
        qb, kb, vb = q[s:e].transpose(0, 1), k[s:e].transpose(0, 1), v[s:e].transpose(0, 1)
        sc = (qb @ kb.transpose(1, 2)) * scale + (torch.tril(torch.ones(n, n)) == 0) * (-10000.0)
        p = R.h(torch.softmax(sc, dim=-1))
        ctx_ref[s:e] = R.h(p @ vb).transpose(0, 1)

    d = lambda t, dt: t.to(cuda, dt).contiguous()
    qkv_d, bias_d = d(qkv, torch.float16), d(bias, torch.float16)
    q_out = torch.empty(T, H, Dh, dtype=torch.float16, device=cuda)
    kc = torch.zeros(B, H, max_len, Dh, dtype=torch.float16, device=cuda)
    vc = torch.zeros_like(kc)
    tb, tp_, of = d(tok_b, torch.int32), d(tok_p, torch.int32), torch.from_numpy(offs).to(cuda)
    capi.check(lib.ftcf_prefill_qkv_rotary_scatter(qkv_d.data_ptr(), bias_d.data_ptr(), q_out.data_ptr(), kc.data_ptr(), vc.data_ptr(),
                                                   tb.data_ptr(), tp_.data_ptr(), T, H, Dh, rot, max_len, stream()))
    ctx = torch.zeros(T, hl, dtype=torch.float16, device=cuda)
    capi.check(lib.ftcf_prefill_attention(q_out.data_ptr(), kc.data_ptr(), vc.data_ptr(), ctx.data_ptr(), of.data_ptr(), B, max(lens),
                                          H, Dh, max_len, scale, stream()))
    torch.cuda.synchronize()
    assert_close("prefill q", q_out.float().cpu(), q, rtol=2e-3, atol=1e-3)
    for b in range(B):
        s, e = offs[b], offs[b + 1]
        assert_close("prefill k cache", kc[b, :, :e - s].float().cpu(), k[s:e].transpose(0, 1), rtol=2e-3, atol=1e-3)
        assert torch.equal(vc[b, :, :e - s].float().cpu(), v[s:e].transpose(0, 1))
    assert_close("prefill ctx", ctx.float().cpu(), ctx_ref.reshape(T, hl), rtol=5e-3, atol=2e-3)


# ------------------------------------------------------------------ tcgen05 kernels (impl = 2)
@pytest.mark.parametrize("m", [1, 16, 17, 33, 64, 100, 128, 129, 300])
@pytest.mark.parametrize("n,k", [(256, 128), (1024, 4096), (5120, 640), (200, 256)])
def test_w8a16_tcgen05_grid(lib, cuda, m, n, k):
    torch.manual_seed(734876213 + m + n)
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = _quant(w)
    x = torch.randn(m, k, device=cuda).half()
    y = _gemm_w8(lib, x, p, s, None, m, n, k, 0, impl=2)
    ref = x.float() @ (q.float() * s.float()[None, :])
    assert_close(f"tcgen05 w8a16 m={m} n={n} k={k}", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=2e-3)


def test_w8a16_tcgen05_exact_dequant_round_trip(lib, cuda):
    torch.manual_seed(1)
    k, n = 256, 384
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = _quant(w)
    x = torch.eye(k, dtype=torch.float16, device=cuda)
    y = _gemm_w8(lib, x, p, s, None, k, n, k, 0, impl=2)
    assert torch.equal(y, (q.float() * s.float()[None, :]).half())


@pytest.mark.parametrize("m", [40, 256])
def test_w8a16_tcgen05_bias_gelu(lib, cuda, m):
    torch.manual_seed(11 + m)
    n, k = 2048, 1024
    w = (torch.randn(k, n, device=cuda) * 0.02).half()
    p, s, q = _quant(w)
    x = torch.randn(m, k, device=cuda).half()
    bias = (torch.randn(n, device=cuda) * 0.1).half()
    y = _gemm_w8(lib, x, p, s, bias, m, n, k, 1, impl=2)
    acc = x.float() @ (q.float() * s.float()[None, :]) + bias.float()
    ref = torch.nn.functional.gelu(acc, approximate="tanh")
    assert_close(f"tcgen05 w8a16 gelu m={m}", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("m", [8, 40, 200])
@pytest.mark.parametrize("out_f32", [0, 1])
def test_f16_tcgen05(lib, cuda, m, out_f32):
    torch.manual_seed(5 + m)
    n, k = 1008, 768
    w_nk = (torch.randn(n, k, device=cuda) * 0.02).half()
    x = torch.randn(m, k, device=cuda).half()
    y = torch.empty(m, n, dtype=torch.float32 if out_f32 else torch.float16, device=cuda)
    capi.check(lib.ftcf_gemm_f16(x.data_ptr(), w_nk.data_ptr(), None, y.data_ptr(), m, n, k, n, 0, out_f32, 2, stream()))
    torch.cuda.synchronize()
    ref = x.float() @ w_nk.float().t()
    if out_f32:
        assert_close("tcgen05 f16 gemm fp32 out", y.cpu(), ref.cpu(), rtol=1e-4, atol=1e-4)
    else:
        assert_close("tcgen05 f16 gemm fp16 out", y.float().cpu(), ref.cpu(), rtol=1e-3, atol=1e-3)


def _ln_prologue(x, ffn, attn, bias, g, b, x_out):
    pro = capi.LnPrologue()
    pro.x, pro.gamma, pro.beta, pro.eps = x.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-5
    if ffn is not None:
        pro.add_ffn, pro.add_attn = ffn.data_ptr(), attn.data_ptr()
        pro.add_bias = bias.data_ptr() if bias is not None else None
    pro.x_out = x_out.data_ptr() if x_out is not None else None
    return pro


@pytest.mark.parametrize("m", [1, 2, 4])
@pytest.mark.parametrize("with_res", [False, True])
@pytest.mark.parametrize("n,k", [(1024, 5120), (300, 256)])
def test_w8a16_fused_residual_layernorm_prologue(lib, cuda, m, with_res, n, k):
    """ftcf_gemm_w8a16_ln == residual add (exact fp16 adds) -> LayerNorm -> INT8 GEMM; the stored residual stream is exact."""
    torch.manual_seed(100 * m + n + int(with_res))
    w = (torch.randn(k, n, device=cuda) * 0.02).half()
    p, s, q = _quant(w)
    x, ffn, attn = [torch.randn(m, k, device=cuda).half() for _ in range(3)]
    bias = (0.1 * torch.randn(k, device=cuda)).half()
    g = (1 + 0.1 * torch.randn(k, device=cuda)).half()
    b = (0.1 * torch.randn(k, device=cuda)).half()
    obias = (0.1 * torch.randn(n, device=cuda)).half()
    x_out = torch.zeros_like(x)
    y = torch.empty(m, n, dtype=torch.float16, device=cuda)
    pro = _ln_prologue(x, ffn if with_res else None, attn, bias, g, b, x_out if with_res else None)
    capi.check(lib.ftcf_gemm_w8a16_ln(pro, p.data_ptr(), s.data_ptr(), obias.data_ptr(), y.data_ptr(), m, n, k, 1, stream()))
    torch.cuda.synchronize()
    r = x.float().cpu()
    if with_res:
        r = R.h(R.h(R.h(ffn.float().cpu() + attn.float().cpu()) + bias.float().cpu()) + r)
        assert torch.equal(x_out.float().cpu(), r)
    a = R.layernorm_ref(r, g.cpu(), b.cpu(), 1e-5)
    # the same input through the unfused kernels must give the same GEMM result up to LayerNorm's reduction-order ulp
    a_dev = a.half().to(cuda)
    y2 = torch.empty_like(y)
    capi.check(lib.ftcf_gemm_w8a16(a_dev.data_ptr(), p.data_ptr(), s.data_ptr(), obias.data_ptr(), y2.data_ptr(), m, n, k, 1, 1, stream()))
    torch.cuda.synchronize()
    ref = R.gelu_f32(a.float() @ (q.float().cpu() * s.float().cpu()[None, :]) + obias.float().cpu())
    assert_close("fused-prologue gemm vs unfused", y.float().cpu(), y2.float().cpu(), rtol=4e-3, atol=4e-3)
    assert_close("fused-prologue gemm vs oracle", y.float().cpu(), ref, rtol=4e-3, atol=6e-3)


@pytest.mark.parametrize("out_f32", [0, 1])
def test_f16_fused_layernorm_prologue(lib, cuda, out_f32):
    torch.manual_seed(5 + out_f32)
    m, n, k = 2, 520, 768
    w = (torch.randn(n, k, device=cuda) * 0.02).half()        # K-major [n, k]
    x, ffn, attn = [torch.randn(m, k, device=cuda).half() for _ in range(3)]
    g = (1 + 0.1 * torch.randn(k, device=cuda)).half()
    b = (0.1 * torch.randn(k, device=cuda)).half()
    y = torch.empty(m, n, dtype=torch.float32 if out_f32 else torch.float16, device=cuda)
    pro = _ln_prologue(x, ffn, attn, None, g, b, None)
    capi.check(lib.ftcf_gemm_f16_ln(pro, w.data_ptr(), None, y.data_ptr(), m, n, k, n, 0, out_f32, stream()))
    torch.cuda.synchronize()
    r = R.h(R.h(ffn.float().cpu() + attn.float().cpu()) + x.float().cpu())
    a = R.layernorm_ref(r, g.cpu(), b.cpu(), 1e-5)
    ref = a.float() @ w.float().cpu().T
    assert_close("fused-prologue f16 gemm", y.float().cpu(), ref, rtol=4e-3, atol=6e-3)


def test_fused_prologue_rejects_large_m(lib, cuda):
    x = torch.zeros(5, 256, dtype=torch.float16, device=cuda)
    g = torch.ones(256, dtype=torch.float16, device=cuda)
    w = torch.zeros(64, 256, dtype=torch.uint8, device=cuda)
    sc = torch.ones(64, dtype=torch.float16, device=cuda)
    y = torch.empty(5, 64, dtype=torch.float16, device=cuda)
    pro = _ln_prologue(x, None, None, None, g, g, None)
    assert lib.ftcf_gemm_w8a16_ln(pro, w.data_ptr(), sc.data_ptr(), None, y.data_ptr(), 5, 64, 256, 0, stream()) == 4   # FTCF_ERR_UNSUPPORTED


@pytest.mark.parametrize("m", [1, 7, 8, 20])
@pytest.mark.parametrize("n,k", [(5120, 20480), (2048, 8192), (5120, 5120), (96, 4096)])
def test_w8a16_launch_hints_do_not_change_results(lib, cuda, m, n, k):
    """The launch hint of ftcf_gemm_w8a16_ex (CTA target, no-PDL) and the k-split it implies on the tcgen05 decode kernel never
    change the result beyond the summation order (same tolerance), and every choice is deterministic."""
    torch.manual_seed(99 + m + n)
    w = (torch.randn(k, n, device=cuda) * 0.002).half()
    p, s, q = _quant(w)
    x = torch.randn(m, k, device=cuda).half()
    bias = (0.05 * torch.randn(n, device=cuda)).half()
    ref = R.gelu_f32((x.float() @ (q.float() * s.float()[None, :])).cpu() + bias.float().cpu())
    for impl in (0, 1, 3):
        for target, no_pdl, stages in ((0, 0, 0), (148, 1, 8), (444, 0, 3)):
            hint = capi.LaunchHint(target, no_pdl, stages)
            ys = []
            for _ in range(2):
                y = torch.empty(m, n, dtype=torch.float16, device=cuda)
                capi.check(lib.ftcf_gemm_w8a16_ex(x.data_ptr(), p.data_ptr(), s.data_ptr(), bias.data_ptr(), y.data_ptr(), m, n, k, 1, impl, hint,
                                                  stream()))
                torch.cuda.synchronize()
                ys.append(y)
            assert torch.equal(ys[0], ys[1])
            assert_close(f"w8a16 impl={impl} target={target} m={m} n={n} k={k}", ys[0].float().cpu(), ref, rtol=2e-3, atol=3e-3)
