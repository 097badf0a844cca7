"""GPU parity of the prefill path against the reference's own kernels in oracle/_ref/libref_kernels.so: the bias + NeoX rotary +
split (invokeAddFusedQKVBiasTranspose, kernels/unfused_attention_kernels.cu:1326-1484), invokeMaskedSoftmax
(kernels/unfused_attention_kernels.cu:255-333) inside the unfused attention chain, and invokeGatherTree as
GptNeoX<T>::setOutputTensors calls it (models/gptneox/GptNeoX.cc:1141-1164)."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi
from helpers import assert_close, stream

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip(f"{REF_SO} is missing (run __graft_entry__.build() in the authoring container)")
    return C.CDLL(REF_SO)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _tokens(lens, S):
    tok_b, tok_p, pad_off = [], [], []
    skipped = 0
    for b, n in enumerate(lens):
        for p in range(n):
            tok_b.append(b)
            tok_p.append(p)
            pad_off.append(skipped)
        skipped += S - n
    return np.asarray(tok_b, np.int32), np.asarray(tok_p, np.int32), np.asarray(pad_off, np.int32)


@pytest.mark.parametrize("Dh,rot", [(64, 16), (128, 128), (128, 32)])
@pytest.mark.parametrize("B,S,H,lens", [(2, 12, 4, [12, 7]), (3, 40, 5, [1, 40, 23])])
def test_prefill_bias_rotary_split_vs_reference_kernel(lib, ref, cuda, Dh, rot, B, S, H, lens):
    """K7.  (Round 1 parked this comparison as failing: the harness passed `.data_ptr()` of temporaries, the caching allocator
    handed tok_b's block to tok_p, and our kernel then scattered K/V rows with batch = position, out of bounds.  Every device
    tensor is held in a named variable now.)"""
    torch.manual_seed(Dh + rot + S)
    max_len = S + 8
    tok_b, tok_p, pad_off = _tokens(lens, S)
    T = len(tok_b)
    qkv = torch.randn(T, 3 * H * Dh, device=cuda).half()
    bias = (0.1 * torch.randn(3 * H * Dh, device=cuda)).half()
    tok_b_d, tok_p_d, pad_off_d = (torch.from_numpy(a).to(cuda) for a in (tok_b, tok_p, pad_off))
    # ours
    q_o = torch.zeros(T, H, Dh, dtype=torch.float16, device=cuda)
    kc = torch.zeros(B, H, max_len, Dh, dtype=torch.float16, device=cuda)
    vc = torch.zeros_like(kc)
    capi.check(lib.ftcf_prefill_qkv_rotary_scatter(qkv.data_ptr(), bias.data_ptr(), q_o.data_ptr(), kc.data_ptr(), vc.data_ptr(),
                                                   tok_b_d.data_ptr(), tok_p_d.data_ptr(), T, H, Dh, rot, max_len, stream()))
    # the reference's kernel: q / k / v [B, H, S, Dh]; it also rewrites its qkv input in place, hence the clone
    q_r = torch.zeros(B, H, S, Dh, dtype=torch.float16, device=cuda)
    k_r, v_r = torch.zeros_like(q_r), torch.zeros_like(q_r)
    qkv_in = qkv.clone()
    assert ref.ref_prefill_qkv_bias_rotary_transpose(_p(q_r), _p(k_r), _p(v_r), _p(qkv_in), _p(bias), _p(pad_off_d), B, S, T, H, Dh, rot,
                                                     C.c_void_p(stream())) == 0
    torch.cuda.synchronize()
    q_o_c, kc_c, vc_c, q_r_c, k_r_c, v_r_c = (x.float().cpu() for x in (q_o, kc, vc, q_r, k_r, v_r))
    for t in range(T):
        b, p = int(tok_b[t]), int(tok_p[t])
        # rotary in fp32 on fp16 inputs, rounded once: identical up to the sincos implementation (one fp16 ulp)
        assert_close(f"q token {t}", q_o_c[t], q_r_c[b, :, p], rtol=2e-3, atol=1e-3)
        assert_close(f"k token {t}", kc_c[b, :, p], k_r_c[b, :, p], rtol=2e-3, atol=1e-3)
        assert torch.equal(vc_c[b, :, p], v_r_c[b, :, p])
        # the non-rotary dims are bias adds only: bit-exact
        assert torch.equal(q_o_c[t][:, rot:], q_r_c[b, :, p][:, rot:])
    # nothing outside the prompts was written (pad gap and the decode part of the cache stay untouched)
    for b, n in enumerate(lens):
        assert not kc_c[b, :, n:].any() and not vc_c[b, :, n:].any()


def test_prefill_attention_vs_reference_softmax_chain(lib, ref, cuda):
    """Our fused causal attention against the reference's unfused chain with ITS masked-softmax kernel in the middle:
    qk = Q.K^T in fp32 (cuBLAS there, torch.matmul here), softmax(qk * scale + (1 - mask) * -10000) -> fp16, out = P.V."""
    torch.manual_seed(3)
    B, S, H, Dh, max_len = 2, 40, 3, 128, 48
    lens = [40, 23]
    tok_b, tok_p, _ = _tokens(lens, S)
    T = len(tok_b)
    q = (0.5 * torch.randn(T, H, Dh, device=cuda)).half()
    kc = (0.5 * torch.randn(B, H, max_len, Dh, device=cuda)).half()
    vc = torch.randn(B, H, max_len, Dh, device=cuda).half()
    seq_off = torch.tensor([0, lens[0], lens[0] + lens[1]], dtype=torch.int32, device=cuda)
    ctx = torch.zeros(T, H * Dh, dtype=torch.float16, device=cuda)
    scale = float(torch.tensor(1.0 / math.sqrt(Dh)).half())
    capi.check(lib.ftcf_prefill_attention(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), ctx.data_ptr(), seq_off.data_ptr(), B, S, H, Dh,
                                          max_len, scale, stream()))
    # reference chain on padded [B, H, S, Dh]
    qp = torch.zeros(B, H, S, Dh, dtype=torch.float16, device=cuda)
    for t in range(T):
        qp[int(tok_b[t]), :, int(tok_p[t])] = q[t]
    qk = torch.matmul(qp.float(), kc[:, :, :S].float().transpose(-1, -2)).contiguous()          # [B, H, S, S] fp32
    mask = torch.zeros(B, S, S, dtype=torch.float16, device=cuda)
    for b, n in enumerate(lens):
        mask[b, :n, :n] = torch.tril(torch.ones(n, n, dtype=torch.float16, device=cuda))        # gpt_kernels.cu:1036-1050
    probs = torch.empty(B, H, S, S, dtype=torch.float16, device=cuda)
    assert ref.ref_masked_softmax_half(_p(probs), _p(qk), _p(mask), B, H, S, S, C.c_float(scale), C.c_void_p(stream())) == 0
    torch.cuda.synchronize()
    out = torch.matmul(probs.float(), vc[:, :, :S].float()).half()                               # [B, H, S, Dh]
    for t in range(T):
        b, p = int(tok_b[t]), int(tok_p[t])
        assert_close(f"ctx token {t}", ctx[t].float().cpu(), out[b, :, p].reshape(-1).float().cpu(), rtol=1e-2, atol=3e-3)


def test_gather_output_vs_reference_gather_tree(lib, ref, cuda):
    g = np.random.default_rng(4)
    B, max_in, out_len, end_id = 3, 6, 5, 99
    max_len = max_in + out_len
    in_len = np.asarray([6, 2, 4], np.int32)
    ids = g.integers(0, 90, size=(max_len, B)).astype(np.int32)
    ids[max_in + 3, 1] = end_id                                   # row 1 produced end_id: everything after is end_id
    seq = np.asarray([max_len - 1, max_in + 3 - 1 + 1, max_len - 1], np.int32)     # counted as the engine counts them (time step t -> t - 1)
    d = lambda a: torch.from_numpy(a).to(cuda)
    ids_d, seq_d, len_d = d(ids), d(seq), d(in_len)
    out_o = torch.zeros(B, max_len, dtype=torch.int32, device=cuda)
    len_o = torch.zeros(B, dtype=torch.int32, device=cuda)
    capi.check(lib.ftcf_gather_output(out_o.data_ptr(), len_o.data_ptr(), ids_d.data_ptr(), seq_d.data_ptr(), len_d.data_ptr(), B, max_in,
                                      max_len, end_id, stream()))
    out_r = torch.zeros(B, 1, max_len, dtype=torch.int32, device=cuda)
    seq_r = seq_d.clone()
    scratch = torch.zeros(B, max_len, dtype=torch.int32, device=cuda)
    ends = torch.full((B,), end_id, dtype=torch.int32, device=cuda)
    assert ref.ref_gather_tree_sampling(_p(out_r), _p(seq_r), _p(scratch), max_len, B, _p(ids_d), _p(ends), _p(len_d), max_in,
                                        C.c_void_p(stream())) == 0
    torch.cuda.synchronize()
    assert torch.equal(out_o, out_r[:, 0]), (out_o.tolist(), out_r[:, 0].tolist())
    assert torch.equal(len_o, seq_r)
