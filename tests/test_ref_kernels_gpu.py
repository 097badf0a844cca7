"""GPU parity against the REFERENCE's OWN CUDA kernels (oracle/_ref/libref_kernels.so: the reference's .cu files compiled for
sm_100a by oracle/Makefile, never copied): LayerNorm (K4), the parallel-residual add (K5), the decode attention (K1) and the
top-k sampling chain (K12) run on the same inputs as our kernels, and as the CPU oracle -- which pins both.  Tolerances are
stated per test; integer results (token ids, finished flags, lengths) and the fp16 residual are compared bit for bit."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi
from oracle import gptneox_ref as R
from oracle import sampling_ref as S
from helpers import assert_close, stream

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        # test infrastructure, not the product: built by __graft_entry__.build() where /root/reference exists (oracle/Makefile) and
        # shipped with the snapshot; a checkout without it can still run every other parity test
        pytest.skip(f"{REF_SO} is missing (run __graft_entry__.build() in the authoring container)")
    lib = C.CDLL(REF_SO)
    lib.ref_curand_state_bytes.restype = C.c_size_t
    return lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("m,n", [(1, 5120), (4, 5120), (3, 768), (32, 5120)])
def test_layernorm_vs_reference_kernel(lib, ref, cuda, m, n):
    torch.manual_seed(m + n)
    x = (torch.randn(m, n, device=cuda) * 2 + 0.3).half()
    g = (1 + 0.1 * torch.randn(n, device=cuda)).half()
    b = (0.1 * torch.randn(n, device=cuda)).half()
    ours, theirs = torch.empty_like(x), torch.empty_like(x)
    capi.check(lib.ftcf_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), ours.data_ptr(), m, n, 1e-5, stream()))
    assert ref.ref_layernorm_half(_p(theirs), _p(x), _p(g), _p(b), C.c_float(1e-5), m, n, C.c_void_p(stream())) == 0
    torch.cuda.synchronize()
    # same arithmetic; only the fp32 reduction order of the statistics differs -> at most one fp16 ulp, on few elements
    assert_close("layernorm vs reference kernel", ours.float().cpu(), theirs.float().cpu(), rtol=2e-3, atol=2e-3)
    assert (ours == theirs).float().mean().item() > 0.97
    oracle = R.layernorm_ref(x.float().cpu(), g.cpu(), b.cpu(), 1e-5)
    assert_close("oracle vs reference kernel", oracle, theirs.float().cpu(), rtol=2e-3, atol=2e-3)
    assert (oracle == theirs.float().cpu()).float().mean().item() > 0.97


@pytest.mark.parametrize("tp", [1, 2, 8])
def test_residual_vs_reference_kernel(lib, ref, cuda, tp):
    torch.manual_seed(tp)
    m, n = 3, 5120
    x, ffn, attn = [torch.randn(m, n, device=cuda).half() for _ in range(3)]
    bias = (0.1 * torch.randn(n, device=cuda)).half()
    ours, theirs = torch.empty_like(x), torch.empty_like(x)
    capi.check(lib.ftcf_add_bias_attn_ffn_residual(ours.data_ptr(), ffn.data_ptr(), attn.data_ptr(), x.data_ptr(), bias.data_ptr(), m, n, tp,
                                                   stream()))
    assert ref.ref_add_bias_attn_ffn_residual_half(_p(theirs), _p(ffn), _p(attn), _p(x), _p(bias), m, n, tp, C.c_void_p(stream())) == 0
    torch.cuda.synchronize()
    assert torch.equal(ours, theirs)                                   # fp16 adds in the reference's order: bit-exact


@pytest.mark.parametrize("Dh,rot", [(128, 128), (128, 32), (64, 16)])
@pytest.mark.parametrize("splits", [1, 4])
def test_mmha_vs_reference_kernel(lib, ref, cuda, Dh, rot, splits):
    torch.manual_seed(Dh + rot + splits)
    B, H, max_in, max_len = 3, 5, 48, 160
    in_len = torch.tensor([48, 17, 33], dtype=torch.int32)
    step = 101                                      # 53 tokens already generated
    tl = torch.full((B,), step - 1, dtype=torch.int32)
    pad = (max_in - in_len).to(torch.int32)
    qkv = torch.randn(B, 3 * H * Dh).half()
    bias = (0.1 * torch.randn(3 * H * Dh)).half()
    kc = torch.randn(B, H, max_len, Dh).half()
    vc = torch.randn(B, H, max_len, Dh).half()
    masked = torch.zeros(B, max_len, dtype=torch.bool)
    for b in range(B):
        masked[b, int(in_len[b]):max_in] = True

    d = lambda t: t.to(cuda).contiguous()
    qkv_d, bias_d = d(qkv), d(bias)
    # ---- ours: K, V = [B, H, max_len, Dh]
    kc_o, vc_o = d(kc), d(vc)
    ctx_o = torch.zeros(B, H * Dh, dtype=torch.float16, device=cuda)
    tl_d, in_d, pad_d = d(tl), d(in_len), d(pad)
    fin = torch.zeros(B, dtype=torch.uint8, device=cuda)
    step_d = torch.tensor([step], dtype=torch.int32, device=cuda)
    part = torch.zeros(B * H * splits * (Dh + 2), dtype=torch.float32, device=cuda)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    p = capi.MmhaParams(qkv_d.data_ptr(), bias_d.data_ptr(), kc_o.data_ptr(), vc_o.data_ptr(), ctx_o.data_ptr(), tl_d.data_ptr(),
                        in_d.data_ptr(), pad_d.data_ptr(), fin.data_ptr(), step_d.data_ptr(), part.data_ptr(), cnt.data_ptr(),
                        B, H, Dh, rot, max_len, max_in, splits, 1.0 / math.sqrt(Dh))
    capi.check(lib.ftcf_mmha_decode(p, stream()))
    # ---- the reference's kernel: K = [B, H, Dh/8, max_len, 8], V = [B, H, max_len, Dh]
    kc_r = d(kc.view(B, H, max_len, Dh // 8, 8).permute(0, 1, 3, 2, 4))
    vc_r = d(vc)
    ctx_r = torch.zeros(B, H * Dh, dtype=torch.float16, device=cuda)
    fin_r = torch.zeros(B, dtype=torch.bool, device=cuda)
    masked_d = d(masked)
    rc = ref.ref_mmha_half(_p(qkv_d), _p(bias_d), _p(kc_r), _p(vc_r), _p(ctx_r), _p(fin_r), _p(tl_d), B, H, Dh, rot, max_len, max_in,
                           _p(pad_d), step, _p(masked_d), C.c_void_p(stream()))
    assert rc == 0
    torch.cuda.synchronize()
    # fp16 outputs of O(0.3): the reference rounds the probabilities to fp16 before P.V and reduces in a different order
    assert_close("decode attention vs reference kernel", ctx_o.float().cpu(), ctx_r.float().cpu(), rtol=5e-3, atol=2e-3)
    k_new_r = kc_r.permute(0, 1, 3, 2, 4).reshape(B, H, max_len, Dh)[:, :, step - 1]
    assert_close("appended K row (bias + NeoX rotary)", kc_o[:, :, step - 1].float().cpu(), k_new_r.float().cpu(), rtol=2e-3, atol=1e-3)
    assert torch.equal(vc_o[:, :, step - 1], vc_r[:, :, step - 1])
    # every other cache row is untouched on both sides
    keep = torch.ones(max_len, dtype=torch.bool)
    keep[step - 1] = False
    assert torch.equal(kc_o[:, :, keep].cpu(), kc[:, :, keep]) and torch.equal(vc_r[:, :, keep].cpu(), vc[:, :, keep])


@pytest.mark.parametrize("k,p,want_probs", [(1, 0.0, False), (4, 1.0, True), (40, 0.9, True), (16, 0.5, False), (200, 0.95, True)])
def test_topk_sampling_chain_vs_reference_kernels(lib, ref, cuda, k, p, want_probs):
    """temperature -> repetition penalty -> end mask (-> softmax) -> top-k sampling over several steps with the same seeds:
    token ids, finished flags and sequence lengths must be identical, cumulative log-probs equal to fp32 rounding."""
    B, V, Vp, max_in, out_len = 3, 5000, 5008, 4, 6
    max_len = max_in + out_len
    end_id = V - 1
    lens = [4, 2, 3]
    g = np.random.default_rng(k)
    prompt = g.integers(0, V - 1, size=(B, max_in))
    ids0 = np.zeros((max_len, B), dtype=np.int32)
    ids0[:max_in] = prompt.T
    ks, ps, _ = S.setup_topk_runtime_args(k, p, B)
    temps = np.asarray([0.7, 1.0, 1.3], np.float32)
    reps = np.asarray([1.2, 1.0, 1.1], np.float32)
    seeds = np.asarray([7, 7, 123456789012], np.int64)
    dev = cuda
    t = lambda a, dt: torch.from_numpy(np.asarray(a)).to(dev, dt)

    # ---- ours
    o_ids, o_seq = t(ids0, torch.int32), torch.full((B,), max_in - 1, dtype=torch.int32, device=dev)
    o_fin, o_cum = torch.zeros(B, dtype=torch.uint8, device=dev), torch.zeros(B, dtype=torch.float32, device=dev)
    d_len, d_k, d_p = t(lens, torch.int32), t(ks, torch.int32), t(ps, torch.float32)
    d_t, d_r = t(temps, torch.float32), t(reps, torch.float32)
    d_step = torch.tensor([max_in], dtype=torch.int32, device=dev)
    d_seeds = t(seeds.astype(np.uint64).view(np.int64), torch.int64)
    o_states = torch.zeros(B * lib.ftcf_curand_state_bytes(), dtype=torch.uint8, device=dev)
    capi.check(lib.ftcf_curand_init(o_states.data_ptr(), d_seeds.data_ptr(), B, stream()))
    max_top_k = int(ks.max())
    ws = torch.zeros(lib.ftcf_sampling_workspace_bytes(B, Vp, max_top_k) + B * max_len * 4 + 256, dtype=torch.uint8, device=dev)
    flag = torch.zeros(2, dtype=torch.int32, device=dev)
    o_logits = torch.empty(B, Vp, dtype=torch.float32, device=dev)
    sp = capi.SamplingParams(o_logits.data_ptr(), o_ids.data_ptr(), o_seq.data_ptr(), o_fin.data_ptr(), o_cum.data_ptr(), d_len.data_ptr(),
                             d_k.data_ptr(), d_p.data_ptr(), d_t.data_ptr(), d_r.data_ptr(), None, None, o_states.data_ptr(),
                             d_step.data_ptr(), flag.data_ptr(), ws.data_ptr(), B, V, Vp, max_top_k, 0, 0, max_in, max_len, end_id,
                             1 if want_probs else 0, 0)
    # ---- the reference's kernels
    r_ids, r_seq = t(ids0, torch.int32), torch.full((B,), max_in - 1, dtype=torch.int32, device=dev)
    r_fin, r_cum = torch.zeros(B, dtype=torch.bool, device=dev), torch.zeros(B, dtype=torch.float32, device=dev)
    r_states = torch.zeros(B * ref.ref_curand_state_bytes(), dtype=torch.uint8, device=dev)
    assert ref.ref_curand_batch_init(_p(r_states), B, _p(d_seeds), C.c_void_p(stream())) == 0
    end_ids = torch.full((B,), end_id, dtype=torch.int32, device=dev)
    wsz = C.c_size_t(0)
    r_logits = torch.empty(B, Vp, dtype=torch.float32, device=dev)
    assert ref.ref_batch_topk_sampling(None, C.byref(wsz), _p(r_logits), _p(r_ids), _p(r_seq), _p(r_fin), _p(r_cum), _p(r_states),
                                       max_top_k, _p(d_k), _p(d_p), Vp, _p(end_ids), B, C.c_void_p(stream())) == 0
    r_ws = torch.zeros(wsz.value + 256, dtype=torch.uint8, device=dev)

    for step in range(max_in, max_len):
        logits = np.random.default_rng(k * 7919 + step).normal(0, 3.0, size=(B, Vp)).astype(np.float32)
        if step == max_in + 2:
            logits[1, end_id] = 60.0                       # row 1 finishes here on both sides
        o_logits.copy_(torch.from_numpy(logits))
        r_logits.copy_(torch.from_numpy(logits))
        capi.check(lib.ftcf_sampling_step(sp, stream()))
        st = C.c_void_p(stream())
        assert ref.ref_temperature_penalty(_p(r_logits), _p(d_t), B, V, Vp, st) == 0
        if step > 1:
            assert ref.ref_repetition_penalty(_p(r_logits), _p(d_r), _p(r_ids), B, Vp, _p(d_len), max_in, step, st) == 0
        assert ref.ref_add_bias_end_mask(_p(r_logits), _p(end_ids), _p(r_fin), B, V, Vp, st) == 0
        if want_probs:
            assert ref.ref_add_bias_softmax(_p(r_logits), _p(end_ids), _p(r_fin), B, Vp, V, st) == 0
        wsz2 = C.c_size_t(wsz.value)
        assert ref.ref_batch_topk_sampling(_p(r_ws), C.byref(wsz2), _p(r_logits), C.c_void_p(r_ids.data_ptr() + step * B * 4), _p(r_seq),
                                           _p(r_fin), _p(r_cum) if want_probs else None, _p(r_states), max_top_k, _p(d_k), _p(d_p), Vp,
                                           _p(end_ids), B, st) == 0
        torch.cuda.synchronize()
        assert torch.equal(o_ids[step], r_ids[step]), f"step {step}: ours {o_ids[step].tolist()} reference {r_ids[step].tolist()}"
        assert torch.equal(o_fin.bool(), r_fin), f"step {step}: finished flags differ"
        assert torch.equal(o_seq, r_seq)
    if want_probs:
        np.testing.assert_allclose(o_cum.cpu().numpy(), r_cum.cpu().numpy(), rtol=1e-4, atol=1e-4)
