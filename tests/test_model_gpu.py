"""End-to-end parity of the engine (GptNeoXOp mirror over the C ABI) against the oracle model on small configurations.

Token ids must match exactly (greedy and seeded top-k); raw logits of every step are compared within an fp16-level
tolerance that is stated here: |dlogit| <= 3e-2 + 2e-2 |logit| (logits of these models are O(0.3); the chain runs
2-3 layers of fp16-rounded activations with fp32 accumulation whose summation order differs from the oracle's).
"""
import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import weights as W
from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp
from helpers import assert_close, oracle_from_rank_weights, tiny_cfg, to_cuda_lists

pytestmark = pytest.mark.gpu

LOGIT_RTOL, LOGIT_ATOL = 2e-2, 3e-2


def _build(cfg, int8_mode, cuda, seed=0, fused=1, tweak=None):
    """fused = 1: the residual + LayerNorm prologue fused into the decode GEMMs (batch <= 4, the default); 0: one kernel per
    operator, nothing fused; 2: LayerNorm-only prologue (the tensor-parallel decode path)."""
    rw = W.make_synthetic(cfg, 1, 0, int8_mode, "cpu", seed=seed, keep_plain=True)
    if tweak is not None:
        tweak(rw)
    ref = oracle_from_rank_weights(cfg, [rw], int8_mode)
    w, q, s = to_cuda_lists(rw, cuda)
    op = GptNeoXOp(None, 0, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size,
                   cfg.rotary_embedding_dim, cfg.start_id, cfg.end_id, 1, 1, int8_mode, 1024, cfg.use_gptj_residual, w, q, s)
    op.set_option("fused_ln", fused)
    return op, ref


def _prompts(B, S, V, lens, seed=1234):
    g = np.random.default_rng(seed)
    ids = g.integers(0, V - 1, size=(B, S)).astype(np.int32)
    for b, n in enumerate(lens):
        ids[b, n:] = V - 1          # right-padded with end_id (codefuse_example.py:700)
    return ids


def _compare(op, ref, cuda, ids, lens, out_len, graph, **kw):
    B, S = ids.shape
    V = ref.cfg.vocab_size
    op.set_option("cuda_graph", 1 if graph else 0)
    trace = torch.zeros(out_len, B, V, dtype=torch.float32, device=cuda) if not graph else None
    t = lambda a, dt: None if a is None else torch.tensor(np.asarray(a), dtype=dt)
    res = op.forward(torch.from_numpy(ids).to(cuda), torch.tensor(lens, dtype=torch.int32, device=cuda), out_len, 1,
                     t(kw.get("top_k"), torch.int32), t(kw.get("top_p"), torch.float32), None, t(kw.get("temperature"), torch.float32),
                     None, t(kw.get("repetition_penalty"), torch.float32), t(kw.get("random_seed"), torch.int64), None, None,
                     kw.get("return_cum_log_probs", 0), None, logits_trace=trace)
    exp = ref.forward(ids, lens, out_len, top_k=kw.get("top_k"), top_p=kw.get("top_p"), temperature=kw.get("temperature"),
                      repetition_penalty=kw.get("repetition_penalty"), random_seed=kw.get("random_seed"),
                      return_cum_log_probs=kw.get("return_cum_log_probs", 0), keep_logits=True)
    if trace is not None:
        n_gen = exp["sequence_lengths"].reshape(-1) - S
        for i, lg in enumerate(exp["logits"]):
            live = n_gen > i          # a finished row skips attention (template.hpp:1176-1178): its logits are don't-care
            assert_close(f"logits step {i}", trace[i].cpu().numpy()[live], lg[live], LOGIT_RTOL, LOGIT_ATOL)
    assert np.array_equal(res[0].cpu().numpy(), exp["output_ids"]), (res[0].cpu().numpy(), exp["output_ids"])
    assert np.array_equal(res[1].cpu().numpy(), exp["sequence_lengths"])
    if kw.get("return_cum_log_probs", 0):
        np.testing.assert_allclose(res[2].cpu().numpy(), exp["cum_log_probs"], rtol=2e-2, atol=2e-2)
    return res, exp


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("int8_mode", [0, 1])
@pytest.mark.parametrize("graph", [False, True])
def test_greedy_full_batch(cuda, int8_mode, graph, fused):
    cfg = tiny_cfg()
    op, ref = _build(cfg, int8_mode, cuda, fused=fused)
    ids = _prompts(2, 12, cfg.vocab_size, [12, 12])
    _compare(op, ref, cuda, ids, [12, 12], 10, graph)


@pytest.mark.parametrize("fused", [0, 1, 2])
@pytest.mark.parametrize("int8_mode", [0, 1])
def test_greedy_ragged_batch(cuda, int8_mode, fused):
    cfg = tiny_cfg()
    op, ref = _build(cfg, int8_mode, cuda, seed=3, fused=fused)
    lens = [16, 5, 11, 1]
    ids = _prompts(4, 16, cfg.vocab_size, lens, seed=7)
    _compare(op, ref, cuda, ids, lens, 8, False)
    _compare(op, ref, cuda, ids, lens, 8, True)


def test_sequential_residual(cuda):
    cfg = tiny_cfg(use_gptj_residual=False)
    op, ref = _build(cfg, 1, cuda, seed=5)
    lens = [9, 4]
    ids = _prompts(2, 9, cfg.vocab_size, lens, seed=9)
    _compare(op, ref, cuda, ids, lens, 6, False)


@pytest.mark.parametrize("fused", [0, 1])
def test_dh128_full_rotary(cuda, fused):
    cfg = tiny_cfg(head_num=2, size_per_head=128, rotary_embedding_dim=128, inter_size=1024, layer_num=3)
    op, ref = _build(cfg, 1, cuda, seed=11, fused=fused)
    lens = [20, 13]
    ids = _prompts(2, 20, cfg.vocab_size, lens, seed=2)
    _compare(op, ref, cuda, ids, lens, 12, False)
    _compare(op, ref, cuda, ids, lens, 12, True)


@pytest.mark.parametrize("fused", [0, 1])
def test_seeded_topk_sampling_and_cum_log_probs(cuda, fused):
    cfg = tiny_cfg()
    op, ref = _build(cfg, 1, cuda, seed=1, fused=fused)
    lens = [10, 7, 10]
    ids = _prompts(3, 10, cfg.vocab_size, lens, seed=4)
    _compare(op, ref, cuda, ids, lens, 8, False, top_k=[8, 8, 8], top_p=[0.9, 0.9, 0.9], temperature=[0.7, 0.7, 0.7],
             repetition_penalty=[1.1, 1.1, 1.1], random_seed=[42, 42, 43], return_cum_log_probs=1)


@pytest.mark.parametrize("fused", [0, 1])
def test_single_token_prompt_runs_decoder_only(cuda, fused):
    cfg = tiny_cfg()
    op, ref = _build(cfg, 1, cuda, seed=2, fused=fused)
    ids = _prompts(2, 1, cfg.vocab_size, [1, 1], seed=3)
    _compare(op, ref, cuda, ids, [1, 1], 6, False)


@pytest.mark.parametrize("int8_mode", [0, 1])
def test_long_context_batch8(cuda, int8_mode):
    """Context > 160 keys (several split-KV units per head), KV tiles that straddle the pad gap of ragged prompts, 8 sequences,
    k extents of 3 and 12 k-steps and a head count that does not divide the CTA count."""
    cfg = tiny_cfg(head_num=6, size_per_head=64, inter_size=1536, layer_num=2, vocab_size=640, rotary_embedding_dim=16, end_id=639)
    op, ref = _build(cfg, int8_mode, cuda, seed=21)
    lens = [230, 1, 97, 230, 161, 64, 200, 33]
    ids = _prompts(8, 230, cfg.vocab_size, lens, seed=5)
    _compare(op, ref, cuda, ids, lens, 5, False)
    _compare(op, ref, cuda, ids, lens, 5, True)


@pytest.mark.parametrize("fused", [0, 1])
def test_finished_rows_stop_advancing(cuda, fused):
    """Rows that sample end_id stop (their attention work disappears from the schedule, the sampler pins them to end_id)
    while the others go on; the request ends early once every row is finished.  The end_id row of the LM head is scaled up
    so that end_id wins within a few steps for some rows."""
    cfg = tiny_cfg()

    def tweak(rw):
        rw.w[12 * cfg.layer_num + 3][cfg.end_id] *= 6.0

    op, ref = _build(cfg, 1, cuda, seed=6, fused=fused, tweak=tweak)
    lens = [6, 6, 3, 5]
    ids = _prompts(4, 6, cfg.vocab_size, lens, seed=8)
    res, exp = _compare(op, ref, cuda, ids, lens, 16, False, top_k=[6, 6, 6, 6], top_p=[1.0] * 4, random_seed=[1, 2, 3, 4])
    n_gen = exp["sequence_lengths"].reshape(-1) - 6
    assert n_gen.min() < 16, "the tweak did not make any row finish early: the test would not cover finished rows"
    _compare(op, ref, cuda, ids, lens, 16, True, top_k=[6, 6, 6, 6], top_p=[1.0] * 4, random_seed=[1, 2, 3, 4])


def test_streaming_callback_and_early_stop(cuda):
    cfg = tiny_cfg()
    op, ref = _build(cfg, 1, cuda, seed=6)
    lens = [6, 6]
    ids = _prompts(2, 6, cfg.vocab_size, lens, seed=8)
    msgs = []
    res = op.forward(torch.from_numpy(ids).to(cuda), torch.tensor(lens, dtype=torch.int32, device=cuda), 7, 1, None, None, None, None, None,
                     None, None, None, None, 0, msgs.append)
    exp = ref.forward(ids, lens, 7)
    assert np.array_equal(res[0].cpu().numpy(), exp["output_ids"])
    assert len(msgs) == 6                                     # every step but the last (GptNeoX.cc:1023-1029)
    out = res[0].cpu().numpy()
    for i, m in enumerate(msgs):
        assert m["idxs"] == [[i], [i]]
        assert m["last_tokens"] == [[int(out[0, 0, 6 + i])], [int(out[1, 0, 6 + i])]]


def test_argument_errors_raise(cuda):
    cfg = tiny_cfg()
    op, _ = _build(cfg, 0, cuda)
    ids = torch.zeros(1, 4, dtype=torch.int64, device=cuda)
    with pytest.raises(RuntimeError):
        op.forward(ids, torch.tensor([4], dtype=torch.int32, device=cuda), 4)
    ids32 = torch.zeros(1, 4, dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError):
        op.forward(ids32, torch.tensor([9], dtype=torch.int32, device=cuda), 4)      # length > max_input_length
    with pytest.raises(RuntimeError):
        op.forward(ids32, torch.tensor([4], dtype=torch.int32, device=cuda), 4, 33)  # beam_width beyond the supported 32
    with pytest.raises(RuntimeError):                                                 # beam search needs a prompt to tile
        op.forward(ids32[:, :1].contiguous(), torch.tensor([1], dtype=torch.int32, device=cuda), 4, 3)


def test_pybind_shim_runs_the_reference_call(cuda):
    """libth_gptneox.GptNeoXOp (the compiled shim a user of codefuse_example.py loads) against the oracle: positional
    arguments exactly as codefuse_example.py:533-536 (constructor) and :575-589 (forward)."""
    import sys
    from fastertransformer4codefuse_b200 import capi
    if capi.LIB_DIR not in sys.path:
        sys.path.append(capi.LIB_DIR)
    import libth_gptneox
    cfg = tiny_cfg()
    rw = W.make_synthetic(cfg, 1, 0, 1, "cpu", seed=5, keep_plain=True)
    ref = oracle_from_rank_weights(cfg, [rw], 1)
    w, q, s = to_cuda_lists(rw, cuda)
    op = libth_gptneox.GptNeoXOp(None, 0, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size,
                                 cfg.rotary_embedding_dim, cfg.start_id, cfg.end_id, 1, 1, 1, 1024, cfg.use_gptj_residual, w, q, s)
    lens = [9, 14]
    ids = _prompts(2, 14, cfg.vocab_size, lens)
    seen = []
    kw = dict(top_k=[4], top_p=[0.9], temperature=[0.7], repetition_penalty=[1.1], random_seed=[7])
    res = op.forward(torch.from_numpy(ids).to(cuda), torch.tensor(lens, dtype=torch.int32, device=cuda), 8, 1,
                     torch.tensor(kw["top_k"], dtype=torch.int32), torch.tensor(kw["top_p"]), None, torch.tensor(kw["temperature"]), None,
                     torch.tensor(kw["repetition_penalty"]), torch.tensor(kw["random_seed"], dtype=torch.int64), None, None, 1, seen.append)
    exp = ref.forward(ids, lens, 8, return_cum_log_probs=1, **kw)
    assert len(res) == 3 and res[0].shape == (2, 1, 22) and res[0].dtype == torch.int32
    assert np.array_equal(res[0].cpu().numpy(), exp["output_ids"])
    assert np.array_equal(res[1].cpu().numpy(), exp["sequence_lengths"])
    np.testing.assert_allclose(res[2].cpu().numpy(), exp["cum_log_probs"], rtol=2e-2, atol=2e-2)
    assert len(seen) >= 1 and set(seen[0]) == {"last_tokens", "idxs"} and len(seen[0]["last_tokens"]) == 2
    with pytest.raises(RuntimeError):
        op.forward(torch.from_numpy(ids), torch.tensor(lens, dtype=torch.int32, device=cuda), 8, 1, None, None, None, None, None,
                   None, None, None, None, 0, None)     # CPU input_ids


def test_pure_top_p_request(cuda):
    """top_k = 0 with top_p > 0 through the whole engine (graph replay), mixed with a top-k row."""
    cfg = tiny_cfg()
    op, ref = _build(cfg, 1, cuda, seed=8, fused=0)
    lens = [10, 6]
    ids = _prompts(2, 10, cfg.vocab_size, lens, seed=3)
    _compare(op, ref, cuda, ids, lens, 12, True, top_k=[0, 3], top_p=[0.9, 0.8], temperature=[0.9, 1.0], random_seed=[21, 22],
             return_cum_log_probs=1)


def test_full_width_properties(cuda):
    """CodeFuse-13B layer shape (h 5120, 40 x 128 heads, inter 20480, vocab 100864; 2 layers so the oracle is not needed):
    size-independent properties at the BASELINE width -- (a) the same request twice gives identical ids (deterministic
    reductions, self-resetting counters), (b) graph replay == eager launches, (c) the fused residual + LayerNorm prologue
    path and the per-operator path agree on the logits within the fp16 tolerance and on every token whose top-2 margin is
    clear, (d) a batch of 3 ragged prompts gives each row the tokens it gets alone."""
    cfg = W.NeoXConfig(head_num=40, size_per_head=128, inter_size=20480, layer_num=2, vocab_size=100864, rotary_embedding_dim=128,
                       start_id=100000, end_id=100863)
    rw = W.make_synthetic_fast(cfg, 1, 0, 1, cuda, seed=3)
    rw.w[12 * cfg.layer_num + 3][cfg.end_id].zero_()
    w, q, s = rw.lists()
    op = GptNeoXOp(None, 0, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim,
                   cfg.start_id, cfg.end_id, 1, 1, 1, 2048, True, w, q, s)
    S, out = 96, 12
    g = np.random.default_rng(2)
    ids = torch.from_numpy(g.integers(0, cfg.vocab_size - 2, size=(3, S)).astype(np.int32)).to(cuda)
    lens = torch.tensor([S, 40, 71], dtype=torch.int32, device=cuda)

    def run(batch_rows, graph, fused, trace=False):
        op.set_option("cuda_graph", graph)
        op.set_option("fused_ln", fused)
        rows = torch.tensor(batch_rows, device=cuda)
        tr = torch.zeros(out, len(batch_rows), cfg.vocab_size, dtype=torch.float32, device=cuda) if trace else None
        res = op.forward(ids[rows].contiguous(), lens[rows].contiguous(), out, logits_trace=tr)
        return res[0][:, 0].cpu().numpy(), res[1].cpu().numpy(), (tr.cpu().numpy() if trace else None)

    a_ids, a_len, _ = run([0], 1, 1)
    b_ids, b_len, _ = run([0], 1, 1)
    assert np.array_equal(a_ids, b_ids) and np.array_equal(a_len, b_len)                       # (a)
    e_ids, _, e_log = run([0], 0, 1, trace=True)
    assert np.array_equal(a_ids, e_ids)                                                        # (b)
    u_ids, _, u_log = run([0], 0, 0, trace=True)
    assert_close("fused vs per-operator logits (step 0)", e_log[0], u_log[0], LOGIT_RTOL, LOGIT_ATOL)
    for t in range(out):                                                                       # (c)
        top2 = np.sort(e_log[t, 0])[-2:]
        if top2[1] - top2[0] > 4 * LOGIT_ATOL:
            assert e_ids[0, S + t] == u_ids[0, S + t], f"step {t}"
        elif e_ids[0, S + t] != u_ids[0, S + t]:
            break
    all_ids, all_len, _ = run([0, 1, 2], 1, 1)                                                 # (d)
    for r in range(3):
        one_ids, one_len, one_log = run([r], 0, 1, trace=True)
        n = int(lens[r])
        assert all_len[r, 0] == one_len[0, 0] == S + out      # counted from max_input_length, GptNeoX.cc:687-695
        for t in range(out):
            top2 = np.sort(one_log[t, 0])[-2:]
            if top2[1] - top2[0] > 4 * LOGIT_ATOL:
                assert all_ids[r, n + t] == one_ids[0, n + t], f"row {r} step {t}"
            elif all_ids[r, n + t] != one_ids[0, n + t]:
                break


@pytest.mark.parametrize("layout", [1, 2])
def test_int8_layouts_converted_at_load(cuda, layout):
    """int8_layout 1 (plain [k,n] int8) and 2 (the reference's sm80 *.q.bin bytes) are re-laid out at engine creation and
    give the tokens of the native B200 layout."""
    from oracle import quant_ref as Q
    cfg = tiny_cfg()
    rw = W.make_synthetic(cfg, 1, 0, 1, "cpu", seed=6, keep_plain=True)
    ref = oracle_from_rank_weights(cfg, [rw], 1)
    w, q, s = to_cuda_lists(rw, cuda)
    if layout == 1:
        q_alt = [p.to(cuda) for p in rw.plain_q]
    else:
        q_alt = [torch.from_numpy(Q.preprocess_weights_ampere(p.numpy())).to(cuda) for p in rw.plain_q]
    op = GptNeoXOp(None, 0, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim,
                   cfg.start_id, cfg.end_id, 1, 1, 1, 1024, cfg.use_gptj_residual, w, q_alt, s, int8_layout=layout)
    lens = [8, 5]
    ids = _prompts(2, 8, cfg.vocab_size, lens, seed=2)
    _compare(op, ref, cuda, ids, lens, 6, True)
