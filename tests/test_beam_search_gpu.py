"""Beam search (beam_width > 1) on the B200:
  * ftcf_beam_search_step against the oracle restatement (oracle/beam_search_ref.py) over seeded multi-step runs: token ids,
    parents, lengths, finished flags and the cache indirection bit-exact, cum_log_probs within 2e-4 (fp32 log-sum-exp in another
    summation order);
  * the same runs against the REFERENCE's OWN kernels (invokeAddBiasApplyPenalties + invokeTopkSoftMax compiled into
    oracle/_ref/libref_kernels.so), with the update rule of OnlineBeamSearchLayer.cu:24-60 applied to their winners;
  * gatherTree with parents against the reference's kernel;
  * the engine end to end (GptNeoXOp.forward(beam_width=K)) against the oracle's forward_beam: output ids [B, K, L], sequence
    lengths and cum_log_probs, int8 and fp16, ragged batch, graph on / off, penalties, stop words."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi, weights as W
from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp
from oracle import beam_search_ref as BS
from helpers import oracle_from_rank_weights, stream, tiny_cfg, to_cuda_lists
from helpers_beam import CASES, OracleRun

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip(f"{REF_SO} is missing (run __graft_entry__.build() in the authoring container)")
    lib = C.CDLL(REF_SO)
    lib.ref_beam_topk_workspace_floats.restype = C.c_size_t
    return lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _State(OracleRun):
    """Device state of one beam-search run (ours) beside the oracle's numpy state."""

    def __init__(self, lib, dev, B, K, V, Vp, max_in, out_len, lens, end_id, seed, **args):
        super().__init__(B, K, V, Vp, max_in, out_len, lens, end_id, seed, **args)
        self.lib, self.dev = lib, dev
        BB = B * K
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
        self.d_ids, self.d_par = t(self.ids0, torch.int32), torch.zeros(self.max_len, BB, dtype=torch.int32, device=dev)
        self.d_seq = torch.full((BB,), max_in - 1, dtype=torch.int32, device=dev)
        self.d_fin, self.d_cum = torch.zeros(BB, dtype=torch.uint8, device=dev), t(self.cum0, torch.float32)
        self.d_len = t(self.lens, torch.int32)
        self.d_ind = torch.zeros(2, BB, self.max_len, dtype=torch.int32, device=dev)
        self.d_step = torch.tensor([max_in], dtype=torch.int32, device=dev)
        self.d_logits = torch.empty(BB, Vp, dtype=torch.float32, device=dev)
        self.ws = torch.zeros(lib.ftcf_beam_workspace_bytes(B, K, Vp, self.max_len), dtype=torch.uint8, device=dev)
        self.flag = torch.zeros(2, dtype=torch.int32, device=dev)
        self.d_stop = t(self.stop, torch.int32) if self.stop is not None else None
        self.bp = capi.BeamParams(self.d_logits.data_ptr(), self.d_ids.data_ptr(), self.d_par.data_ptr(), self.d_seq.data_ptr(),
                                  self.d_fin.data_ptr(), self.d_cum.data_ptr(), self.d_len.data_ptr(), self.d_ind.data_ptr(),
                                  self.d_stop.data_ptr() if self.d_stop is not None else None, self.d_step.data_ptr(), self.flag.data_ptr(),
                                  None, self.ws.data_ptr(), B, K, V, Vp, self.stop.shape[2] if self.stop is not None else 0, max_in,
                                  self.max_len, end_id, args.get("temperature", 1.0), args.get("repetition_penalty", 1.0),
                                  args.get("diversity_rate", 0.0), args.get("length_penalty", 0.0), 0)

    def ours(self, x):
        self.d_logits.copy_(torch.from_numpy(x))
        capi.check(self.lib.ftcf_beam_search_step(self.bp, stream()))
        torch.cuda.synchronize()

    def compare(self, step, what):
        par = (step - self.max_in) % 2
        assert np.array_equal(self.d_ids.cpu().numpy()[step], self.ids[step]), f"{what} step {step}: ids"
        assert np.array_equal(self.d_par.cpu().numpy()[step], self.par[step]), f"{what} step {step}: parents"
        assert np.array_equal(self.d_seq.cpu().numpy(), self.seq), f"{what} step {step}: lengths"
        assert np.array_equal(self.d_fin.cpu().numpy().astype(bool), self.fin), f"{what} step {step}: finished"
        np.testing.assert_allclose(self.d_cum.cpu().numpy(), self.cum, rtol=0, atol=2e-4, err_msg=f"{what} step {step}: cum_log_probs")
        assert np.array_equal(self.d_ind.cpu().numpy()[1 - par], self.ind[1 - par]), f"{what} step {step}: cache indirection"
        assert int(self.d_step.item()) == step + 1


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"B{c['B']}K{c['K']}V{c['V']}")
def test_beam_step_vs_oracle(lib, cuda, case):
    case = dict(case)
    B, K, V, Vp, lens = (case.pop(k) for k in ("B", "K", "V", "Vp", "lens"))
    max_in, out_len = 6, 9
    st = _State(lib, cuda, B, K, V, Vp, max_in, out_len, lens, V - 1, seed=3, **case)
    mixed = False
    for step in range(max_in, max_in + out_len):
        x = st.logits(step, 17)
        st.ours(x)
        st.advance(x, step)
        st.compare(step, "oracle")
        mixed |= bool(st.fin.any() and not st.fin.all())
    assert mixed or "diversity_rate" in case                       # the run saw finished and live beams side by side


def test_beam_stop_words_vs_oracle(lib, cuda):
    """A stop word made of the two tokens the best beam of batch 0 produces at steps 2 and 3 (found with a free run of the
    oracle), matched through the parents."""
    B, K, V, Vp, max_in, out_len = 2, 3, 800, 800, 5, 8
    free = OracleRun(B, K, V, Vp, max_in, out_len, [5, 4], V - 1, seed=4)
    for step in range(max_in, max_in + 4):
        free.advance(free.logits(step, 23), step)
    step = max_in + 3
    tok_b = int(free.ids[step, 0])
    tok_a = int(free.ids[step - 1, int(free.par[step, 0])])
    stop = np.full((B, 2, 3), -1, np.int32)
    stop[0, 0, :2], stop[0, 1, 0] = [tok_a, tok_b], 2
    stop[1, 0, 0], stop[1, 1, 0] = 12345 % V, 1
    st = _State(lib, cuda, B, K, V, Vp, max_in, out_len, [5, 4], V - 1, seed=4, stop_words=stop)
    hit = False
    for step in range(max_in, max_in + out_len):
        x = st.logits(step, 23)
        st.ours(x)
        st.advance(x, step)
        st.compare(step, "stop words")
        hit |= bool((st.fin & (st.ids[step] != V - 1)).any())
    assert hit, "the stop word never fired"


@pytest.mark.parametrize("case", CASES[:5], ids=lambda c: f"B{c['B']}K{c['K']}V{c['V']}")
def test_beam_step_vs_reference_kernels(lib, ref, cuda, case):
    """The reference's own penalty + softmax/top-k kernels decide the winners; lengths / finished / parents follow the update rule of
    OnlineBeamSearchLayer.cu:24-60 and the indirection BaseBeamSearchLayer.cu:24-52 (both are layer files, not compiled here)."""
    case = dict(case)
    B, K, V, Vp, lens = (case.pop(k) for k in ("B", "K", "V", "Vp", "lens"))
    max_in, out_len = 6, 9
    BB = B * K
    dev = cuda
    st = _State(lib, dev, B, K, V, Vp, max_in, out_len, lens, V - 1, seed=5, **case)
    r_ids, r_par = st.d_ids.clone(), st.d_par.clone()
    r_seq, r_fin, r_cum = st.d_seq.clone(), torch.zeros(BB, dtype=torch.bool, device=dev), st.d_cum.clone()
    r_logits = torch.empty(BB, Vp, dtype=torch.float32, device=dev)
    r_win = torch.zeros(BB, dtype=torch.int32, device=dev)
    end_ids = torch.full((B,), V - 1, dtype=torch.int32, device=dev)
    nws = ref.ref_beam_topk_workspace_floats(B)
    r_ws = torch.zeros(nws, dtype=torch.float32, device=dev)
    a = case
    for step in range(max_in, max_in + out_len):
        x = st.logits(step, 29)
        st.ours(x)
        r_logits.copy_(torch.from_numpy(x))
        s = C.c_void_p(stream())
        assert ref.ref_beam_penalties(_p(r_logits), step, _p(r_ids), _p(r_par), _p(st.d_len), _p(r_seq), max_in, B, K, V, Vp, _p(end_ids),
                                      C.c_float(a.get("temperature", 1.0)), C.c_float(a.get("repetition_penalty", 1.0)), s) == 0
        assert ref.ref_beam_topk_softmax(_p(r_logits), _p(r_fin), _p(r_seq), _p(r_cum), _p(r_win), _p(r_ws), C.c_size_t(nws), B, K, Vp,
                                         _p(end_ids), C.c_float(a.get("diversity_rate", 0.0)), C.c_float(a.get("length_penalty", 0.0)), s) == 0
        torch.cuda.synchronize()
        word = r_win.cpu().numpy().astype(np.int64)
        parent, tok = (word // Vp) % K, word % Vp
        seq, fin = r_seq.cpu().numpy(), r_fin.cpu().numpy()
        base = (np.arange(BB) // K) * K
        new_seq = seq[base + parent] + np.where(fin[base + parent], 0, 1)
        r_seq.copy_(torch.from_numpy(new_seq.astype(np.int32)))
        r_fin.copy_(torch.from_numpy(tok == V - 1))
        r_ids[step].copy_(torch.from_numpy(tok.astype(np.int32)))
        r_par[step].copy_(torch.from_numpy(parent.astype(np.int32)))
        assert np.array_equal(st.d_ids.cpu().numpy()[step], tok), f"step {step}: ids differ from the reference kernels"
        assert np.array_equal(st.d_par.cpu().numpy()[step], parent), f"step {step}: parents"
        assert np.array_equal(st.d_seq.cpu().numpy(), new_seq), f"step {step}: lengths"
        np.testing.assert_allclose(st.d_cum.cpu().numpy(), r_cum.cpu().numpy(), rtol=0, atol=2e-4, err_msg=f"step {step}: cum_log_probs")
    assert bool(r_fin.any()) or "diversity_rate" in case


def test_gather_tree_beams_vs_reference_and_oracle(lib, ref, cuda):
    B, K, V, max_in, out_len = 3, 4, 500, 6, 10
    max_len = max_in + out_len
    lens = [6, 3, 5]
    st = _State(lib, cuda, B, K, V, V, max_in, out_len, lens, V - 1, seed=6)
    for step in range(max_in, max_len):
        x = st.logits(step, 31)
        st.ours(x)
        st.advance(x, step)
    dev, BB = cuda, B * K
    out = torch.zeros(B, K, max_len, dtype=torch.int32, device=dev)
    out_len_t = torch.zeros(B, K, dtype=torch.int32, device=dev)
    capi.check(lib.ftcf_gather_output_beams(out.data_ptr(), out_len_t.data_ptr(), st.d_ids.data_ptr(), st.d_par.data_ptr(), st.d_seq.data_ptr(),
                                            st.d_len.data_ptr(), B, K, max_in, max_len, V - 1, stream()))
    exp, exp_len = BS.gather_tree(st.ids, st.par, st.seq, st.lens, max_in, max_len, V - 1, K)
    assert np.array_equal(out.cpu().numpy(), exp)
    assert np.array_equal(out_len_t.cpu().numpy(), exp_len)
    r_out = torch.zeros(B, K, max_len, dtype=torch.int32, device=dev)
    r_seq = st.d_seq.clone()
    scratch = torch.zeros(max_len, BB, dtype=torch.int32, device=dev)
    end_ids = torch.full((B,), V - 1, dtype=torch.int32, device=dev)
    assert ref.ref_gather_tree_beams(_p(r_out), _p(r_seq), _p(scratch), max_len, B, K, _p(st.d_ids), _p(st.d_par), _p(end_ids), _p(st.d_len),
                                     max_in, C.c_void_p(stream())) == 0
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), r_out.cpu().numpy()), "gatherTree with parents differs from the reference kernel"
    assert np.array_equal(out_len_t.cpu().numpy().reshape(-1), r_seq.cpu().numpy())


# ------------------------------------------------------------------------------------------------- engine end to end
def _op(cfg, rw, dev, int8_mode):
    w, q, s = to_cuda_lists(rw, dev)
    return GptNeoXOp(None, 0, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim,
                     cfg.start_id, cfg.end_id, 1, 1, int8_mode, 1024, cfg.use_gptj_residual, w, q, s)


@pytest.mark.parametrize("int8_mode,gptj,K,kw,seed", [
    (1, True, 3, {}, 19),
    (0, True, 4, {}, 27),
    (1, False, 2, {}, 12),
    (1, True, 3, dict(temperature=0.8, repetition_penalty=1.2, beam_search_diversity_rate=0.3, len_penalty=0.6), 33),
])
def test_engine_beam_search_vs_oracle(cuda, int8_mode, gptj, K, kw, seed):
    """A random small model gives almost flat distributions: the score gap between neighbouring candidates is ~1e-3, the size of
    fp16 noise between two correct implementations, so two runs may legitimately pick different beams.  The comparison is therefore
    made in two halves that do not hinge on near-ties:
      (1) the engine's raw logits of every step are traced; the ORACLE takes its beam decisions on those logits -- ids, lengths,
          parents must then be identical (tiling, penalties, search, stop, gatherTree), and
      (2) the oracle's OWN logits along that same trajectory must agree with the engine's within fp16 tolerance for every live row
          (prefill of beam 0 only, cache indirection in the decode attention, new K/V rows in the beam's own slot).
    Where the oracle's free run is decisive (min score gap >= 5e-3) the free-run ids are compared too."""
    dev = cuda
    cfg = tiny_cfg(use_gptj_residual=gptj)
    rw = W.make_synthetic(cfg, 1, 0, int8_mode, "cpu", seed=21, keep_plain=True)
    ref_model = oracle_from_rank_weights(cfg, [rw], int8_mode)
    op = _op(cfg, rw, dev, int8_mode)
    g = np.random.default_rng(seed)
    lens = [10, 6]
    ids = g.integers(0, cfg.vocab_size - 1, size=(2, 10)).astype(np.int32)
    for b, n in enumerate(lens):
        ids[b, n:] = cfg.end_id
    out_len = 9
    targs = {k: torch.tensor([v], dtype=torch.float32) for k, v in kw.items()}
    d_ids, d_lens = torch.from_numpy(ids).to(dev), torch.tensor(lens, dtype=torch.int32, device=dev)
    trace = torch.zeros(out_len, 2 * K, cfg.vocab_size, dtype=torch.float32, device=dev)
    op.set_option("cuda_graph", 0)
    got = op.forward(d_ids, d_lens, out_len, beam_width=K, return_cum_log_probs=1, logits_trace=trace, **targs)
    assert tuple(got[0].shape) == (2, K, 10 + out_len) and tuple(got[1].shape) == (2, K) and tuple(got[2].shape) == (2, K)
    steps = op.last_stats["steps"]
    exp = ref_model.forward_beam(ids, lens, out_len, K, decide_on_logits=trace.cpu().numpy(), **kw)
    assert exp["steps"] <= steps                      # the engine may run kExitLag more steps after every row finished
    assert np.array_equal(got[0].cpu().numpy(), exp["output_ids"]), "beam ids differ from the oracle deciding on the same logits"
    assert np.array_equal(got[1].cpu().numpy(), exp["sequence_lengths"])
    np.testing.assert_allclose(got[2].cpu().numpy(), exp["cum_log_probs"], rtol=0, atol=1e-3)
    tr = trace.cpu().numpy()
    worst = 0.0
    for i in range(exp["steps"]):
        live = ~exp["finished_before"][i]
        worst = max(worst, float(np.abs(tr[i][live] - exp["logits"][i][live]).max()))
    assert worst < 4e-2, f"logits along the common trajectory differ by {worst} (fp16 tolerance 4e-2 on logits of magnitude ~4)"
    # graph replay gives the same answer as the eager steps
    for graph in (1, 1):
        op.set_option("cuda_graph", graph)
        again = op.forward(d_ids, d_lens, out_len, beam_width=K, return_cum_log_probs=1, **targs)
        assert all(torch.equal(a, b) for a, b in zip(again, got)), "graph replay differs from eager beam search"
    free = ref_model.forward_beam(ids, lens, out_len, K, **kw)
    if free["min_margin"] >= 5e-3:
        assert np.array_equal(got[0].cpu().numpy(), free["output_ids"])
    # a sampling request afterwards still works on the same engine (buffers and graph cache are shared)
    got1 = op.forward(d_ids, d_lens, out_len)
    exp1 = ref_model.forward(ids, lens, out_len)
    assert np.array_equal(got1[0].cpu().numpy(), exp1["output_ids"])


def test_engine_beam_search_callback_and_stop_words(cuda):
    dev = cuda
    cfg = tiny_cfg()
    rw = W.make_synthetic(cfg, 1, 0, 1, "cpu", seed=22, keep_plain=True)
    ref_model = oracle_from_rank_weights(cfg, [rw], 1)
    op = _op(cfg, rw, dev, 1)
    ids = np.random.default_rng(9).integers(0, cfg.vocab_size - 1, size=(1, 8)).astype(np.int32)
    free = ref_model.forward_beam(ids, [8], 8, 3)
    raw, par = free["raw_output_ids"], free["parent_ids"]
    tok_b = int(raw[8 + 3, 0])
    tok_a = int(raw[8 + 2, int(par[8 + 3, 0])])
    stop = np.full((1, 2, 2), -1, np.int32)
    stop[0, 0, :], stop[0, 1, 0] = [tok_a, tok_b], 2
    exp = ref_model.forward_beam(ids, [8], 8, 3, stop_words_list=stop)
    seen = []
    got = op.forward(torch.from_numpy(ids).to(dev), torch.tensor([8], dtype=torch.int32, device=dev), 8, beam_width=3,
                     stop_words_list=torch.from_numpy(stop).to(dev), callback=lambda m: seen.append(m))
    assert np.array_equal(got[0].cpu().numpy(), exp["output_ids"])
    assert np.array_equal(got[1].cpu().numpy(), exp["sequence_lengths"])
    assert seen and all(len(m["last_tokens"]) == 1 and len(m["last_tokens"][0]) == 3 for m in seen)
