"""CPU checks of the beam-search oracle (oracle/beam_search_ref.py): properties the reference's algorithm guarantees, so that the
GPU parity tests (tests/test_beam_search_gpu.py) compare against a restatement that is at least self-consistent."""
import numpy as np
import pytest

from oracle import beam_search_ref as BS
from helpers_beam import CASES, OracleRun


@pytest.mark.parametrize("case", CASES[:4] + CASES[5:], ids=lambda c: f"B{c['B']}K{c['K']}V{c['V']}")
def test_oracle_run_properties(case):
    case = dict(case)
    B, K, V, Vp, lens = (case.pop(k) for k in ("B", "K", "V", "Vp", "lens"))
    max_in, out_len = 6, 9
    run = OracleRun(B, K, V, Vp, max_in, out_len, lens, V - 1, seed=3, **case)
    plain = not any(k in case for k in ("diversity_rate", "length_penalty"))
    mixed = False
    for step in range(max_in, max_in + out_len):
        x = run.logits(step, 17)
        fin_before, cum_before, seq_before = run.fin.copy(), run.cum.copy(), run.seq.copy()
        run.advance(x, step)
        par = run.par[step].reshape(B, K)
        tok = run.ids[step].reshape(B, K)
        assert ((0 <= par) & (par < K)).all() and ((0 <= tok) & (tok < V)).all()      # never a padded-vocabulary id
        if step == max_in:
            assert (par == 0).all()                                   # only beam 0 is alive at the first step (cum -1e20 elsewhere)
            if plain and "temperature" not in case and "repetition_penalty" not in case:
                for b in range(B):
                    order = np.lexsort((np.arange(Vp), -x[b * K].astype(np.float64)))
                    order = order[order < V][:K] if Vp == V else order[:K]
                    assert np.array_equal(tok[b], order)
        if plain:
            c = run.cum.reshape(B, K)
            assert (np.diff(c, axis=1) <= 0).all()                    # winners come out best first
        for bb in range(B * K):
            p = (bb // K) * K + int(run.par[step, bb])
            if fin_before[p]:                                         # a finished beam only continues with end_id, score unchanged
                assert run.ids[step, bb] == V - 1 and run.fin[bb] and run.cum[bb] == cum_before[p] and run.seq[bb] == seq_before[p]
            else:
                assert run.seq[bb] == seq_before[p] + 1
                assert run.cum[bb] <= cum_before[p] + 1e-6            # log-probabilities are <= 0
            if not run.fin[bb]:
                tgt = run.ind[1 - (step - max_in) % 2][bb]
                assert tgt[step] == bb % K and (tgt[:max_in] == 0).all()
        mixed |= bool(run.fin.any() and not run.fin.all())
    # finished and live beams side by side at some step (not with a diversity bonus: after the length normalisation the
    # + diversity * (candidate index % K) term outweighs the end_id candidate, which is always index 0 of its beam)
    assert mixed or "diversity_rate" in case


def test_oracle_gather_tree_follows_parents():
    B, K, V, max_in, out_len = 2, 3, 300, 5, 7
    run = OracleRun(B, K, V, V, max_in, out_len, [5, 3], V - 1, seed=8)
    for step in range(max_in, max_in + out_len):
        run.advance(run.logits(step, 5), step)
    out, out_len_a = BS.gather_tree(run.ids, run.par, run.seq, run.lens, max_in, run.max_len, V - 1, K)
    for b in range(B):
        n_in = int(run.lens[b * K])
        for j in range(K):
            # follow the parents by hand from the last level
            path, slot = [], j
            for level in range(run.max_len - 1, max_in - 1, -1):
                path.append(int(run.ids[level, b * K + slot]))
                slot = int(run.par[level, b * K + slot])
            path = path[::-1]
            if V - 1 in path:                                         # everything after the first end_id is end_id
                k = path.index(V - 1)
                path = path[:k] + [V - 1] * (len(path) - k)
            want = list(run.ids0[:n_in, b * K]) + path
            want = want + [V - 1] * (run.max_len - len(want))
            assert list(out[b, j]) == want
            assert out_len_a[b, j] == run.seq[b * K + j] + 1


def test_oracle_repetition_penalty_walks_the_beam_history():
    """Two beams with different histories must penalise different tokens."""
    B, K, V, max_in = 1, 2, 64, 3
    run = OracleRun(B, K, V, V, max_in, 4, [3], V - 1, seed=1)
    run.advance(run.logits(max_in, 2), max_in)
    step = max_in + 1
    x = np.full((K, V), 1.0, np.float32)
    BS.apply_penalties(x, step, run.ids, run.par, run.lens, max_in, K, V, 1.0, 2.0)
    for j in range(K):
        hist = {int(run.ids[step - 1, j])} | {int(t) for t in run.ids0[:max_in, 0]}
        assert {int(i) for i in np.nonzero(x[j] == 0.5)[0]} == hist


def test_oracle_forward_beam_output_contract():
    """Model-level restatement (GptNeoXRef.forward_beam): output shapes and the gatherTree contract on a ragged batch -- every beam
    starts with its request's prompt, the pad gap is gone, sequence_lengths counts max_input_length + generated (pad gap NOT
    subtracted, as for sampling), cum_log_probs come out best first, and deciding on the run's own traced logits reproduces it."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fastertransformer4codefuse_b200 import weights as W
    from helpers import oracle_from_rank_weights, tiny_cfg
    cfg = tiny_cfg()
    rw = W.make_synthetic(cfg, 1, 0, 1, "cpu", seed=21, keep_plain=True)
    ref = oracle_from_rank_weights(cfg, [rw], 1)
    g = np.random.default_rng(12)
    lens, S, out_len, K = [7, 4], 7, 5, 3
    ids = g.integers(0, cfg.vocab_size - 1, size=(2, S)).astype(np.int32)
    ids[1, 4:] = cfg.end_id
    res = ref.forward_beam(ids, lens, out_len, K)
    out, sl, cum = res["output_ids"], res["sequence_lengths"], res["cum_log_probs"]
    assert out.shape == (2, K, S + out_len) and sl.shape == (2, K) and cum.shape == (2, K)
    for b in range(2):
        for j in range(K):
            assert np.array_equal(out[b, j, :lens[b]], ids[b, :lens[b]])
            n_gen = int(sl[b, j]) - S
            assert 1 <= n_gen <= out_len
            assert (out[b, j, lens[b] + n_gen:] == cfg.end_id).all()
        assert (np.diff(cum[b]) <= 0).all()
    again = ref.forward_beam(ids, lens, out_len, K, decide_on_logits=np.stack(res["logits"]))
    assert np.array_equal(again["output_ids"], out) and np.array_equal(again["sequence_lengths"], sl)
