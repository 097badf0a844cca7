"""Bit-exact parity of the sampling stack (C ABI: ftcf_sampling_step / ftcf_gather_output) against oracle/sampling_ref.py."""
import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi
from oracle import gptneox_ref as R
from oracle import sampling_ref as S
from helpers import stream

pytestmark = pytest.mark.gpu


def _run_steps(lib, cuda, B, V, Vp, max_in, out_len, lens, top_k, top_p, temperature, rep, seeds, want_probs, logits_fn,
               optional_last=None, stop_words=None, end_id=None, tie_heavy=False):
    """Drives `out_len` sampling steps on both sides with the same per-step logits; returns both states."""
    end_id = V - 1 if end_id is None else end_id
    max_len = max_in + out_len
    rng = np.random.default_rng(1234)
    prompt = rng.integers(0, V - 1, size=(B, max_in))
    out_ids = np.zeros((max_len, B), dtype=np.int64)
    out_ids[:max_in] = prompt.T
    ks, ps, _ = S.setup_topk_runtime_args(top_k, top_p, B)
    max_top_k = max(1, int(ks.max()))
    temp = np.broadcast_to(np.asarray(temperature, np.float32).reshape(-1), (B,)).copy()
    reps = np.broadcast_to(np.asarray(rep, np.float32).reshape(-1), (B,)).copy()
    seeds = np.broadcast_to(np.asarray(seeds, np.int64).reshape(-1), (B,)).copy()

    # ---- device state
    dev = cuda
    d_ids = torch.from_numpy(out_ids.astype(np.int32)).to(dev)
    d_seq = torch.full((B,), max_in - 1, dtype=torch.int32, device=dev)
    d_fin = torch.zeros(B, dtype=torch.uint8, device=dev)
    d_cum = torch.zeros(B, dtype=torch.float32, device=dev)
    d_len = torch.tensor(lens, dtype=torch.int32, device=dev)
    d_k = torch.from_numpy(ks.astype(np.int32)).to(dev)
    d_p = torch.from_numpy(ps.astype(np.float32)).to(dev)
    d_t = torch.from_numpy(temp).to(dev)
    d_r = torch.from_numpy(reps).to(dev)
    d_step = torch.tensor([max_in], dtype=torch.int32, device=dev)
    d_seeds = torch.from_numpy(seeds.astype(np.uint64).view(np.int64)).to(dev)
    states = torch.zeros(B * lib.ftcf_curand_state_bytes(), dtype=torch.uint8, device=dev)
    capi.check(lib.ftcf_curand_init(states.data_ptr(), d_seeds.data_ptr(), B, stream()))
    ws = torch.zeros(lib.ftcf_sampling_workspace_bytes(B, Vp, max_top_k) + B * max_len * 4 + 256, dtype=torch.uint8, device=dev)
    flag = torch.zeros(2, dtype=torch.int32, device=dev)
    d_logits = torch.empty(B, Vp, dtype=torch.float32, device=dev)
    d_last = torch.from_numpy(optional_last.astype(np.int32)).to(dev) if optional_last is not None else None
    d_stop = torch.from_numpy(stop_words.astype(np.int32)).to(dev) if stop_words is not None else None
    sp = capi.SamplingParams(d_logits.data_ptr(), d_ids.data_ptr(), d_seq.data_ptr(), d_fin.data_ptr(), d_cum.data_ptr(),
                             d_len.data_ptr(), d_k.data_ptr(), d_p.data_ptr(),
                             d_t.data_ptr() if not np.all(temp == 1.0) else None,
                             d_r.data_ptr() if not np.all(reps == 1.0) else None,
                             d_last.data_ptr() if d_last is not None else None,
                             d_stop.data_ptr() if d_stop is not None else None,
                             states.data_ptr(), d_step.data_ptr(), flag.data_ptr(), ws.data_ptr(),
                             B, V, Vp, max_top_k, d_last.shape[1] if d_last is not None else 0,
                             d_stop.shape[2] if d_stop is not None else 0, max_in, max_len, end_id, 1 if want_probs else 0,
                             1 if (ks == 0).any() else 0)

    # ---- oracle state
    seq_len = np.full(B, max_in - 1, dtype=np.int64)
    finished = np.zeros(B, dtype=bool)
    cum = np.zeros(B, dtype=np.float32)
    rngs = [S.CurandXorwow(int(s)) for s in seeds]
    lens_np = np.asarray(lens, dtype=np.int64)

    for step in range(max_in, max_len):
        logits = logits_fn(step).astype(np.float32)
        d_logits.copy_(torch.from_numpy(logits))
        capi.check(lib.ftcf_sampling_step(sp, stream()))
        # oracle, same order as gptneox_ref.GptNeoXRef.forward
        lg = logits.copy()
        if step == max_in and optional_last is not None:
            S.select_optional_last_tokens(lg, optional_last)
        if not np.all(temp == 1.0):
            S.apply_temperature(lg, temp, V)
        if step > 1 and not np.all(reps == 1.0):
            S.apply_repetition_penalty(lg, reps, out_ids, lens_np, max_in, step)
        S.add_bias_end_mask(lg, np.full(B, end_id), finished, V)
        if want_probs:
            lg = S.softmax_probs(lg)
        for b in range(B):
            if finished[b]:
                out_ids[step, b] = end_id
                continue
            if int(ks[b]) == 0:
                pr = lg[b] if want_probs else S.softmax_probs(lg[b:b + 1])[0]
                tok, val = S.topp_sampling_row(pr, ps[b], rngs[b])
            else:
                tok, val = S.topk_sampling_row(lg[b], int(ks[b]), ps[b], rngs[b], max_top_k, want_probs)
            out_ids[step, b] = tok
            if want_probs:
                cum[b] += np.float32(np.log(val))
            seq_len[b] += 1
            finished[b] = tok == end_id
        if stop_words is not None:
            S.stop_words_criterion(out_ids, stop_words, finished, step)
        torch.cuda.synchronize()
        got = d_ids[step].cpu().numpy()
        assert np.array_equal(got, out_ids[step]), f"step {step}: device {got} oracle {out_ids[step]}"
        assert np.array_equal(d_fin.cpu().numpy().astype(bool), finished), f"step {step}: finished flags differ"
        assert int(d_step.item()) == step + 1
        assert int(flag[0].item()) == int(finished.sum()) and int(flag[1].item()) == step
    assert np.array_equal(d_seq.cpu().numpy(), seq_len)
    if want_probs:
        np.testing.assert_allclose(d_cum.cpu().numpy(), cum, rtol=1e-4, atol=1e-4)
    return d_ids, d_seq, d_len, out_ids, seq_len


def _random_logits(B, Vp, scale=3.0, seed=0):
    def fn(step):
        return np.random.default_rng(seed * 7919 + step).normal(0, scale, size=(B, Vp))
    return fn


@pytest.mark.parametrize("want_probs", [0, 1])
def test_greedy(lib, cuda, want_probs):
    B, V = 4, 1000
    _run_steps(lib, cuda, B, V, V, 6, 5, [6, 3, 6, 1], 1, 0.0, 1.0, 1.0, 0, want_probs, _random_logits(B, V))


@pytest.mark.parametrize("k,p", [(4, 1.0), (40, 0.9), (16, 0.5), (17, 1.0), (33, 0.3), (200, 0.95)])
def test_topk_topp_sampling_matches_curand_stream(lib, cuda, k, p):
    B, V, Vp = 3, 5000, 5008
    _run_steps(lib, cuda, B, V, Vp, 4, 6, [4, 2, 3], k, p, 1.0, 1.0, [7, 7, 123456789012], 1, _random_logits(B, Vp, seed=k))


def test_batch_varying_k_p_temperature_repetition(lib, cuda):
    B, V, Vp = 4, 3000, 3000
    _run_steps(lib, cuda, B, V, Vp, 8, 8, [8, 5, 8, 2], np.array([1, 5, 40, 0]), np.array([0.0, 0.7, 0.9, 0.0], np.float32),
               np.array([1.0, 0.2, 0.7, 1.3], np.float32), np.array([1.0, 1.1, 1.3, 1.0], np.float32), [1, 2, 3, 4], 1,
               _random_logits(B, Vp, seed=5))


def test_ties_follow_the_reference_order(lib, cuda):
    # heavily quantised logits: many exact ties inside and across the 8 vocabulary slices
    B, V = 2, 4096

    def fn(step):
        return np.round(np.random.default_rng(step).normal(0, 1.0, size=(B, V)) * 2) / 2
    _run_steps(lib, cuda, B, V, V, 3, 6, [3, 3], 8, 1.0, 1.0, 1.0, 5, 0, fn)
    _run_steps(lib, cuda, B, V, V, 3, 6, [3, 3], 1, 0.0, 1.0, 1.0, 5, 0, fn)


def test_end_id_finishes_and_masks(lib, cuda):
    B, V = 3, 512
    end_id = 77

    def fn(step):
        lg = np.random.default_rng(step).normal(0, 1, size=(B, V))
        if step >= 6:
            lg[1, end_id] = 50.0          # row 1 ends at step 6
        return lg
    d_ids, d_seq, d_len, out_ids, seq_len = _run_steps(lib, cuda, B, V, V, 4, 6, [4, 4, 2], 1, 0.0, 1.0, 1.0, 0, 1, fn, end_id=end_id)
    # output gather against the oracle's gatherTree restatement
    max_in, max_len = 4, 10
    g = torch.empty(B, max_len, dtype=torch.int32, device=cuda)
    gl = torch.empty(B, dtype=torch.int32, device=cuda)
    capi.check(lib.ftcf_gather_output(g.data_ptr(), gl.data_ptr(), d_ids.data_ptr(), d_seq.data_ptr(), d_len.data_ptr(), B, max_in, max_len,
                                      end_id, stream()))
    torch.cuda.synchronize()
    ref, ref_len = R.gather_output(out_ids, seq_len, np.array([4, 4, 2]), max_in, max_len, end_id)
    assert np.array_equal(g.cpu().numpy(), ref[:, 0, :])
    assert np.array_equal(gl.cpu().numpy(), ref_len[:, 0])


def test_optional_last_tokens_and_stop_words(lib, cuda):
    B, V = 2, 600
    optional_last = np.array([[5, 9, 300, -1], [17, -1, -1, -1]])
    # stop phrases: row 0 stops on the bigram (11, 12); row 1 on token 13
    stop = np.full((B, 2, 3), -1, dtype=np.int64)
    stop[0, 0, :2] = [11, 12]
    stop[0, 1, 0] = 2
    stop[1, 0, 0] = 13
    stop[1, 1, 0] = 1

    def fn(step):
        lg = np.random.default_rng(step).normal(0, 1, size=(B, V))
        if step == 5:
            lg[0, 11] = 40.0
        if step == 6:
            lg[0, 12] = 40.0
            lg[1, 13] = 40.0
        return lg
    _run_steps(lib, cuda, B, V, V, 4, 6, [4, 3], 1, 0.0, 1.0, 1.0, 0, 1, fn, optional_last=optional_last, stop_words=stop)


@pytest.mark.parametrize("p", [0.3, 0.9, 1.0])
@pytest.mark.parametrize("want_probs", [False, True])
def test_pure_top_p_sampling(lib, cuda, p, want_probs):
    """top_k = 0, top_p > 0: nucleus sampling (TopPSamplingLayer) -- ids exact against the oracle's sorted walk."""
    B, V, Vp = 3, 1000, 1008
    _run_steps(lib, cuda, B, V, Vp, 4, 24, [4, 2, 3], [0], [p], [1.0], [1.0], [11, 12, 13], want_probs,
               _random_logits(B, Vp, scale=2.0, seed=int(p * 10)))


def test_top_p_head_shortcut_and_mixed_batch(lib, cuda):
    """Peaked rows take the largest probability directly (topp_beam_topk_kernel); top-k and top-p rows share a batch."""
    B, V, Vp = 4, 512, 512

    def logits(step):
        lg = _random_logits(B, Vp, scale=1.0, seed=100)(step)
        lg[0, (7 * step) % V] = 30.0          # row 0: one token holds ~all the mass
        return lg
    _run_steps(lib, cuda, B, V, Vp, 3, 16, [3, 3, 1, 2], [0, 5, 0, 1], [0.8, 0.9, 0.95, 0.0], [1.0, 0.7, 1.3, 1.0], [1.0, 1.2, 1.0, 1.0],
               [1, 2, 3, 4], True, logits)


def test_top_p_ties_follow_id_order(lib, cuda):
    """Equal probabilities are walked in ascending id order (the reference's radix sort is stable)."""
    B, V, Vp = 2, 64, 64

    def logits(step):
        lg = np.zeros((B, Vp), dtype=np.float32)      # uniform rows: every prefix sum is a tie group
        lg[1, :8] = 1.0
        return lg
    _run_steps(lib, cuda, B, V, Vp, 2, 20, [2, 2], [0], [0.9], [1.0], [1.0], [5, 6], False, logits)
