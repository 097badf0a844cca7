"""Oracle (numpy restatement) of the reference's logit post-processing and
top-k / top-p sampling for beam_width == 1.

TEST INFRASTRUCTURE -- never imported by the product path.

Follows (paths relative to /root/reference/src/fastertransformer):
  * DynamicDecodeLayer<float>::forward        layers/DynamicDecodeLayer.cc:192-495
  * BaseSamplingLayer<T>::forward             layers/sampling_layers/BaseSamplingLayer.cc:255-357
  * TopKSamplingLayer: setup rules            layers/sampling_layers/TopKSamplingLayer.cu:28-78
                       runSampling            layers/sampling_layers/TopKSamplingLayer.cu:189-265
  * select_optional_last_tokens               kernels/select_optional_last_tokens.cu:21-117
  * batchApplyTemperaturePenalty              kernels/sampling_penalty_kernels.cu:115-143
  * batchApplyRepetitionPenalty               kernels/sampling_penalty_kernels.cu:366-425
  * addBiasEndMask / addBiasSoftMax           kernels/sampling_topk_kernels.cu:68-93, sampling_topp_kernels.cu:1296-1345
  * topk_stage1 / topk_stage2_sampling        kernels/sampling_topk_kernels.cu:131-312 (CASE_K table :411-417)
  * TopK_2 / reduce_topk_op_2                 kernels/reduce_kernel_utils.cuh:325-348
  * curand XORWOW, curand_init(seed, 0, 0)    kernels/sampling_topk_kernels.cu:32-55 (generator: CUDA toolkit's
                                              curand_kernel.h, closed library -- its published header algorithm is
                                              restated in `CurandXorwow`)
  * stop_words_criterion / length_criterion   kernels/stop_criteria_kernels.cu:24-156

Tie-breaking restated from the reduction structure: a thread keeps the FIRST
maximum it meets (`elem > u`, strict); cub::BlockReduce (warp shuffle-down,
then warp aggregates in order) combined with `a.u > b.u ? a : b` hands ties to
the HIGHER thread index.  So among equal maxima the winner is the element with
the highest (index % BLOCK_SIZE), then the lowest index.

Parity status: the top-k chain is pinned against the reference's OWN kernels on
the B200 (oracle/_ref/libref_kernels.so, tests/test_ref_kernels_gpu.py: same
seeds, ids / finished / lengths identical over multi-step runs); the tie rule
above is derived from source and confirmed on the GPU (tests/test_sampling_gpu.py).
Pure top-p rows and the stop-word criterion are pinned the same way
(tests/test_ref_kernels_sampling_gpu.py).
"""
from __future__ import annotations

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)
MAX_BLOCKS_PER_BEAM = 8


def topk_block_sizes(max_top_k: int):
    """CASE_K table, sampling_topk_kernels.cu:411-417 -> (BLOCK_SIZE_1, BLOCK_SIZE_2)."""
    if 1 <= max_top_k <= 16:
        return 128, 128
    if max_top_k <= 32:
        return 256, 128
    if max_top_k <= 1024:
        return 256, 256
    raise ValueError("top-k kernel supports 1<=k<=1024")


class CurandXorwow:
    """curandStateXORWOW after curand_init(seed, subsequence=0, offset=0)."""

    M = 0xFFFFFFFF

    def __init__(self, seed: int):
        seed &= 0xFFFFFFFFFFFFFFFF
        s0 = (seed & self.M) ^ 0xAAD26B49
        s1 = (seed >> 32) ^ 0xF7DCEFDD
        t0 = (1099087573 * s0) & self.M
        t1 = (2591861531 * s1) & self.M
        self.d = (6615241 + t1 + t0) & self.M
        self.v = [(123456789 + t0) & self.M, 362436069 ^ t0, (521288629 + t1) & self.M, 88675123 ^ t1,
                  (5783321 + t0) & self.M]

    def next_u32(self) -> int:
        v = self.v
        t = v[0] ^ (v[0] >> 2)
        v[0], v[1], v[2], v[3] = v[1], v[2], v[3], v[4]
        v[4] = ((v[4] ^ ((v[4] << 4) & self.M)) ^ (t ^ ((t << 1) & self.M))) & self.M
        self.d = (self.d + 362437) & self.M
        return (v[4] + self.d) & self.M

    def uniform(self) -> np.float32:
        """curand_uniform: x * 2^-32 + 2^-33 in fp32 (contracted to one fma by nvcc)."""
        x = np.float64(np.float32(self.next_u32()))
        return np.float32(x * np.float64(np.float32(2.3283064e-10)) + np.float64(np.float32(2.3283064e-10)) / 2.0)


def setup_topk_runtime_args(top_k, top_p, batch):
    """TopKSamplingLayer.cu:28-78 (k clipped to 1024; k=0&p=0 -> k=1; k>0&p=0 -> p=1)."""
    ks = np.broadcast_to(np.asarray(top_k, dtype=np.int64).reshape(-1), (batch,)).copy()
    ps = np.broadcast_to(np.asarray(top_p, dtype=np.float32).reshape(-1), (batch,)).copy()
    skip = np.zeros(batch, dtype=bool)
    for i in range(batch):
        k, p = int(ks[i]), np.float32(ps[i])
        if k == 0 and p == 0.0:
            k = 1
        if k > 0 and p == 0.0:
            p = np.float32(1.0)
        ks[i] = min(k, 1024)
        ps[i] = min(max(p, np.float32(0.0)), np.float32(1.0))
        skip[i] = k == 0
    return ks, ps, skip


def select_optional_last_tokens(logits, optional_last_tokens):
    """select_optional_last_tokens.cu:74-83: ids not listed -> -inf.  In place."""
    for b in range(logits.shape[0]):
        allowed = optional_last_tokens[b]
        allowed = allowed[allowed >= 0]
        mask = np.ones(logits.shape[1], dtype=bool)
        mask[allowed] = False
        logits[b, mask] = -np.inf


def apply_temperature(logits, temperature, vocab_size):
    inv = (np.float32(1.0) / (np.asarray(temperature, np.float32) + np.float32(1e-6))).astype(np.float32)
    logits[:, :vocab_size] = (logits[:, :vocab_size] * inv[:, None]).astype(np.float32)
    logits[:, vocab_size:] = -FLT_MAX


def apply_repetition_penalty(logits, penalties, output_ids, input_lengths, max_input_length, step):
    """Multiplicative; every distinct id once; pad gap skipped.  output_ids: [maxlen, B]."""
    for b in range(logits.shape[0]):
        pen = np.float32(penalties[b])
        idx = [i for i in range(step) if not (input_lengths[b] <= i < max_input_length)]
        ids = output_ids[idx, b]
        vals = logits[b, ids].copy()
        vals = np.where(vals < 0, vals * pen, vals / pen).astype(np.float32)
        logits[b, ids] = vals


def add_bias_end_mask(logits, end_ids, finished, vocab_size):
    logits[:, vocab_size:] = -FLT_MAX
    for b in range(logits.shape[0]):
        if finished[b]:
            logits[b, :vocab_size] = -FLT_MAX
            logits[b, end_ids[b]] = FLT_MAX


def softmax_probs(logits):
    """addBiasSoftMax (after the end mask): exp(x-max)/(sum+1e-6), fp32."""
    m = logits.max(axis=1, keepdims=True)
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp((logits - m).astype(np.float32)).astype(np.float32)
    s = e.sum(axis=1, keepdims=True, dtype=np.float32)
    return (e / (s + np.float32(1e-6))).astype(np.float32)


def _argmax_tie(vals, idx, block):
    """Winner among (vals, idx): max value; ties -> highest idx % block, then lowest idx.
    Returns position in the arrays, or -1 if nothing exceeds -FLT_MAX."""
    m = vals.max()
    if not (m > -FLT_MAX):
        return -1
    cand = np.nonzero(vals == m)[0]
    if len(cand) == 1:
        return int(cand[0])
    ci = idx[cand]
    key = (-(ci % block)).astype(np.int64) * (1 << 40) + ci.astype(np.int64)
    return int(cand[np.argmin(key)])


def topk_sampling_row(row, k, p, rng: CurandXorwow, max_top_k, is_prob):
    """One unfinished row of topk_stage1 + topk_stage2_sampling.  row: fp32 [V_padded].
    Returns (token id, value used for log-prob i.e. exp_logit)."""
    bs1, bs2 = topk_block_sizes(max_top_k)
    V = row.shape[0]
    elem = np.arange(V)
    lane_of = (elem // bs1) % MAX_BLOCKS_PER_BEAM
    tmp_ids = np.full(MAX_BLOCKS_PER_BEAM * k, -1, dtype=np.int64)
    tmp_val = np.full(MAX_BLOCKS_PER_BEAM * k, -FLT_MAX, dtype=np.float32)
    for lane in range(MAX_BLOCKS_PER_BEAM):
        sel = np.nonzero(lane_of == lane)[0]
        if len(sel) == 0:
            continue
        vals = row[sel].copy()
        # partial selection is enough: only the k best of a lane can matter
        for ite in range(k):
            w = _argmax_tie(vals, sel, bs1)
            if w < 0:
                break       # total.p stays -1 in the reference; such slots never win stage 2
            tmp_ids[lane * k + ite] = sel[w]
            tmp_val[lane * k + ite] = vals[w]
            vals[w] = -FLT_MAX
    # stage 2
    size = k * MAX_BLOCKS_PER_BEAM
    pos = np.arange(size)
    s_val = tmp_val.copy()
    s_id, s_val2 = [], []
    s_sum = np.float32(0.0)
    s_max = None
    for ite in range(k):
        w = _argmax_tie(s_val, pos, bs2)
        if w < 0:
            w = size - 1 if False else int(np.argmax(s_val))   # degenerate (k > #valid); not exercised
        u = np.float32(s_val[w])
        if ite == 0:
            s_max = u
        s_val[w] = -FLT_MAX
        if not is_prob:
            u = np.float32(np.exp(np.float32(u - s_max)))
        s_id.append(w)
        s_val2.append(u)
        s_sum = np.float32(s_sum + u)
    rand_num = np.float32(np.float32(rng.uniform() * np.float32(p)) * s_sum)
    for i in range(k):
        rand_num = np.float32(rand_num - s_val2[i])
        if rand_num <= 0.0 or i == k - 1:
            return int(tmp_ids[s_id[i]] % V), s_val2[i]
    raise AssertionError


def topp_sampling_row(probs, p, rng: CurandXorwow):
    """One unfinished pure top-p row (top_k == 0): topp_beam_topk_kernel<MAX_K=1> + stable descending sort + topp_sampling
    (kernels/sampling_topp_kernels.cu:801-1004).  probs: fp32 softmax of the row.  Returns (token id, its probability).
    The reference accumulates the sorted probabilities with cub::BlockScan in fp32; this restatement uses the exact
    (float64) prefix sums, so the two can only differ when the draw falls within fp32 rounding of a prefix sum."""
    rand = np.float32(rng.uniform() * np.float32(p))          # drawn before the head check (:917-920)
    top = int(np.argmax(probs))
    if np.float32(probs[top]) >= np.float32(p):               # :842-856
        return top, np.float32(probs[top])
    order = np.argsort(-probs.astype(np.float64), kind="stable")
    cum = np.cumsum(probs[order].astype(np.float64))
    hit = np.nonzero(cum >= np.float64(rand))[0]
    if len(hit) == 0:
        return top, np.float32(probs[top])
    j = int(order[hit[0]])
    return j, np.float32(probs[j])


def stop_words_criterion(output_ids, stop_words, finished, step):
    """stop_criteria_kernels.cu:24-81.  output_ids [maxlen, B] time-major; stop_words [B, 2, n]."""
    B = output_ids.shape[1]
    n = stop_words.shape[2]
    for b in range(B):
        base, offs = stop_words[b, 0], stop_words[b, 1]
        for idx in range(n):
            if offs[idx] < 0:
                continue
            item_end = offs[idx]
            item_start = offs[idx - 1] if idx > 0 else 0
            item_size = item_end - item_start
            if step + 1 < item_size:
                continue
            ok = True
            for t in range(item_size - 1, -1, -1):
                if output_ids[step - (item_size - 1) + t, b] != base[item_start + t]:
                    ok = False
                    break
            if ok:
                finished[b] = True
