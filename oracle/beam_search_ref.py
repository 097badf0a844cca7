"""Oracle: CPU restatement of the reference's online beam search step (beam_width > 1).

TEST INFRASTRUCTURE -- never imported by the product path.

Follows (paths relative to /root/reference/src/fastertransformer):
  * layer wiring          layers/DynamicDecodeLayer.cc:308-408 (beam branch: element 0 of every runtime argument is used for the
                          whole batch -- the per-row tensors are passed down unsliced), layers/beam_search_layers/
                          BaseBeamSearchLayer.cu:170-285 (penalties -> softmax/top-k -> cache indirection update)
  * penalties             kernels/beam_search_penalty_kernels.cu:24-50 (temperature, padded vocabulary -> -FLT_MAX),
                          :84-150 (repetition penalty along the beam's own history: walks parent_ids backwards, skips the pad gap)
  * softmax + top-k       kernels/online_softmax_beamsearch_kernels.cu:366-452 (per row, per vocabulary part: online softmax
                          (max, sum) and the top 2K by (value desc, id asc)), :455-520 (merge of the parts; candidate value =
                          logit - max - log(sum) + cum_log_prob of the row), :101-262 (per batch: K winners among the K x 2K
                          candidates by (score desc, candidate index asc); score = value [/ length^len_penalty] + diversity *
                          (candidate index % K); the stored cum_log_prob is the value WITHOUT length penalty / diversity)
                          kernels/reduce_kernel_utils.cuh:275-322 (TopK insert: ties go to the smaller id)
  * finished rows         online_softmax_beamsearch_kernels.cu:390-398: a finished beam offers end_id at +FLT_MAX and everything
                          else at -FLT_MAX, i.e. exactly one live candidate (end_id, value = its cum_log_prob)
  * update                layers/beam_search_layers/OnlineBeamSearchLayer.cu:24-60 (sequence_length of the slot := the PARENT's,
                          +1 unless the parent had finished; finished := token == end_id; parent_ids / output_ids of the step)
  * cache indirection     BaseBeamSearchLayer.cu:24-52: tgt[b][beam][t] = (t == step) ? beam : src[b][parent][t] for t <= step;
                          rows of beams that are finished AFTER the update are left untouched
  * stop words            kernels/stop_criteria_kernels.cu:24-84 (walks parent_ids for beam_width > 1)
  * output                kernels/decoding_kernels.cu:452-580 (gatherTree with parents: every beam of a batch is walked from the
                          same last level max_b(sequence_length) + 1)
  * initial state         kernels/decoding_kernels.cu:24-60 (cum_log_probs 0 for beam 0, -1e20 for the others)

Pin: tests/test_beam_search_gpu.py runs the reference's own invokeTopkSoftMax / invokeAddBiasApplyPenalties (compiled into
oracle/_ref/libref_kernels.so) on the B200 beside this restatement and beside our kernels.
"""
from __future__ import annotations

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)


def decoding_initialize(batch, beam):
    """finished, cum_log_probs of invokeDecodingInitialize (decoding_kernels.cu:24-60)."""
    cum = np.zeros((batch, beam), dtype=np.float32)
    cum[:, 1:] = np.float32(-1e20)
    return np.zeros(batch * beam, dtype=bool), cum.reshape(-1)


def apply_penalties(logits, step, output_ids, parent_ids, input_lengths, max_input_length, beam, vocab_size, temperature,
                    repetition_penalty):
    """logits [BB, Vp] fp32, modified in place (beam_search_penalty_kernels.cu:171-257 with bias == nullptr, min_length 0)."""
    BB, Vp = logits.shape
    if temperature != 1.0 or vocab_size != Vp:
        inv = np.float32(1.0) / (np.float32(temperature) + np.float32(1e-6))
        logits[:, :vocab_size] = (logits[:, :vocab_size] * inv).astype(np.float32)
        logits[:, vocab_size:] = -FLT_MAX
    if repetition_penalty != 1.0 and step > 0:
        pen = np.float32(repetition_penalty)
        for bb in range(BB):
            batch = bb // beam
            in_len = int(input_lengths[bb])
            idx, val = [], []

            def push(tok):
                x = logits[bb, tok]
                idx.append(tok)
                val.append(x / pen if x > 0 else x * pen)

            # the slot's own last token is filed under level step - 1 and the write-back skips pad-gap levels (:141-147): on the
            # first step of a row shorter than max_input_length it is not penalised
            if not (in_len <= step - 1 < max_input_length):
                push(int(output_ids[step - 1, bb]))
            parent = bb % beam
            for i in range(step - 2, -1, -1):
                if in_len <= i < max_input_length:
                    continue
                parent = int(parent_ids[i, batch * beam + parent])
                push(int(output_ids[i, batch * beam + parent]))
            for t, v in zip(idx, val):                   # every value was computed from the unpenalised logit
                logits[bb, t] = np.float32(v)


def _top(vals, ids, n):
    """first n of (value desc, id asc)."""
    order = np.lexsort((ids, -vals.astype(np.float64)))
    return order[:n]


def row_candidates(row, finished, end_id, cum, k2):
    """2K candidates (absolute-in-row ids, values) of one row (online_softmax_beamsearch_kernels.cu:366-520)."""
    V = row.shape[0]
    if finished:
        x = np.full(V, -FLT_MAX, dtype=np.float32)
        x[end_id] = FLT_MAX
    else:
        x = row
    m = np.float32(x.max())
    with np.errstate(over="ignore", under="ignore"):
        d = np.float32(np.exp((x.astype(np.float64) - np.float64(m))).sum())
        sel = _top(x, np.arange(V), k2)
        val = (x[sel].astype(np.float32) - m) - np.float32(np.log(np.float64(d)))   # -FLT_MAX - FLT_MAX -> -inf for finished rows
        val = (val + np.float32(cum)).astype(np.float32)
    return sel.astype(np.int64), val


def beam_step(logits, step, output_ids, parent_ids, seq_len, finished, cum_log, cache_indir_src, cache_indir_tgt, input_lengths,
              max_input_length, beam, vocab_size, end_id, temperature=1.0, repetition_penalty=1.0, diversity_rate=0.0,
              length_penalty=0.0, stop_words=None, margins=None):
    """One decoding step for every batch.  logits [B*beam, Vp] fp32 (modified in place); output_ids / parent_ids [max_len, B*beam]
    time-major; seq_len / finished / cum_log [B*beam]; cache_indir_* [B*beam, max_len].  Everything is updated in place."""
    BB, Vp = logits.shape
    B = BB // beam
    K = beam
    apply_penalties(logits, step, output_ids, parent_ids, input_lengths, max_input_length, beam, vocab_size, temperature,
                    repetition_penalty)
    old_seq, old_fin = seq_len.copy(), finished.copy()
    for b in range(B):
        cid, cval = [], []
        for j in range(K):
            bb = b * K + j
            ids, vals = row_candidates(logits[bb], bool(old_fin[bb]), end_id, cum_log[bb], 2 * K)
            cid.append(ids + bb * Vp)                    # "absolute" ids, :409 / :506
            cval.append(vals)
        cid, cval = np.concatenate(cid), np.concatenate(cval)
        score = cval.copy()
        if length_penalty != 0.0:
            # :152-157 index finished / sequence_lengths with the BATCH index (vector_id), not batch * beam + beam
            length = int(old_seq[b]) if old_fin[b] else int(old_seq[b]) + 1
            if length != 1:
                score = (score / np.float32(np.power(np.float32(length), np.float32(length_penalty)))).astype(np.float32)
        score = (score + np.float32(diversity_rate) * (np.arange(cid.shape[0]) % K).astype(np.float32)).astype(np.float32)
        win = _top(score, np.arange(cid.shape[0]), K)
        if margins is not None:                          # smallest gap between neighbours among the K + 1 best scores: how decisive the step was
            order = _top(score, np.arange(cid.shape[0]), K + 1)
            with np.errstate(invalid="ignore"):
                margins.append(float(np.min(np.abs(np.diff(score[order].astype(np.float64))))))
        for j in range(K):
            bb = b * K + j
            word = int(cid[win[j]])
            parent = (word // Vp) % K
            tok = word % Vp
            cum_log[bb] = cval[win[j]]
            seq_len[bb] = old_seq[b * K + parent] + (0 if old_fin[b * K + parent] else 1)
            finished[bb] = tok == end_id
            parent_ids[step, bb] = parent
            output_ids[step, bb] = tok
    # cache indirection (BaseBeamSearchLayer.cu:24-52), with the finished flags of AFTER the update
    for bb in range(BB):
        if finished[bb]:
            continue
        b, j = divmod(bb, K)
        parent = int(parent_ids[step, bb])
        n = min(step + 1, cache_indir_src.shape[1])
        cache_indir_tgt[bb, :n] = cache_indir_src[b * K + parent, :n]
        if step < cache_indir_tgt.shape[1]:
            cache_indir_tgt[bb, step] = j
    if stop_words is not None:
        stop_words_criterion_beams(output_ids, parent_ids, np.asarray(stop_words), finished, step, beam)


def stop_words_criterion_beams(output_ids, parent_ids, stop_words, finished, step, beam):
    """stop_criteria_kernels.cu:24-84; stop_words [B, 2, n]."""
    BB = output_ids.shape[1]
    n = stop_words.shape[2]
    for bb in range(BB):
        b, j = divmod(bb, beam)
        words, offs = stop_words[b, 0], stop_words[b, 1]
        for i in range(n):
            if offs[i] < 0:
                continue
            end, start = int(offs[i]), int(offs[i - 1]) if i > 0 else 0
            size = end - start
            if step + 1 < size:
                continue
            ok, parent = True, j
            for t in range(size - 1, -1, -1):
                lvl = step - (size - 1) + t
                if output_ids[lvl, b * beam + parent] != words[start + t]:
                    ok = False
                    break
                parent = int(parent_ids[lvl, b * beam + parent])
                if parent < 0 or parent >= beam:
                    ok = False
                    break
            if ok:
                finished[bb] = True


def gather_tree(step_ids, parent_ids, seq_len, input_lengths, max_input_length, max_time, end_id, beam):
    """gatherTree with parents (decoding_kernels.cu:452-580).  Returns output_ids [B, beam, max_time], sequence_lengths [B, beam]."""
    BB = step_ids.shape[1]
    B = BB // beam
    beams = np.zeros((max_time, BB), dtype=np.int64)
    out_len = np.zeros((B, beam), dtype=np.int32)
    for b in range(B):
        max_len = max(int(seq_len[b * beam + j]) + 1 for j in range(beam))
        for j in range(beam):
            out_len[b, j] = int(seq_len[b * beam + j]) + 1
        msl = min(max_time, max_len)
        if msl <= 0:
            continue
        for j in range(beam):
            i = b * beam + j
            input_len = int(input_lengths[i])
            pad = max_input_length - input_len
            beams[msl - 1 - pad, i] = step_ids[msl - 1, i]
            parent = int(parent_ids[msl - 1, i]) % beam
            found_bad = False
            for level in range(msl - 2, -1, -1):
                if input_len <= level < max_input_length:
                    continue
                tgt = level - pad if level >= max_input_length else level
                if parent < 0 or parent > beam:
                    beams[tgt, i] = end_id
                    parent = -1
                    found_bad = True
                else:
                    beams[tgt, i] = step_ids[level, b * beam + parent]
                    parent = int(parent_ids[level, b * beam + parent]) % beam
            for index in range(max_len - pad, max_time):
                beams[index, i] = end_id
            if not found_bad:
                fin = False
                start = 1 if max_input_length == 0 else max_input_length
                for time in range(start, msl):
                    if fin:
                        beams[time, i] = end_id
                    elif beams[time, i] == end_id:
                        fin = True
    return beams.T.reshape(B, beam, max_time).astype(np.int32), out_len
