"""Oracle: CPU restatement of the reference's GPT-NeoX (CodeFuse) request path --
prefill, per-token decode, LM head, sampling, output gather -- with the
reference's fp16 rounding points made explicit (fp32 math, `.half()` where the
reference stores or computes in half).

TEST INFRASTRUCTURE -- never imported by the product path.

Follows (paths relative to /root/reference/src/fastertransformer):
  * request loop          models/gptneox/GptNeoX.cc:385-1052 (buffers :84-156, vocab padding :319-323,
                          output gather setOutputTensors :1090-1181 + kernels/decoding_kernels.cu:519-560)
  * prefill layers        models/gptneox/GptNeoXContextDecoder.cc:223-512
  * decode layers         models/gptneox/GptNeoXDecoder.cc:197-389 (parallel residual :301-360, sequential :313-331,361-368)
  * decode attention      kernels/decoder_masked_multihead_attention/decoder_masked_multihead_attention_template.hpp:1099-1919
                          (bias add, NeoX rotary :1312-1365, masked softmax :1498-1643, fp16 probabilities :1643)
  * rotary                kernels/decoder_masked_multihead_attention_utils.h:1325-1337
  * prefill attention     layers/attention_layers/GptContextAttentionLayer.cc:25-403,
                          kernels/unfused_attention_kernels.cu:255-333 (masked softmax, -10000 mask), :1326-1484 (bias+rotary)
  * LayerNorm             kernels/layernorm_kernels.cu:158-286 (fp32 statistics, normalisation in half2)
  * GELU                  kernels/activation_kernels.cu:50-72 (fp16 path), cutlass epilogue
                          cutlass_extensions/include/cutlass_extensions/epilogue/thread/ft_fused_activations.h:61-84 (int8 path)
  * residual              kernels/add_residual_kernels.cu:116-176
  * INT8 GEMM rounding    cutlass_extensions/include/cutlass_extensions/gemm/threadblock/dq_mma_multistage.h:475-522
                          (u8 -> fp16, times scale IN fp16, then mma with fp32 accumulate)
  * weight list order     th_op/gptneox/GptNeoXOp.h:121-174, examples/pytorch/codefuse/codefuse_example.py:182-419
  * sampling              oracle/sampling_ref.py
  * beam search           oracle/beam_search_ref.py (forward_beam below: tiling, cache indirection, gatherTree with parents)

Parity status.  The reference as a whole cannot be built or run for this path on this image or on sm_100 (SURVEY.md
section 8c), but its kernels can: oracle/Makefile compiles the reference's own .cu files (LayerNorm, residual add, decode
attention, prefill bias + rotary + split, masked softmax, top-k / top-p sampling, penalties, stop criteria, beam-search
penalties + softmax/top-k, gatherTree) for sm_100a into oracle/_ref/libref_kernels.so, and tests/test_ref_kernels*_gpu.py,
tests/test_beam_search_gpu.py run them on the B200 beside our kernels AND beside this restatement.  Pinned that way: LayerNorm
(bit-level: > 97 % of the elements identical, the rest one fp16 ulp from the fp32 reduction order), the parallel-residual add
(bit-exact), the decode attention incl. bias + NeoX rotary + cache append (fp16 tolerance stated in the test), the prefill
bias + rotary + split, the prefill softmax chain, the sampling chain and the beam search (token ids / parents / finished flags /
lengths identical over multi-step seeded runs), the output gather with and without parents.  The quantiser is pinned by the
reference's object code on the CPU (oracle/_ref/libref_quant.so) and its KATs; the model wiring by HuggingFace
GPTNeoXForCausalLM goldens (tests/golden/, tests/golden/make_golden.py); the request loop as a whole by the reference's
UNCHANGED driver (oracle/_ref/codefuse_example.compiled, tests/test_driver_gpu.py).  Still PARITY UNPINNED by the reference:
only the INT8 GEMM rounding (its CUTLASS kernel refuses sm >= 90); it is held by the exact dequant round trip and the
reference's own tolerance (rtol 1e-3 / atol 2e-3) instead.

Tensor parallelism is emulated: `ranks` weight sets are evaluated one after the
other and summed where the reference all-reduces.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import sampling_ref as S


def h(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 and come back to fp32 (a reference `half` store)."""
    return x.to(torch.float16).to(torch.float32)


@dataclass
class RefConfig:
    head_num: int
    size_per_head: int
    inter_size: int
    layer_num: int
    vocab_size: int
    rotary_embedding_dim: int
    start_id: int
    end_id: int
    tensor_para_size: int = 1
    int8_mode: int = 0
    use_gptj_residual: bool = True
    layernorm_eps: float = 1e-5

    @property
    def hidden(self):
        return self.head_num * self.size_per_head


@dataclass
class RankWeights:
    """One TP rank's tensors in GptNeoXOp order (index = field * L + layer)."""
    w: List[torch.Tensor]                       # fp16, 12L + 4
    q: List[Optional[np.ndarray]] = field(default_factory=list)      # plain int8 [k, n], 4L
    scale: List[Optional[torch.Tensor]] = field(default_factory=list)  # fp16 [n], 4L


# ------------------------------------------------------------------ primitives
def layernorm_ref(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    """layernorm_kernels.cu:158-286.  x fp32-held fp16 values [m, n]."""
    n = x.shape[-1]
    mean = x.sum(-1, keepdim=True) / n
    var = (x * x).sum(-1, keepdim=True) / n - mean * mean + eps
    rstd = torch.rsqrt(var)
    mean_h, rstd_h = h(mean), h(rstd)
    # (x - mean) and (* rstd) round to fp16 each; (* gamma + beta) is ONE fused multiply-add in the reference's object code
    # (SASS of generalAddBiasResidualLayerNormOpt2<half2>: HADD2, HMUL2, HFMA2) -- pinned against that kernel on the GPU by
    # tests/test_ref_kernels_gpu.py.  float64 holds the fp16 x fp16 + fp16 result exactly, so one rounding to fp16 follows.
    t = h(h(x - mean_h) * rstd_h)
    return (t.double() * gamma.double() + beta.double()).half().float()


def gelu_f32(x: torch.Tensor) -> torch.Tensor:
    return x * (0.5 * (1.0 + torch.tanh(0.7978845608028654 * (x + 0.044715 * x * x * x))))


def gelu_half2(val: torch.Tensor) -> torch.Tensor:
    """activation_kernels.cu:59-72 on fp16 values held in fp32."""
    pow3 = h(val * h(val * val))
    cdf = 0.5 * (1.0 + torch.tanh(0.7978845608028654 * (val + 0.044715 * pow3)))
    return h(val * h(cdf))


def rotary_coef(pos: torch.Tensor, rot: int) -> (torch.Tensor, torch.Tensor):
    """decoder_masked_multihead_attention_utils.h:1325-1329; pos [*] -> cos/sin [*, rot/2]."""
    i = torch.arange(rot // 2, dtype=torch.float32)
    inv = pos.to(torch.float32)[..., None] / torch.pow(torch.tensor(10000.0), 2.0 * i / rot)
    return torch.cos(inv), torch.sin(inv)


def apply_rotary_neox(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, rot: int) -> torch.Tensor:
    """NeoX pairing (i, i + rot/2) (template.hpp:1318-1353); x [..., Dh] fp16-valued."""
    half = rot // 2
    a, b = x[..., :half], x[..., half:rot]
    ra = h(cos * a - sin * b)
    rb = h(cos * b + sin * a)
    return torch.cat([ra, rb, x[..., rot:]], dim=-1)


def gather_output(step_ids, seq_len, input_lengths, max_input_length, max_time, end_id):
    """gatherTree for beam_width == 1 (decoding_kernels.cu:452-580): removes the pad gap
    [input_len, max_input_length), fills the tail with end_id, and reports
    sequence_length = internal length + 1 (i.e. max_input_length + #generated: the pad gap is
    NOT subtracted, :497-501 with max_prefix_soft_prompt_length == 0)."""
    B = step_ids.shape[1]
    beams = np.zeros((max_time, B), dtype=np.int64)
    out_len = np.zeros((B, 1), dtype=np.int32)
    for b in range(B):
        input_len = int(input_lengths[b])
        max_len = int(seq_len[b]) + 1
        out_len[b, 0] = max_len
        msl = min(max_time, max_len)
        if msl <= 0:
            continue
        pad = max_input_length - input_len
        beams[msl - 1 - pad, b] = step_ids[msl - 1, b]
        for level in range(msl - 2, -1, -1):
            if input_len <= level < max_input_length:
                continue
            tgt = level - pad if level >= max_input_length else level
            beams[tgt, b] = step_ids[level, b]
        for index in range(max_len - pad, max_time):
            beams[index, b] = end_id
        fin = False
        start = 1 if max_input_length == 0 else max_input_length
        for time in range(start, msl):
            if fin:
                beams[time, b] = end_id
            elif beams[time, b] == end_id:
                fin = True
    return beams.T.reshape(B, 1, max_time).astype(np.int32), out_len


class GptNeoXRef:
    def __init__(self, cfg: RefConfig, ranks: List[RankWeights]):
        assert len(ranks) == cfg.tensor_para_size
        self.cfg = cfg
        self.ranks = ranks
        L = cfg.layer_num
        self._eff = []          # per rank: dict (kind, layer) -> fp32 [k, n] effective weight
        for rw in ranks:
            eff = {}
            for kind, widx in enumerate((2, 4, 6, 8)):
                for l in range(L):
                    if cfg.int8_mode == 1:
                        qv = torch.from_numpy(rw.q[kind * L + l].astype(np.float32))
                        sc = rw.scale[kind * L + l].float()
                        eff[(kind, l)] = h(qv * sc[None, :])      # fp16(q) * scale in fp16
                    else:
                        eff[(kind, l)] = rw.w[widx * L + l].float()
            self._eff.append(eff)
        t = cfg.tensor_para_size
        v = cfg.vocab_size
        self.vocab_padded = int(math.ceil(math.ceil(v / t) / 8.0) * 8 * t)      # GptNeoX.cc:319-323

    # ---- GEMMs
    def _gemm(self, x, r, kind, l, bias=None, act=False):
        acc = x @ self._eff[r][(kind, l)]
        if self.cfg.int8_mode == 1:
            if bias is not None:
                acc = acc + bias.float()
            if act:
                acc = gelu_f32(acc)
            return h(acc)
        y = h(acc)
        if bias is not None and act:
            return gelu_half2(h(y + bias.float()))
        return y

    def _W(self, r, f, l):
        return self.ranks[r].w[f * self.cfg.layer_num + l]

    # ---- one layer on [m, hidden] activations; attn_fn(r, l, qkv[m, 3h/t]) -> ctx [m, h/t]
    def _layer(self, x, l, attn_fn):
        cfg = self.cfg
        t = cfg.tensor_para_size
        L = cfg.layer_num
        outs = []
        if cfg.use_gptj_residual:
            for r in range(t):
                n1 = layernorm_ref(x, self._W(r, 1, l), self._W(r, 0, l), cfg.layernorm_eps)
                qkv = self._gemm(n1, r, 0, l)
                ctx = attn_fn(r, l, qkv)
                attn = self._gemm(ctx, r, 1, l)
                n2 = layernorm_ref(x, self._W(r, 11, l), self._W(r, 10, l), cfg.layernorm_eps)
                inter = self._gemm(n2, r, 2, l, bias=self._W(r, 7, l), act=True)
                ffn = self._gemm(inter, r, 3, l)
                # add_residual_kernels.cu:116-152: (half)(x/t) + ffn + attn + bias, fp16 adds
                xs = h(x / t) if t > 1 else x
                o = h(h(h(ffn + attn) + self._W(r, 9, l).float()) + xs)
                outs.append(o)
            return h(torch.stack(outs).sum(0))
        # sequential residual
        attn_sum = []
        for r in range(t):
            n1 = layernorm_ref(x, self._W(r, 1, l), self._W(r, 0, l), cfg.layernorm_eps)
            qkv = self._gemm(n1, r, 0, l)
            ctx = attn_fn(r, l, qkv)
            attn_sum.append(self._gemm(ctx, r, 1, l))
        attn = h(torch.stack(attn_sum).sum(0))
        # invokeGeneralAddBiasResidualPreLayerNorm: y = attn + bias + x (stored), n2 = LN(y)
        y = h((self._W(0, 5, l).float() + x) + attn)                  # fp32 sum, one rounding (layernorm_kernels.cu:196-232)
        ffn_sum = []
        for r in range(t):
            n2 = layernorm_ref(y, self._W(r, 11, l), self._W(r, 10, l), cfg.layernorm_eps)
            inter = self._gemm(n2, r, 2, l, bias=self._W(r, 7, l), act=True)
            ffn_sum.append(self._gemm(inter, r, 3, l))
        ffn = h(torch.stack(ffn_sum).sum(0))
        return h(h(ffn + y) + self._W(0, 9, l).float())          # add_residual_kernels.cu:44-46

    # ---- request
    @torch.no_grad()
    def forward(self, input_ids, input_lengths, output_len, top_k=None, top_p=None, temperature=None,
                repetition_penalty=None, random_seed=None, stop_words_list=None, optional_last_tokens=None,
                return_cum_log_probs=0, keep_logits=False, callback=None):
        cfg = self.cfg
        t = cfg.tensor_para_size
        L, H, Dh, rot = cfg.layer_num, cfg.head_num, cfg.size_per_head, cfg.rotary_embedding_dim
        Hl = H // t
        hl = Hl * Dh
        input_ids = np.asarray(input_ids, dtype=np.int64)
        lens = np.asarray(input_lengths, dtype=np.int64)
        B, S_in = input_ids.shape
        maxlen = S_in + output_len
        wte = self.ranks[0].w[12 * L].float()
        lnf_g, lnf_b = self.ranks[0].w[12 * L + 1], self.ranks[0].w[12 * L + 2]
        lm_head = self.ranks[0].w[12 * L + 3].float()
        Vp = self.vocab_padded
        inv_sqrt_dh = 1.0 / math.sqrt(Dh)

        kc = [[torch.zeros(B, Hl, maxlen, Dh) for _ in range(L)] for _ in range(t)]
        vc = [[torch.zeros(B, Hl, maxlen, Dh) for _ in range(L)] for _ in range(t)]
        out_ids = np.zeros((maxlen, B), dtype=np.int64)          # time-major, GptNeoX.cc output_ids_buf_
        out_ids[:S_in] = input_ids.T
        seq_len = np.full(B, S_in - 1, dtype=np.int64)           # invokeDecodingInitialize(max_input_length - 1)
        finished = np.zeros(B, dtype=bool)
        cum_log = np.zeros(B, dtype=np.float32)
        masked = np.zeros((B, maxlen), dtype=bool)
        for b in range(B):
            masked[b, lens[b]:S_in] = True                        # gpt_kernels.cu:1036-1050
        pad_count = np.zeros(B, dtype=np.int64)
        logits_trace = []

        ks, ps, _ = S.setup_topk_runtime_args(1 if top_k is None else top_k, 0.0 if top_p is None else top_p, B)
        temp = np.broadcast_to(np.asarray(1.0 if temperature is None else temperature, np.float32).reshape(-1), (B,))
        rep = np.broadcast_to(np.asarray(1.0 if repetition_penalty is None else repetition_penalty,
                                         np.float32).reshape(-1), (B,))
        seeds = np.broadcast_to(np.asarray(0 if random_seed is None else random_seed, np.int64).reshape(-1), (B,))
        rngs = [S.CurandXorwow(int(s)) for s in seeds]
        max_top_k = max(1, int(ks.max()))

        def bias_rotary(r, l, qkv, pos):
            """qkv [m, 3*hl] + bias (fp16), split, rotary at pos [m]."""
            qkv = h(qkv + self._W(r, 3, l).float())
            q, k, v = [z.reshape(-1, Hl, Dh) for z in qkv.split(hl, dim=-1)]
            cos, sin = rotary_coef(pos, rot)
            q = apply_rotary_neox(q, cos[:, None, :], sin[:, None, :], rot)
            k = apply_rotary_neox(k, cos[:, None, :], sin[:, None, :], rot)
            return q, k, v

        x_last = None
        if S_in > 1:
            # ---------------- prefill on compacted tokens (GptNeoXContextDecoder.cc:285-308)
            tok_b = np.concatenate([np.full(lens[b], b) for b in range(B)])
            tok_p = np.concatenate([np.arange(lens[b]) for b in range(B)])
            x = h(wte[torch.from_numpy(input_ids[tok_b, tok_p])])
            offs = np.concatenate([[0], np.cumsum(lens)])

            def ctx_attn(r, l, qkv):
                q, k, v = bias_rotary(r, l, qkv, torch.from_numpy(tok_p))
                ctx = torch.zeros(q.shape[0], Hl, Dh)
                for b in range(B):
                    s, e = offs[b], offs[b + 1]
                    n = e - s
                    kc[r][l][b, :, :n] = k[s:e].transpose(0, 1)
                    vc[r][l][b, :, :n] = v[s:e].transpose(0, 1)
                    qb, kb, vb = q[s:e].transpose(0, 1), k[s:e].transpose(0, 1), v[s:e].transpose(0, 1)
                    sc = (qb @ kb.transpose(1, 2))                               # fp32 QK^T (GptContextAttentionLayer.cc:207-229)
                    mask = torch.tril(torch.ones(n, n)) == 0
                    sc = sc * h(torch.tensor(inv_sqrt_dh)) + mask * (-10000.0)    # unfused_attention_kernels.cu:255-333
                    p = h(torch.softmax(sc, dim=-1))
                    ctx[s:e] = h(p @ vb).transpose(0, 1)
                return ctx.reshape(-1, hl)

            for l in range(L):
                x = self._layer(x, l, ctx_attn)
            x_last = x[torch.from_numpy(offs[1:] - 1)]                            # invokeLookupHiddenStateOfLastToken
        else:
            seq_len[:] = 0

        steps_done = 0
        for step in range(S_in, maxlen):
            if not (S_in > 1 and step == S_in):
                # ---------------- one decode step (GptNeoXDecoder.cc:197-389), timestep = step - 1
                ids_prev = out_ids[step - 1]
                x = h(wte[torch.from_numpy(ids_prev)])
                tl = seq_len.copy()                                               # cache slot / #keys (template.hpp:1204-1209)
                pos = torch.from_numpy((step - 1) - pad_count)

                def dec_attn(r, l, qkv):
                    q, k, v = bias_rotary(r, l, qkv, pos)
                    ctx = torch.zeros(B, Hl, Dh)
                    for b in range(B):
                        if finished[b]:
                            continue                                               # template.hpp:1176-1178
                        tlen = int(tl[b])
                        kc[r][l][b, :, tlen] = k[b]
                        vc[r][l][b, :, tlen] = v[b]
                        keys = kc[r][l][b, :, :tlen + 1]
                        vals = vc[r][l][b, :, :tlen + 1]
                        sc = (keys @ q[b][:, :, None]).squeeze(-1) * inv_sqrt_dh   # [Hl, tlen+1] fp32
                        mk = torch.from_numpy(masked[b, :tlen + 1])
                        sc_m = sc.masked_fill(mk[None, :], float("-inf"))
                        mx = sc_m.max(dim=-1, keepdim=True).values
                        e = torch.exp(sc - mx).masked_fill(mk[None, :], 0.0)
                        p = h(e * (1.0 / (e.sum(-1, keepdim=True) + 1e-6)))        # fp16 probabilities (:1643)
                        ctx[b] = h((p[:, None, :] @ vals).squeeze(1))
                    return ctx.reshape(B, hl)

                for l in range(L):
                    x = self._layer(x, l, dec_attn)
                x_last = x

            # ---------------- LM head + sampling (GptNeoX.cc:854-1022, DynamicDecodeLayer.cc:192-495)
            n = layernorm_ref(x_last, lnf_g, lnf_b, cfg.layernorm_eps)
            logits = torch.zeros(B, Vp)
            logits[:, :cfg.vocab_size] = n @ lm_head.t()
            logits = logits.numpy().astype(np.float32)
            if keep_logits:
                logits_trace.append(logits[:, :cfg.vocab_size].copy())
            if step == S_in and optional_last_tokens is not None:
                S.select_optional_last_tokens(logits, np.asarray(optional_last_tokens))
            if not np.all(temp == 1.0):
                S.apply_temperature(logits, temp, cfg.vocab_size)
            if step > 1 and not np.all(rep == 1.0):
                S.apply_repetition_penalty(logits, rep, out_ids, lens, S_in, step)
            S.add_bias_end_mask(logits, np.full(B, cfg.end_id), finished, cfg.vocab_size)
            is_prob = bool(return_cum_log_probs)
            if is_prob:
                logits = S.softmax_probs(logits)
            for b in range(B):
                if finished[b]:
                    out_ids[step, b] = cfg.end_id
                    continue
                if int(ks[b]) == 0:      # pure top-p row: always on probabilities (TopPSamplingLayer.cu:293-300)
                    pr = logits[b] if is_prob else S.softmax_probs(logits[b:b + 1])[0]
                    tok, val = S.topp_sampling_row(pr, ps[b], rngs[b])
                else:
                    tok, val = S.topk_sampling_row(logits[b], int(ks[b]), ps[b], rngs[b], max_top_k, is_prob)
                out_ids[step, b] = tok
                if is_prob:
                    cum_log[b] += np.float32(np.log(val))
                seq_len[b] += 1
                finished[b] = tok == cfg.end_id
            if stop_words_list is not None:
                S.stop_words_criterion(out_ids, np.asarray(stop_words_list), finished, step)
            finished |= step >= maxlen                                             # length criterion (never true in-loop)
            steps_done += 1
            if callback is not None and step + 1 < maxlen:
                callback(step, out_ids, seq_len.copy())
            if finished.all():
                break
            if step == S_in:
                pad_count = S_in - lens                                            # invokeUpdatePaddingCount

        # ---------------- gather (setOutputTensors GptNeoX.cc:1090-1181 -> gatherTree decoding_kernels.cu:452-580)
        output, out_len = gather_output(out_ids, seq_len, lens, S_in, maxlen, cfg.end_id)
        res = {"output_ids": output, "sequence_lengths": out_len, "cum_log_probs": cum_log.reshape(B, 1),
               "raw_output_ids": out_ids, "steps": steps_done}
        if keep_logits:
            res["logits"] = logits_trace
        return res

    # ------------------------------------------------------------------------------------------- beam search (beam_width > 1)
    def forward_beam(self, input_ids, input_lengths, output_len, beam_width, temperature=None, repetition_penalty=None,
                     beam_search_diversity_rate=None, len_penalty=None, stop_words_list=None, decide_on_logits=None):
        """GptNeoX.cc:385-1052 with beam_width > 1: inputs tiled over the beams (invokeTileGptInputs), prefill and decode on
        batch x beam rows, decode attention through the cache indirection (template.hpp:1494-1522,1709-1761), the online beam
        search of oracle/beam_search_ref.py, gatherTree with parents.  Runtime arguments: element 0 is used for every row
        (DynamicDecodeLayer.cc:308-408 passes the tensors down unsliced).
        decide_on_logits [steps, B*K, vocab] (optional): the beam decisions are taken on THESE logits (e.g. the ones a device run
        traced) while the model's own logits are still computed and returned in "logits" -- a comparison that does not hinge on
        near-ties between candidates (random small models give score gaps of 1e-3, the size of fp16 noise)."""
        from oracle import beam_search_ref as BS
        cfg = self.cfg
        t = cfg.tensor_para_size
        L, H, Dh, rot = cfg.layer_num, cfg.head_num, cfg.size_per_head, cfg.rotary_embedding_dim
        Hl = H // t
        hl = Hl * Dh
        K = int(beam_width)
        input_ids = np.repeat(np.asarray(input_ids, dtype=np.int64), K, axis=0)      # [B*K, S]
        lens = np.repeat(np.asarray(input_lengths, dtype=np.int64), K)
        BB, S_in = input_ids.shape
        B = BB // K
        maxlen = S_in + output_len
        wte = self.ranks[0].w[12 * L].float()
        lnf_g, lnf_b = self.ranks[0].w[12 * L + 1], self.ranks[0].w[12 * L + 2]
        lm_head = self.ranks[0].w[12 * L + 3].float()
        Vp = self.vocab_padded
        inv_sqrt_dh = 1.0 / math.sqrt(Dh)
        first = lambda a, d: float(np.asarray(d if a is None else a, np.float32).reshape(-1)[0])
        temp, rep = first(temperature, 1.0), first(repetition_penalty, 1.0)
        div, lp = first(beam_search_diversity_rate, 0.0), first(len_penalty, 0.0)

        kc = [[torch.zeros(BB, Hl, maxlen, Dh) for _ in range(L)] for _ in range(t)]
        vc = [[torch.zeros(BB, Hl, maxlen, Dh) for _ in range(L)] for _ in range(t)]
        out_ids = np.zeros((maxlen, BB), dtype=np.int64)
        out_ids[:S_in] = input_ids.T
        parent_ids = np.zeros((maxlen, BB), dtype=np.int64)
        seq_len = np.full(BB, S_in - 1, dtype=np.int64)
        finished, cum_log = BS.decoding_initialize(B, K)
        indir = [np.zeros((BB, maxlen), dtype=np.int64), np.zeros((BB, maxlen), dtype=np.int64)]
        masked = np.zeros((BB, maxlen), dtype=bool)
        for b in range(BB):
            masked[b, lens[b]:S_in] = True
        pad_count = np.zeros(BB, dtype=np.int64)
        margins = []                       # per step and batch: the smallest score gap among the K + 1 best candidates
        logits_trace, finished_trace = [], []

        def bias_rotary(r, l, qkv, pos):
            qkv = h(qkv + self._W(r, 3, l).float())
            q, k, v = [z.reshape(-1, Hl, Dh) for z in qkv.split(hl, dim=-1)]
            cos, sin = rotary_coef(pos, rot)
            q = apply_rotary_neox(q, cos[:, None, :], sin[:, None, :], rot)
            k = apply_rotary_neox(k, cos[:, None, :], sin[:, None, :], rot)
            return q, k, v

        assert S_in > 1, "the beam-search restatement covers requests with a prompt (max_input_length > 1)"
        tok_b = np.concatenate([np.full(lens[b], b) for b in range(BB)])
        tok_p = np.concatenate([np.arange(lens[b]) for b in range(BB)])
        x = h(wte[torch.from_numpy(input_ids[tok_b, tok_p])])
        offs = np.concatenate([[0], np.cumsum(lens)])

        def ctx_attn(r, l, qkv):
            q, k, v = bias_rotary(r, l, qkv, torch.from_numpy(tok_p))
            ctx = torch.zeros(q.shape[0], Hl, Dh)
            for b in range(BB):
                s, e = offs[b], offs[b + 1]
                n = e - s
                kc[r][l][b, :, :n] = k[s:e].transpose(0, 1)
                vc[r][l][b, :, :n] = v[s:e].transpose(0, 1)
                qb, kb, vb = q[s:e].transpose(0, 1), k[s:e].transpose(0, 1), v[s:e].transpose(0, 1)
                sc = (qb @ kb.transpose(1, 2))
                mask = torch.tril(torch.ones(n, n)) == 0
                sc = sc * h(torch.tensor(inv_sqrt_dh)) + mask * (-10000.0)
                p = h(torch.softmax(sc, dim=-1))
                ctx[s:e] = h(p @ vb).transpose(0, 1)
            return ctx.reshape(-1, hl)

        for l in range(L):
            x = self._layer(x, l, ctx_attn)
        x_last = x[torch.from_numpy(offs[1:] - 1)]

        steps_done = 0
        for step in range(S_in, maxlen):
            src, tgt = indir[(step - S_in) % 2], indir[1 - (step - S_in) % 2]
            if step != S_in:
                ids_prev = out_ids[step - 1]
                x = h(wte[torch.from_numpy(ids_prev)])
                tl = seq_len.copy()
                pos = torch.from_numpy((step - 1) - pad_count)

                def dec_attn(r, l, qkv):
                    q, k, v = bias_rotary(r, l, qkv, pos)
                    ctx = torch.zeros(BB, Hl, Dh)
                    for bb in range(BB):
                        if finished[bb]:
                            continue
                        tlen = int(tl[bb])
                        kc[r][l][bb, :, tlen] = k[bb]
                        vc[r][l][bb, :, tlen] = v[bb]
                    for bb in range(BB):
                        if finished[bb]:
                            continue
                        tlen = int(tl[bb])
                        rows = torch.from_numpy((bb // K) * K + src[bb, :tlen + 1])
                        rows[tlen] = bb                                              # the new token's slot is the row's own
                        ar = torch.arange(tlen + 1)
                        keys = kc[r][l][rows, :, ar].transpose(0, 1)                 # [Hl, tlen+1, Dh]
                        vals = vc[r][l][rows, :, ar].transpose(0, 1)
                        sc = (keys @ q[bb][:, :, None]).squeeze(-1) * inv_sqrt_dh
                        mk = torch.from_numpy(masked[bb, :tlen + 1])
                        sc_m = sc.masked_fill(mk[None, :], float("-inf"))
                        mx = sc_m.max(dim=-1, keepdim=True).values
                        e = torch.exp(sc - mx).masked_fill(mk[None, :], 0.0)
                        p = h(e * (1.0 / (e.sum(-1, keepdim=True) + 1e-6)))
                        ctx[bb] = h((p[:, None, :] @ vals).squeeze(1))
                    return ctx.reshape(BB, hl)

                for l in range(L):
                    x = self._layer(x, l, dec_attn)
                x_last = x
            finished_trace.append(finished.copy())
            n = layernorm_ref(x_last, lnf_g, lnf_b, cfg.layernorm_eps)
            logits = torch.zeros(BB, Vp)
            logits[:, :cfg.vocab_size] = n @ lm_head.t()
            logits = logits.numpy().astype(np.float32)
            logits_trace.append(logits[:, :cfg.vocab_size].copy())
            if decide_on_logits is not None:
                logits = np.zeros((BB, Vp), dtype=np.float32)
                logits[:, :cfg.vocab_size] = np.asarray(decide_on_logits[steps_done], dtype=np.float32)
            BS.beam_step(logits, step, out_ids, parent_ids, seq_len, finished, cum_log, src, tgt, lens, S_in, K, cfg.vocab_size,
                         cfg.end_id, temp, rep, div, lp, stop_words_list, margins)
            steps_done += 1
            if finished.all():
                break
            if step == S_in:
                pad_count = S_in - lens
        output, out_len = BS.gather_tree(out_ids, parent_ids, seq_len, lens, S_in, maxlen, cfg.end_id, K)
        return {"output_ids": output, "sequence_lengths": out_len, "cum_log_probs": cum_log.reshape(B, K).copy(),
                "raw_output_ids": out_ids, "parent_ids": parent_ids, "steps": steps_done, "min_margin": min(margins),
                "logits": logits_trace, "finished_before": finished_trace}
