"""FT-layout weight containers for the GPT-NeoX / CodeFuse path: synthetic initialisation, the tensor-parallel split
and INT8 preparation.  Load-time tooling, not the hot path.

Layouts and ordering follow the reference:
  * list order `w[field * L + layer]`, 12 per-layer fields then wte, final-LN weight, final-LN bias, lm_head
    (th_op/gptneox/GptNeoXOp.h:121-174; Python side examples/pytorch/codefuse/codefuse_example.py:182-292,347-372);
  * every Linear is stored [in, out]; QKV output columns ordered [3, heads, dh]; tensor-parallel split on the last axis
    for QKV / FFN1 (+ their biases), on axis 0 for O / FFN2; row-parallel biases divided by t and, for the parallel
    residual, summed into one (examples/pytorch/codefuse/huggingface_convert.py:35-82,192-206);
  * `int8_w[kind * L + layer]`, `scale[...]`, kind in {qkv, o, ffn1, ffn2}.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch


@dataclass
class NeoXConfig:
    head_num: int
    size_per_head: int
    inter_size: int
    layer_num: int
    vocab_size: int
    rotary_embedding_dim: int
    start_id: int = 0
    end_id: int = 0
    use_gptj_residual: bool = True

    @property
    def hidden(self) -> int:
        return self.head_num * self.size_per_head


CODEFUSE_13B = NeoXConfig(head_num=40, size_per_head=128, inter_size=20480, layer_num=40, vocab_size=100864,
                          rotary_embedding_dim=128, start_id=100000, end_id=100001)
NEOX_125M = NeoXConfig(head_num=12, size_per_head=64, inter_size=3072, layer_num=12, vocab_size=50304,
                       rotary_embedding_dim=16, start_id=0, end_id=0)

# field indices inside the `w` list
LN1_B, LN1_G, QKV_W, QKV_B, O_W, O_B, FFN1_W, FFN1_B, FFN2_W, FFN2_B, LN2_B, LN2_G = range(12)
KIND_FIELDS = (QKV_W, O_W, FFN1_W, FFN2_W)


def synthetic_layer(cfg: NeoXConfig, layer: int, device, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Unsplit fp16 tensors of one layer, N(0, std) matrices and biases, LN gamma ~ 1, beta ~ 0 (slightly perturbed so
    that a wrong gamma/beta wiring shows up in tests)."""
    g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + layer)
    h, inter = cfg.hidden, cfg.inter_size

    def rn(*shape, s=std):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * s).to(torch.float16).to(device)

    return {
        "ln1_g": (1.0 + rn(h, s=0.05).float()).half(), "ln1_b": rn(h, s=0.05),
        "qkv_w": rn(h, 3 * h), "qkv_b": rn(3 * h),
        "o_w": rn(h, h), "o_b": rn(h),
        "ffn1_w": rn(h, inter), "ffn1_b": rn(inter),
        "ffn2_w": rn(inter, h), "ffn2_b": rn(h),
        "ln2_g": (1.0 + rn(h, s=0.05).float()).half(), "ln2_b": rn(h, s=0.05),
    }


def synthetic_globals(cfg: NeoXConfig, device, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + 999983)
    h, v = cfg.hidden, cfg.vocab_size

    def rn(*shape, s=std):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * s).to(torch.float16).to(device)

    return {"wte": rn(v, h, s=1.0), "lnf_g": (1.0 + rn(h, s=0.05).float()).half(), "lnf_b": rn(h, s=0.05), "lm_head": rn(v, h)}


def split_layer(full: Dict[str, torch.Tensor], cfg: NeoXConfig, t: int, rank: int) -> List[torch.Tensor]:
    """One layer's 12 tensors for tensor-parallel rank `rank` of `t`, in field order."""
    h, hl, il = cfg.hidden, cfg.hidden // t, cfg.inter_size // t
    qkv_w = full["qkv_w"].reshape(h, 3, h)[:, :, rank * hl:(rank + 1) * hl].reshape(h, 3 * hl).contiguous()
    qkv_b = full["qkv_b"].reshape(3, h)[:, rank * hl:(rank + 1) * hl].reshape(3 * hl).contiguous()
    o_w = full["o_w"][rank * hl:(rank + 1) * hl, :].contiguous()
    ffn1_w = full["ffn1_w"][:, rank * il:(rank + 1) * il].contiguous()
    ffn1_b = full["ffn1_b"][rank * il:(rank + 1) * il].contiguous()
    ffn2_w = full["ffn2_w"][rank * il:(rank + 1) * il, :].contiguous()
    o_b = (full["o_b"].float() / t).half()
    ffn2_b = (full["ffn2_b"].float() / t).half()
    if cfg.use_gptj_residual:
        # "mlp.attention.bias.sum": (o_b + ffn2_b) / t, huggingface_convert.py:192-206
        ffn2_b = ((full["o_b"].float() + full["ffn2_b"].float()) / t).half()
    return [full["ln1_b"], full["ln1_g"], qkv_w, qkv_b, o_w, o_b, ffn1_w, ffn1_b, ffn2_w, ffn2_b, full["ln2_b"], full["ln2_g"]]


def quantize_on_device(w_kn: torch.Tensor):
    """Weight-only INT8 of one [k, n] matrix with torch ops on the tensor's device -- the same arithmetic as
    ftcf_symmetric_quantize_int8_host (cutlass_preprocessors.cc:603-640): scale = absmax/128 in fp32, stored in the
    weight dtype; q = clip(round_half_away(w / scale_fp32), -128, 127); processed = (q + 128)^T as uint8 bytes.
    Returns (processed int8-typed [k, n]-shaped tensor holding the B200-layout bytes, scale [n], plain q [k, n])."""
    w32 = w_kn.float()
    scale = w32.abs().amax(dim=0) * (1.0 / 128.0)
    s = w32 / scale[None, :]
    r = torch.sign(s) * torch.floor(s.abs() + 0.5)
    r = torch.where(torch.isnan(r), torch.full_like(r, 127.0), r)
    q = r.clamp_(-128.0, 127.0).to(torch.int8)
    processed = (q.t().contiguous().to(torch.int16) + 128).to(torch.uint8).view(torch.int8).reshape(w_kn.shape)
    return processed, scale.to(w_kn.dtype), q


class RankWeights:
    """What one process hands to GptNeoXOp: `w` (12L+4 fp16), `int8_w`, `scale` (4L each, empty when int8_mode == 0)."""

    def __init__(self, cfg: NeoXConfig, t: int, rank: int, int8_mode: int):
        self.cfg, self.t, self.rank, self.int8_mode = cfg, t, rank, int8_mode
        L = cfg.layer_num
        self.w: List[Optional[torch.Tensor]] = [None] * (12 * L + 4)
        self.int8_w: List[Optional[torch.Tensor]] = [None] * (4 * L)
        self.scale: List[Optional[torch.Tensor]] = [None] * (4 * L)
        self.plain_q: List[Optional[torch.Tensor]] = [None] * (4 * L)   # kept only when asked (tests / oracle)

    def set_layer(self, layer: int, tensors: List[torch.Tensor], keep_plain: bool = False, keep_fp16: bool = True) -> None:
        L = self.cfg.layer_num
        for f, tns in enumerate(tensors):
            self.w[f * L + layer] = tns
        if self.int8_mode == 1:
            for kind, f in enumerate(KIND_FIELDS):
                p, s, q = quantize_on_device(tensors[f])
                self.int8_w[kind * L + layer] = p
                self.scale[kind * L + layer] = s
                if keep_plain:
                    self.plain_q[kind * L + layer] = q
                if not keep_fp16:
                    # the driver keeps the fp16 matrices as empty placeholders once quantised (codefuse_example.py:400-406)
                    self.w[f * L + layer] = torch.empty(0, dtype=torch.float16, device=tensors[f].device)

    def set_globals(self, g: Dict[str, torch.Tensor]) -> None:
        L = self.cfg.layer_num
        self.w[12 * L + 0] = g["wte"]
        self.w[12 * L + 1] = g["lnf_g"]      # weight, then bias (GptNeoXOp.h:172-173)
        self.w[12 * L + 2] = g["lnf_b"]
        self.w[12 * L + 3] = g["lm_head"]

    def lists(self):
        dev = self.w[-1].device
        empty8 = torch.empty(0, dtype=torch.int8, device=dev)
        empty16 = torch.empty(0, dtype=torch.float16, device=dev)
        return (list(self.w), [x if x is not None else empty8 for x in self.int8_w],
                [x if x is not None else empty16 for x in self.scale])


def make_synthetic(cfg: NeoXConfig, t: int, rank: int, int8_mode: int, device, seed: int = 0, keep_plain: bool = False,
                   keep_fp16: bool = True) -> RankWeights:
    """Synthetic weights of rank `rank`: every rank derives them from the same seeded full tensors, layer by layer,
    so memory stays bounded by one unsplit layer."""
    rw = RankWeights(cfg, t, rank, int8_mode)
    for layer in range(cfg.layer_num):
        full = synthetic_layer(cfg, layer, device, seed)
        rw.set_layer(layer, split_layer(full, cfg, t, rank), keep_plain=keep_plain, keep_fp16=keep_fp16)
    rw.set_globals(synthetic_globals(cfg, device, seed))
    return rw


def make_synthetic_fast(cfg: NeoXConfig, t: int, rank: int, int8_mode: int, device, seed: int = 0) -> RankWeights:
    """Benchmark-sized synthetic weights generated ON the device (torch.randn there), one layer at a time; values are
    N(0, 0.02) like make_synthetic but not bit-identical to it.  All ranks use the same per-layer seed so that the
    shards are slices of one model."""
    rw = RankWeights(cfg, t, rank, int8_mode)
    h, inter, v = cfg.hidden, cfg.inter_size, cfg.vocab_size
    gen = torch.Generator(device=device)

    def rn(*shape, s=0.02):
        return (torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * s).to(torch.float16)

    for layer in range(cfg.layer_num):
        gen.manual_seed(seed * 1000003 + layer)
        full = {"ln1_g": (1.0 + rn(h, s=0.05).float()).half(), "ln1_b": rn(h, s=0.05), "qkv_w": rn(h, 3 * h), "qkv_b": rn(3 * h),
                "o_w": rn(h, h), "o_b": rn(h), "ffn1_w": rn(h, inter), "ffn1_b": rn(inter), "ffn2_w": rn(inter, h),
                "ffn2_b": rn(h), "ln2_g": (1.0 + rn(h, s=0.05).float()).half(), "ln2_b": rn(h, s=0.05)}
        rw.set_layer(layer, split_layer(full, cfg, t, rank), keep_fp16=(int8_mode == 0))
        del full
    gen.manual_seed(seed * 1000003 + 999983)
    rw.set_globals({"wte": rn(v, h, s=1.0), "lnf_g": (1.0 + rn(h, s=0.05).float()).half(), "lnf_b": rn(h, s=0.05),
                    "lm_head": rn(v, h)})
    return rw
