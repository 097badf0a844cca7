"""On-disk checkpoint tooling of the CodeFuse / GPT-NeoX path (load-time, CPU): SURVEY.md section 8(f) rank 2.

Three steps, each writing / reading exactly the files the reference's scripts do, so that directories are interchangeable
with `examples/pytorch/codefuse/` of the reference:

  convert_hf(model_or_dir, out_dir, t)     HuggingFace GPTNeoXForCausalLM -> `<out_dir>/<t>-gpu/model.*.bin` + `config.ini`
                                           (what huggingface_convert.py:22-82,84-206 produces: every Linear transposed to
                                           [in, out]; QKV output features re-ordered [H,3,Dh] -> [3,H,Dh]; column split of
                                           QKV / FFN1 (+ biases), row split of O / FFN2; row-parallel biases divided by t;
                                           `mlp.attention.bias.sum` for the parallel residual)
  quantize_dir(in_dir, out_dir, t)         adds `*.q.bin` / `*.s.bin` next to the fp files (quant_and_save.py:30-101), with
                                           OUR quantiser: same scales and rounding as the reference, int8 bytes in the B200
                                           layout (W^T, k contiguous, q + 128)
  load_rank(ckpt_dir, rank, ...)           one rank's `(weights, int8_weights, scale)` lists in GptNeoXOp order, as
                                           GptNeoXWeights.load does (codefuse_example.py:340-419)

    python -m fastertransformer4codefuse_b200.checkpoint convert  -i <hf dir> -o <out dir> -t 2 [--dtype fp16]
    python -m fastertransformer4codefuse_b200.checkpoint quantize -i <out dir>/2-gpu -o <out dir>/2-gpu-int8 -t 2
"""
from __future__ import annotations

import argparse
import configparser
import os
import shutil
from typing import List, Tuple

import numpy as np
import torch

from . import quant
from .weights import NeoXConfig

_NP = {"fp16": np.float16, "fp32": np.float32, "float16": np.float16, "float32": np.float32}
# file stems in GptNeoXOp field order (codefuse_example.py:347-358); %d = tensor-parallel rank
_SPLIT_W = ("attention.query_key_value.weight", "attention.dense.weight", "mlp.dense_h_to_4h.weight", "mlp.dense_4h_to_h.weight")


def _rotary_dim(hf_cfg, head_size: int) -> int:
    """`rotary_pct` moved into `rope_parameters["partial_rotary_factor"]` in transformers 5."""
    pct = getattr(hf_cfg, "rotary_pct", None)
    if pct is None:
        rp = getattr(hf_cfg, "rope_parameters", None) or {}
        pct = rp.get("partial_rotary_factor", 1.0)
    return int(head_size * pct)


def convert_hf(model, out_dir: str, tensor_para_size: int, weight_data_type: str = "fp16", model_name: str = "codefuse") -> str:
    """`model`: a GPTNeoXForCausalLM or a directory for `from_pretrained`.  Returns the `<t>-gpu` directory written."""
    if isinstance(model, str):
        from transformers import GPTNeoXForCausalLM
        model = GPTNeoXForCausalLM.from_pretrained(model)
    hf = model.config
    t = int(tensor_para_size)
    dt = _NP[weight_data_type]
    H, h, L = hf.num_attention_heads, hf.hidden_size, hf.num_hidden_layers
    dh = h // H
    if H % t or hf.intermediate_size % t:
        raise ValueError(f"head_num {H} / inter_size {hf.intermediate_size} not divisible by tensor_para_size {t}")
    parallel = bool(hf.use_parallel_residual)
    saved = os.path.join(out_dir, f"{t}-gpu")
    os.makedirs(saved, exist_ok=True)

    ini = configparser.ConfigParser()
    ini["gptneox"] = {
        "model_name": model_name, "head_num": str(H), "size_per_head": str(dh), "inter_size": str(hf.intermediate_size),
        "num_layer": str(L), "rotary_embedding": str(_rotary_dim(hf, dh)), "vocab_size": str(hf.vocab_size),
        "start_id": str(hf.bos_token_id), "end_id": str(hf.eos_token_id), "use_gptj_residual": str(int(parallel)),
        "weight_data_type": weight_data_type,
    }
    with open(os.path.join(saved, "config.ini"), "w") as f:
        ini.write(f)

    sd = {k: v.detach().to(torch.float32).cpu().numpy() for k, v in model.state_dict().items()}

    def put(name: str, a: np.ndarray) -> None:
        np.ascontiguousarray(a).astype(dt).tofile(os.path.join(saved, f"model.{name}.bin"))

    put("wte", sd["gpt_neox.embed_in.weight"])
    put("final_layernorm.weight", sd["gpt_neox.final_layer_norm.weight"])
    put("final_layernorm.bias", sd["gpt_neox.final_layer_norm.bias"])
    put("lm_head.weight", sd["embed_out.weight"])
    for l in range(L):
        p = f"gpt_neox.layers.{l}."
        o = f"layers.{l}."
        for ln in ("input_layernorm", "post_attention_layernorm"):
            put(o + ln + ".weight", sd[p + ln + ".weight"])
            put(o + ln + ".bias", sd[p + ln + ".bias"])
        # QKV: [in, out] with the output features [H, 3, Dh] -> [3, H, Dh]; rank r takes its heads of each third
        qkv_w = sd[p + "attention.query_key_value.weight"].T.reshape(h, H, 3, dh).transpose(0, 2, 1, 3).reshape(h, 3, h)
        qkv_b = sd[p + "attention.query_key_value.bias"].reshape(H, 3, dh).transpose(1, 0, 2).reshape(3, h)
        o_w = sd[p + "attention.dense.weight"].T            # [h, h]: rows split
        f1_w = sd[p + "mlp.dense_h_to_4h.weight"].T         # [h, inter]: columns split
        f1_b = sd[p + "mlp.dense_h_to_4h.bias"]
        f2_w = sd[p + "mlp.dense_4h_to_h.weight"].T         # [inter, h]: rows split
        for r, (a, b, c, d, e, f) in enumerate(zip(np.split(qkv_w, t, axis=-1), np.split(qkv_b, t, axis=-1), np.split(o_w, t, axis=0),
                                                   np.split(f1_w, t, axis=-1), np.split(f1_b, t, axis=-1), np.split(f2_w, t, axis=0))):
            put(o + f"attention.query_key_value.weight.{r}", a)
            put(o + f"attention.query_key_value.bias.{r}", b)
            put(o + f"attention.dense.weight.{r}", c)
            put(o + f"mlp.dense_h_to_4h.weight.{r}", d)
            put(o + f"mlp.dense_h_to_4h.bias.{r}", e)
            put(o + f"mlp.dense_4h_to_h.weight.{r}", f)
        # row-parallel biases: every rank adds bias / t, the all-reduce restores it (huggingface_convert.py:35-41); the
        # reference divides in the OUTPUT dtype and, for the parallel residual, sums the two stored files (:192-206)
        ob = (sd[p + "attention.dense.bias"].astype(dt) / t) if t > 1 else sd[p + "attention.dense.bias"].astype(dt)
        fb = (sd[p + "mlp.dense_4h_to_h.bias"].astype(dt) / t) if t > 1 else sd[p + "mlp.dense_4h_to_h.bias"].astype(dt)
        put(o + "attention.dense.bias", ob)
        put(o + "mlp.dense_4h_to_h.bias", fb)
        if parallel:
            put(o + "mlp.attention.bias.sum", (ob.astype(dt) + fb.astype(dt)).astype(dt))
    return saved


def read_config(ckpt_dir: str) -> Tuple[NeoXConfig, np.dtype]:
    ini = configparser.ConfigParser()
    if not ini.read(os.path.join(ckpt_dir, "config.ini")):
        raise FileNotFoundError(os.path.join(ckpt_dir, "config.ini"))
    g = ini["gptneox"]
    cfg = NeoXConfig(head_num=int(g["head_num"]), size_per_head=int(g["size_per_head"]), inter_size=int(g["inter_size"]),
                     layer_num=int(g["num_layer"]), vocab_size=int(g["vocab_size"]), rotary_embedding_dim=int(g["rotary_embedding"]),
                     start_id=int(g["start_id"]), end_id=int(g["end_id"]), use_gptj_residual=g["use_gptj_residual"] == "1")
    return cfg, _NP[g["weight_data_type"]]


def _shapes(cfg: NeoXConfig, t: int):
    h, hl, il = cfg.hidden, cfg.hidden // t, cfg.inter_size // t
    return {"attention.query_key_value.weight": (h, 3 * hl), "attention.dense.weight": (hl, h),
            "mlp.dense_h_to_4h.weight": (h, il), "mlp.dense_4h_to_h.weight": (il, h)}


def quantize_dir(in_dir: str, out_dir: str, tensor_para_size: int) -> None:
    """quant_and_save.py: copy the directory, add `model.layers.<l>.<name>.<rank>.q.bin` (int8, processed layout) and
    `.s.bin` (per-column scales in the weight dtype) for the four GEMM weights of every layer and rank."""
    cfg, dt = read_config(in_dir)
    if os.path.abspath(in_dir) != os.path.abspath(out_dir):
        if os.path.exists(out_dir):
            shutil.rmtree(out_dir)
        shutil.copytree(in_dir, out_dir)
    for rank in range(tensor_para_size):
        for name, shape in _shapes(cfg, tensor_para_size).items():
            for l in range(cfg.layer_num):
                stem = os.path.join(out_dir, f"model.layers.{l}.{name}.{rank}")
                w = torch.from_numpy(np.fromfile(stem + ".bin", dtype=dt).reshape(shape))
                q, s = quant.symmetric_quantize_last_axis_of_batched_matrix_int8(w)
                q.numpy().tofile(stem + ".q.bin")
                s.numpy().tofile(stem + ".s.bin")
    # the *.q.bin bytes are in the B200 layout (W^T, k contiguous, q + 128), NOT the reference's sm80 interleaved layout: say so in
    # config.ini (an extra key the reference driver ignores), so that a loader given the wrong FTCF_INT8_LAYOUT fails loudly
    ini = configparser.ConfigParser()
    ini.read(os.path.join(out_dir, "config.ini"))
    ini["gptneox"][INT8_LAYOUT_KEY] = "b200"
    with open(os.path.join(out_dir, "config.ini"), "w") as f:
        ini.write(f)


INT8_LAYOUT_KEY = "int8_weight_layout"       # "b200" (ours, FTCF_INT8_LAYOUT=0); absent: made by the reference's quant_and_save.py (=2)


def int8_layout_of(ckpt_dir: str) -> int:
    """ftcf_gptneox_config.int8_layout the pre-quantised files of this directory need (0: ours, 2: the reference's sm80 layout)."""
    ini = configparser.ConfigParser()
    ini.read(os.path.join(ckpt_dir, "config.ini"))
    return 0 if ini["gptneox"].get(INT8_LAYOUT_KEY, "") == "b200" else 2


def load_rank(ckpt_dir: str, tensor_para_rank: int, tensor_para_size: int, int8_mode: int = 0, enable_int8_weights: bool = False,
              device="cpu") -> Tuple[NeoXConfig, List[torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
    """(cfg, weights, int8_weights, scale) of one rank in GptNeoXOp order (th_op/gptneox/GptNeoXOp.h:121-174): fp16 tensors,
    empty tensors where a field is absent.  int8_mode = 1 with enable_int8_weights reads `*.q.bin` / `*.s.bin`; without it
    the fp weights are quantised here, as the reference driver does at load (codefuse_example.py:388-403)."""
    cfg, dt = read_config(ckpt_dir)
    L, t, r = cfg.layer_num, tensor_para_size, tensor_para_rank
    h, hl, il = cfg.hidden, cfg.hidden // t, cfg.inter_size // t
    shapes = _shapes(cfg, t)
    empty = torch.empty(0, dtype=torch.float16)

    def rd(stem: str, shape) -> torch.Tensor:
        return torch.from_numpy(np.fromfile(os.path.join(ckpt_dir, f"model.{stem}.bin"), dtype=dt).reshape(shape)).to(torch.float16)

    preq = int8_mode == 1 and enable_int8_weights
    if preq:
        want = int(os.environ.get("FTCF_INT8_LAYOUT", "0"))
        have = int8_layout_of(ckpt_dir)
        if want != have:
            raise ValueError(f"{ckpt_dir}: the *.q.bin files are in int8 layout {have} ({'B200, written by checkpoint.quantize_dir' if have == 0 else 'sm80 interleaved, written by the reference quant_and_save.py'})"
                             f" but FTCF_INT8_LAYOUT selects {want}; set FTCF_INT8_LAYOUT={have}")
    fields = [("input_layernorm.bias", (h,)), ("input_layernorm.weight", (h,)),
              (None if preq else f"attention.query_key_value.weight.{r}", shapes["attention.query_key_value.weight"]),
              (f"attention.query_key_value.bias.{r}", (3 * hl,)),
              (None if preq else f"attention.dense.weight.{r}", shapes["attention.dense.weight"]),
              (None if cfg.use_gptj_residual else "attention.dense.bias", (h,)),
              (None if preq else f"mlp.dense_h_to_4h.weight.{r}", shapes["mlp.dense_h_to_4h.weight"]),
              (f"mlp.dense_h_to_4h.bias.{r}", (il,)),
              (None if preq else f"mlp.dense_4h_to_h.weight.{r}", shapes["mlp.dense_4h_to_h.weight"]),
              ("mlp.attention.bias.sum" if cfg.use_gptj_residual else "mlp.dense_4h_to_h.bias", (h,)),
              ("post_attention_layernorm.bias", (h,)), ("post_attention_layernorm.weight", (h,))]
    w: List[torch.Tensor] = []
    for stem, shape in fields:
        for l in range(L):
            w.append(empty if stem is None else rd(f"layers.{l}.{stem}", shape))
    w.append(rd("wte", (cfg.vocab_size, h)))
    w.append(rd("final_layernorm.weight", (h,)))
    w.append(rd("final_layernorm.bias", (h,)))
    w.append(rd("lm_head.weight", (cfg.vocab_size, h)))

    int8_w: List[torch.Tensor] = []
    scale: List[torch.Tensor] = []
    if int8_mode == 1:
        for kind, name in enumerate(_SPLIT_W):
            for l in range(L):
                if preq:
                    stem = os.path.join(ckpt_dir, f"model.layers.{l}.{name}.{r}")
                    int8_w.append(torch.from_numpy(np.fromfile(stem + ".q.bin", dtype=np.int8)))     # loaded flat, as the reference
                    scale.append(torch.from_numpy(np.fromfile(stem + ".s.bin", dtype=dt)).to(torch.float16))
                else:
                    idx = (2, 4, 6, 8)[kind] * L + l
                    q, s = quant.symmetric_quantize_last_axis_of_batched_matrix_int8(w[idx])
                    int8_w.append(q)
                    scale.append(s)
                    w[idx] = empty
    dev = torch.device(device)
    return cfg, [x.to(dev) for x in w], [x.to(dev) for x in int8_w], [x.to(dev) for x in scale]


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("convert")
    c.add_argument("-i", "--in_file", required=True)
    c.add_argument("-o", "--saved_dir", required=True)
    c.add_argument("-t", "--infer_gpu_num", type=int, required=True)
    c.add_argument("--dtype", default="fp16", choices=["fp16", "fp32"])
    c.add_argument("--model_name", default="codefuse")
    q = sub.add_parser("quantize")
    q.add_argument("-i", "--in_dir", required=True)
    q.add_argument("-o", "--out_dir", required=True)
    q.add_argument("-t", "--tensor_para_size", type=int, required=True)
    a = ap.parse_args()
    if a.cmd == "convert":
        print(convert_hf(a.in_file, a.saved_dir, a.infer_gpu_num, a.dtype, a.model_name))
    else:
        quantize_dir(a.in_dir, a.out_dir, a.tensor_para_size)


if __name__ == "__main__":
    main()
