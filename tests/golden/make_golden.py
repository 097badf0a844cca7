#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ (run in the authoring container; needs /root/reference
only for `quant_ref_*.npz`).  TEST INFRASTRUCTURE.

  neox_hf_*.npz    model-wiring goldens: a tiny random-init HuggingFace `GPTNeoXForCausalLM` (fp32 math on fp16-
                   representable weights, tanh GELU as FasterTransformer uses, parallel / sequential residual), its
                   weights converted to the FasterTransformer layout with the rules of the reference converter
                   (examples/pytorch/codefuse/huggingface_convert.py:22-82,151-206: [in,out] transposes, QKV columns
                   [H,3,Dh] -> [3,H,Dh]), the prompt, greedy token ids and the logits of every generated step.
  quant_ref_*.npz  quantiser goldens: outputs of the REFERENCE's own object code (oracle/_ref/libref_quant.so, built
                   from /root/reference/src/fastertransformer/kernels/cutlass_kernels/cutlass_preprocessors.cc) --
                   unprocessed int8, sm80-processed bytes and fp16 scales for seeded fp16 / fp32 matrices.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def hf_to_ft_lists(model, cfg, use_parallel_residual):
    """HF state dict -> GptNeoXOp weight list (t = 1), index = field * L + layer (th_op/gptneox/GptNeoXOp.h:121-174)."""
    L, H = cfg.num_hidden_layers, cfg.num_attention_heads
    h = cfg.hidden_size
    dh = h // H
    sd = {k: v.detach().float().numpy() for k, v in model.state_dict().items()}
    w = [None] * (12 * L + 4)
    for l in range(L):
        p = f"gpt_neox.layers.{l}."
        qkv_w = sd[p + "attention.query_key_value.weight"].T                      # [in, out]
        qkv_w = qkv_w.reshape(h, H, 3, dh).transpose(0, 2, 1, 3).reshape(h, 3 * h)
        qkv_b = sd[p + "attention.query_key_value.bias"].reshape(H, 3, dh).transpose(1, 0, 2).reshape(3 * h)
        o_b, f2_b = sd[p + "attention.dense.bias"], sd[p + "mlp.dense_4h_to_h.bias"]
        fields = [sd[p + "input_layernorm.bias"], sd[p + "input_layernorm.weight"], qkv_w, qkv_b,
                  sd[p + "attention.dense.weight"].T, o_b, sd[p + "mlp.dense_h_to_4h.weight"].T,
                  sd[p + "mlp.dense_h_to_4h.bias"], sd[p + "mlp.dense_4h_to_h.weight"].T,
                  (o_b + f2_b) if use_parallel_residual else f2_b,                 # "mlp.attention.bias.sum", :192-206
                  sd[p + "post_attention_layernorm.bias"], sd[p + "post_attention_layernorm.weight"]]
        for f, a in enumerate(fields):
            w[f * L + l] = np.ascontiguousarray(a).astype(np.float16)
    w[12 * L + 0] = sd["gpt_neox.embed_in.weight"].astype(np.float16)
    w[12 * L + 1] = sd["gpt_neox.final_layer_norm.weight"].astype(np.float16)
    w[12 * L + 2] = sd["gpt_neox.final_layer_norm.bias"].astype(np.float16)
    w[12 * L + 3] = sd["embed_out.weight"].astype(np.float16)
    return w


def make_hf(name, seed, parallel, heads=4, dh=16, layers=2, inter=256, vocab=128, rot_pct=0.5, B=2, S=9, out=6, lens=(9, 5)):
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM
    torch.manual_seed(seed)
    cfg = GPTNeoXConfig(hidden_size=heads * dh, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=inter,
                        vocab_size=vocab, hidden_act="gelu_new", use_parallel_residual=parallel, max_position_embeddings=64,
                        tie_word_embeddings=False, layer_norm_eps=1e-5, attention_dropout=0.0, hidden_dropout=0.0,
                        rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "partial_rotary_factor": rot_pct})
    model = GPTNeoXForCausalLM(cfg).eval()
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if "layernorm" in n_ or "layer_norm" in n_:
                p.add_(torch.randn_like(p) * 0.05)
            elif p.dim() == 1:
                p.normal_(0.0, 0.05)
            else:
                p.normal_(0.0, 0.08)
            p.copy_(p.half().float())                                              # fp16-representable
    g = np.random.default_rng(seed)
    ids = g.integers(0, vocab - 1, size=(B, S)).astype(np.int32)
    end_id = vocab - 1
    rows, logits = [], []
    with torch.no_grad():
        for b in range(B):                                                         # one sequence at a time: no padding
            cur = torch.from_numpy(ids[b:b + 1, :lens[b]].astype(np.int64))
            lg = []
            for _ in range(out):
                o = model(cur).logits[0, -1]
                lg.append(o.numpy().copy())
                cur = torch.cat([cur, o.argmax().reshape(1, 1)], dim=1)
            rows.append(cur[0, lens[b]:].numpy().astype(np.int32))
            logits.append(np.stack(lg))
    w = hf_to_ft_lists(model, cfg, parallel)
    arrays = {f"w{i}": a for i, a in enumerate(w)}
    for b in range(B):
        ids[b, lens[b]:] = end_id                                                   # right padding with end_id, codefuse_example.py:700
    np.savez_compressed(os.path.join(HERE, name), ids=ids, lens=np.asarray(lens, np.int32), out_len=np.int32(out),
                        gen=np.stack(rows), logits=np.stack(logits).astype(np.float32),
                        meta=np.asarray([heads, dh, inter, layers, vocab, int(dh * rot_pct), end_id, int(parallel)], np.int32), **arrays)
    print("wrote", name)


def make_quant(name, seed, k, n, dtype):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_quant.so"))
    g = np.random.default_rng(seed)
    w = (g.standard_normal((k, n)) * 0.02).astype(dtype)
    w[3, 5] = 0.0
    proc = np.empty((k, n), np.int8)
    unproc = np.empty((k, n), np.int8)
    scales = np.empty(n, np.float16)
    shape = (C.c_size_t * 2)(k, n)
    fn = lib.ref_symmetric_quantize_half if dtype == np.float16 else lib.ref_symmetric_quantize_float
    rc = fn(proc.ctypes.data_as(C.c_void_p), unproc.ctypes.data_as(C.c_void_p), scales.ctypes.data_as(C.c_void_p),
            w.ctypes.data_as(C.c_void_p), shape, 2)
    assert rc == 0
    np.savez_compressed(os.path.join(HERE, name), w=w, processed=proc, unprocessed=unproc, scales=scales)
    print("wrote", name)


if __name__ == "__main__":
    make_hf("neox_hf_parallel.npz", 11, True)
    make_hf("neox_hf_sequential.npz", 12, False)
    make_hf("neox_hf_fullrot.npz", 13, True, heads=2, dh=32, rot_pct=1.0, B=1, S=7, out=5, lens=(7,))
    make_quant("quant_ref_f16_128x64.npz", 21, 128, 64, np.float16)
    make_quant("quant_ref_f32_64x128.npz", 22, 64, 128, np.float32)
