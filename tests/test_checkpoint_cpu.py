"""CPU tests of the checkpoint tooling: HuggingFace GPT-NeoX -> FT files -> per-rank weight
lists.  The converted shards, fed to the oracle (tensor parallelism emulated), must reproduce HuggingFace's own logits
and greedy tokens; file names, config.ini keys and the *.q.bin / *.s.bin pairs are the ones the reference's driver reads
(examples/pytorch/codefuse/codefuse_example.py:340-419, huggingface_convert.py, quant_and_save.py)."""
import os

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import checkpoint as CK
from fastertransformer4codefuse_b200 import quant
from oracle import gptneox_ref as R


def _tiny_hf(parallel, seed=3):
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM
    torch.manual_seed(seed)
    cfg = GPTNeoXConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256, vocab_size=96,
                        hidden_act="gelu_new", use_parallel_residual=parallel, max_position_embeddings=64, tie_word_embeddings=False,
                        layer_norm_eps=1e-5, attention_dropout=0.0, hidden_dropout=0.0, bos_token_id=0, eos_token_id=95,
                        rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "partial_rotary_factor": 0.5})
    model = GPTNeoXForCausalLM(cfg).eval()
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if "layernorm" in n_ or "layer_norm" in n_:
                p.add_(torch.randn_like(p) * 0.05)
            elif p.dim() == 1:
                p.normal_(0.0, 0.05)
            else:
                p.normal_(0.0, 0.08)
            p.copy_(p.half().float())
    return model


def _hf_greedy(model, ids, out):
    cur = torch.from_numpy(ids.astype(np.int64))[None]
    logits = []
    with torch.no_grad():
        for _ in range(out):
            o = model(cur).logits[0, -1]
            logits.append(o.numpy().copy())
            cur = torch.cat([cur, o.argmax().reshape(1, 1)], dim=1)
    return cur[0, len(ids):].numpy(), np.stack(logits)


@pytest.mark.parametrize("parallel", [True, False])
@pytest.mark.parametrize("t", [1, 2])
def test_convert_load_matches_huggingface(tmp_path, parallel, t):
    model = _tiny_hf(parallel)
    if t > 1 and not parallel:
        # Reference quirk kept for file compatibility: the converter stores the row-parallel biases divided by t
        # (huggingface_convert.py:35-41), but the sequential-residual decoder adds them once AFTER the all-reduce
        # (GptNeoXDecoder.cc:313-331,361-368), so t > 1 with use_gptj_residual = 0 ends up with bias / t.  CodeFuse uses
        # the parallel residual, where the sum restores the bias.  Zero those two biases here so that the rest of the
        # wiring can still be compared with HuggingFace.
        with torch.no_grad():
            for layer in model.gpt_neox.layers:
                layer.attention.dense.bias.zero_()
                layer.mlp.dense_4h_to_h.bias.zero_()
    saved = CK.convert_hf(model, str(tmp_path), t, "fp16")
    assert os.path.basename(saved) == f"{t}-gpu"
    cfg, _ = CK.read_config(saved)
    assert (cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size, cfg.rotary_embedding_dim) == (4, 16, 256, 2, 96, 8)
    assert cfg.use_gptj_residual == parallel and cfg.end_id == 95
    for name in ("model.wte.bin", "model.lm_head.weight.bin", "model.final_layernorm.weight.bin",
                 f"model.layers.1.attention.query_key_value.weight.{t - 1}.bin", "model.layers.0.mlp.dense_4h_to_h.bias.bin"):
        assert os.path.exists(os.path.join(saved, name)), name
    assert os.path.exists(os.path.join(saved, "model.layers.0.mlp.attention.bias.sum.bin")) == parallel
    ranks = []
    for r in range(t):
        c2, w, q, s = CK.load_rank(saved, r, t)
        assert len(w) == 12 * cfg.layer_num + 4 and not q and not s
        assert tuple(w[2 * cfg.layer_num].shape) == (64, 3 * 64 // t) and tuple(w[8 * cfg.layer_num].shape) == (256 // t, 64)
        ranks.append(R.RankWeights(w=w))
    rcfg = R.RefConfig(head_num=4, size_per_head=16, inter_size=256, layer_num=2, vocab_size=96, rotary_embedding_dim=8, start_id=0,
                       end_id=95, tensor_para_size=t, int8_mode=0, use_gptj_residual=parallel)
    ref = R.GptNeoXRef(rcfg, ranks)
    ids = np.random.default_rng(1).integers(0, 94, size=(1, 7)).astype(np.int32)
    gen, logits = _hf_greedy(model, ids[0], 5)
    res = ref.forward(ids, [7], 5, keep_logits=True)
    for s_ in range(5):
        np.testing.assert_allclose(res["logits"][s_][0], logits[s_], rtol=0, atol=5e-2)
        top2 = np.sort(logits[s_])[-2:]
        if top2[1] - top2[0] > 0.1:
            assert res["output_ids"][0, 0, 7 + s_] == gen[s_]
        elif res["output_ids"][0, 0, 7 + s_] != gen[s_]:
            break


def test_quantize_dir_writes_what_load_time_quantisation_gives(tmp_path):
    model = _tiny_hf(True, seed=5)
    saved = CK.convert_hf(model, str(tmp_path), 2, "fp16")
    out = str(tmp_path / "int8")
    CK.quantize_dir(saved, out, 2)
    for r in range(2):
        cfg, w, q, s = CK.load_rank(out, r, 2, int8_mode=1, enable_int8_weights=True)
        cfg2, w2, q2, s2 = CK.load_rank(saved, r, 2, int8_mode=1, enable_int8_weights=False)
        assert len(q) == len(s) == 4 * cfg.layer_num == len(q2)
        for a, b in zip(q, q2):
            assert torch.equal(a.reshape(-1), b.reshape(-1))
        for a, b in zip(s, s2):
            assert torch.equal(a, b)
        # with pre-quantised files the fp GEMM weights are absent (empty tensors), the rest is there
        L = cfg.layer_num
        assert all(w[f * L + l].numel() == 0 for f in (2, 4, 6, 8) for l in range(L))
        assert w[3 * L].numel() > 0 and w[12 * L].numel() > 0
        # the scale is absmax / 128 of the fp column, in fp16
        wf = torch.from_numpy(np.fromfile(os.path.join(saved, f"model.layers.0.attention.dense.weight.{r}.bin"), dtype=np.float16))
        wf = wf.reshape(64 // 2, 64).float()
        exp = (wf.abs().amax(dim=0) / 128.0).half()
        assert torch.equal(s[1 * L + 0], exp)


def test_int8_layout_marker_guards_against_the_wrong_loader_setting(tmp_path, monkeypatch):
    """quantize_dir writes the B200 byte layout and says so in config.ini; a directory without the marker counts as made by the
    reference's quant_and_save.py (sm80 layout).  Loading with the other FTCF_INT8_LAYOUT fails instead of producing garbage."""
    model = _tiny_hf(True, seed=6)
    saved = CK.convert_hf(model, str(tmp_path), 1, "fp16")
    out = str(tmp_path / "int8")
    CK.quantize_dir(saved, out, 1)
    assert CK.int8_layout_of(out) == 0 and CK.int8_layout_of(saved) == 2
    CK.read_config(out)                                               # the extra key does not disturb the config reader
    monkeypatch.setenv("FTCF_INT8_LAYOUT", "2")
    with pytest.raises(ValueError, match="FTCF_INT8_LAYOUT=0"):
        CK.load_rank(out, 0, 1, int8_mode=1, enable_int8_weights=True)
    monkeypatch.delenv("FTCF_INT8_LAYOUT")
    CK.load_rank(out, 0, 1, int8_mode=1, enable_int8_weights=True)
