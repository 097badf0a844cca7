"""The reference's UNCHANGED Python driver against the drop-in modules: examples/pytorch/codefuse/codefuse_example.py (byte-compiled
where it lies into oracle/_ref/codefuse_example.compiled by oracle/Makefile -- the GPU box has no /root/reference) is executed under
`torchrun --nproc_per_node 1` exactly as its README does, with --lib_path pointing at fastertransformer4codefuse_b200/lib
(libth_gptneox.so / libth_common.so), a two-layer checkpoint directory written by our converter (checkpoint.py ==
huggingface_convert.py's files) and a small word-level tokenizer directory.  Its printed generations must be the oracle's tokens,
decoded by the same tokenizer, for int8_mode 0 and 1 (the driver quantises at load through OUR libth_common) and for pre-quantised
*.q.bin / *.s.bin files (enable_int8_weights = 1).  The last request is a beam search (beam_width 3, as line 3 of the reference's
input_demo.jsonl, streamed through the callback): its three printed beams must be what our ctypes GptNeoXOp gives in this process on
the same checkpoint (the beam search itself is pinned against the oracle and the reference's kernels in tests/test_beam_search_gpu.py;
a free-running oracle comparison would hinge on 1e-3 score gaps of this random model)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi, checkpoint as CK, weights as W
from oracle import gptneox_ref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "codefuse_example.compiled")
VOCAB, EOS = 96, 95


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny_hf():
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM
    torch.manual_seed(5)
    cfg = GPTNeoXConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, vocab_size=VOCAB,
                        hidden_act="gelu_new", use_parallel_residual=True, max_position_embeddings=128, tie_word_embeddings=False,
                        layer_norm_eps=1e-5, attention_dropout=0.0, hidden_dropout=0.0, bos_token_id=0, eos_token_id=EOS,
                        rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "partial_rotary_factor": 0.5})
    model = GPTNeoXForCausalLM(cfg).eval()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "layernorm" in name or "layer_norm" in name:
                p.add_(torch.randn_like(p) * 0.05)
            elif p.dim() == 1:
                p.normal_(0.0, 0.05)
            else:
                p.normal_(0.0, 0.08)
            p.copy_(p.half().float())
        model.embed_out.weight[EOS].zero_()          # end_id never wins: every request generates out_seq_length tokens
    return model


def _tokenizer_dir(path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    # the special tokens must not be substrings of ordinary words: added tokens are matched before the pre-tokenizer runs ("w18" would
    # split into the special token "w1" + "8")
    names = {0: "<s>", 1: "<unk>", EOS: "</s>"}
    vocab = {names.get(i, f"w{i}"): i for i in range(VOCAB)}
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", eos_token="</s>", bos_token="<s>")
    fast.save_pretrained(path)
    return fast


def _oracle(ckpt_dir, int8_mode):
    cfg, w, _, _ = CK.load_rank(ckpt_dir, 0, 1, int8_mode=0)
    L = cfg.layer_num
    rcfg = R.RefConfig(head_num=cfg.head_num, size_per_head=cfg.size_per_head, inter_size=cfg.inter_size, layer_num=L,
                       vocab_size=cfg.vocab_size, rotary_embedding_dim=cfg.rotary_embedding_dim, start_id=cfg.start_id, end_id=cfg.end_id,
                       tensor_para_size=1, int8_mode=int8_mode, use_gptj_residual=cfg.use_gptj_residual)
    q, s = [None] * (4 * L), [None] * (4 * L)
    if int8_mode == 1:
        for kind, f in enumerate(W.KIND_FIELDS):
            for l in range(L):
                _, sc, plain = W.quantize_on_device(w[f * L + l])
                q[kind * L + l], s[kind * L + l] = plain.numpy(), sc
    return R.GptNeoXRef(rcfg, [R.RankWeights(w=w, q=q, scale=s)])


def _run_driver(tmp_path, ckpt, tok_dir, requests, int8_mode, enable_int8_weights):
    inp = tmp_path / f"input_{int8_mode}_{enable_int8_weights}.jsonl"
    inp.write_text("\n".join(json.dumps(r) for r in requests) + "\n")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1", "--master-port",
           str(_free_port()), DRIVER, "--world_size", "1", "--lib_path", capi.LIB_DIR, "--ckpt_path", str(ckpt), "--tokenizer_path",
           str(tok_dir), "--int8_mode", str(int8_mode), "--enable_int8_weights", str(enable_int8_weights), "--input_file", str(inp)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout + "\n" + r.stderr)[-6000:]
    results, lines = [], r.stdout.splitlines()
    for i, line in enumerate(lines):
        if line.strip() == "- result":
            results.append(lines[i + 1].strip())
    return results, r.stdout


@pytest.mark.parametrize("int8_mode,enable_int8_weights", [(0, 0), (1, 0), (1, 1)])
def test_unchanged_reference_driver(cuda, tmp_path, int8_mode, enable_int8_weights):
    if not os.path.exists(DRIVER):
        pytest.skip(f"{DRIVER} is missing (oracle/Makefile builds it where /root/reference exists)")
    model = _tiny_hf()
    fp_dir = CK.convert_hf(model, str(tmp_path / "ft"), 1, "fp16")
    ckpt = fp_dir
    if enable_int8_weights:
        ckpt = str(tmp_path / "ft" / "1-gpu-int8")
        CK.quantize_dir(fp_dir, ckpt, 1)
    tok = _tokenizer_dir(str(tmp_path / "tok"))
    g = np.random.default_rng(8)
    prompts = [[int(x) for x in g.integers(2, VOCAB - 2, size=n)] for n in (9, 5, 12)]
    text = lambda ids: " ".join(f"w{i}" for i in ids)
    requests = [
        {"prompts": [{"prompt": text(prompts[0]), "top_k": 1}], "out_seq_length": 8},                               # greedy, batch 1
        {"prompts": [{"prompt": text(prompts[1]), "top_k": 1}, {"prompt": text(prompts[2]), "top_k": 1}], "out_seq_length": 6},   # ragged batch
        {"prompts": [{"prompt": text(prompts[0]), "top_k": 40, "top_p": 0.9, "temperature": 0.2, "repetition_penalty": 1.1,
                      "random_seed": 7}], "out_seq_length": 8},                                                    # input_demo.jsonl-style sampling
        {"prompts": [{"prompt": text(prompts[1])}, {"prompt": text(prompts[2])}], "out_seq_length": 6, "beam_width": 3,
         "stream": True},                                                                                        # beam search, streamed
    ]
    got, out = _run_driver(tmp_path, ckpt, tmp_path / "tok", requests, int8_mode, enable_int8_weights)
    ref = _oracle(fp_dir, int8_mode)

    def expect(ids_list, out_len, **kw):
        S = max(len(x) for x in ids_list)
        ids = np.full((len(ids_list), S), EOS, dtype=np.int32)
        for b, x in enumerate(ids_list):
            ids[b, :len(x)] = x
        lens = [len(x) for x in ids_list]
        res = ref.forward(ids, lens, out_len, return_cum_log_probs=1, **kw)["output_ids"]
        return [tok.decode([int(t) for t in res[b, 0, lens[b]:lens[b] + out_len]]) for b in range(len(ids_list))]

    want = expect([prompts[0]], 8, top_k=[1], top_p=[0.0])
    want += expect([prompts[1], prompts[2]], 6, top_k=[1, 1], top_p=[0.0, 0.0])
    want += expect([prompts[0]], 8, top_k=[40], top_p=[0.9], temperature=[0.2], repetition_penalty=[1.1], random_seed=[7])
    # The same requests through the ctypes mirror of the op in THIS process, on the same checkpoint files: the driver's output must be
    # exactly that (the subprocess differs only in the host path: reference loader -> libth_common -> libth_gptneox).  The beam request
    # is only compared this way (the beam search itself is pinned in tests/test_beam_search_gpu.py; a free-running oracle comparison
    # would hinge on 1e-3 score gaps of this random model).
    from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp
    dev = cuda
    cfg2, w2, q2, s2 = CK.load_rank(ckpt, 0, 1, int8_mode=int8_mode, enable_int8_weights=bool(enable_int8_weights))
    op = GptNeoXOp(None, 0, cfg2.head_num, cfg2.size_per_head, cfg2.inter_size, cfg2.layer_num, cfg2.vocab_size, cfg2.rotary_embedding_dim,
                   cfg2.start_id, cfg2.end_id, 1, 1, int8_mode, 1024, bool(cfg2.use_gptj_residual),
                   [x.to(dev) for x in w2], [x.to(dev) for x in q2], [x.to(dev) for x in s2])

    def ours(ids_list, out_len, beam=1, **kw):
        S = max(len(x) for x in ids_list)
        ids = np.full((len(ids_list), S), EOS, dtype=np.int32)
        for b, x in enumerate(ids_list):
            ids[b, :len(x)] = x
        lens = [len(x) for x in ids_list]
        targs = {k: torch.tensor(v, dtype=torch.int32 if k == "top_k" else torch.int64 if k == "random_seed" else torch.float32)
                 for k, v in kw.items()}
        res = op.forward(torch.from_numpy(ids).to(dev), torch.tensor(lens, dtype=torch.int32, device=dev), out_len, beam_width=beam,
                         return_cum_log_probs=1, **targs)[0].cpu().numpy()
        texts = []
        for b, n in enumerate(lens):
            for j in range(beam):
                gen = [int(t) for t in res[b, j, n:]]
                texts.append(tok.decode(gen[:gen.index(EOS)] if EOS in gen else gen))
        return texts

    mine = ours([prompts[0]], 8, top_k=[1], top_p=[0.0])
    mine += ours([prompts[1], prompts[2]], 6, top_k=[1, 1], top_p=[0.0, 0.0])
    mine += ours([prompts[0]], 8, top_k=[40], top_p=[0.9], temperature=[0.2], repetition_penalty=[1.1], random_seed=[7])
    mine += ours([prompts[1], prompts[2]], 6, beam=3)
    assert len(got) == len(mine) == 4 + 6
    assert got == mine, f"driver printed {got}\nour op gives  {mine}\n---- driver output ----\n{out[-1500:]}"
    assert mine[:4] == want, f"our op gives {mine[:4]}\noracle gives {want}"
