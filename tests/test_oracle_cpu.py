"""CPU tests (no GPU): the oracle against the reference's own golden vectors / object code and the committed fixtures.

  * quantiser KATs copied in spirit from the reference's unit tests
    (tests/weight_only_quant_ops/th_weight_quant_ops_unit_tests.py:36-37,110-116,133);
  * oracle quantiser vs outputs of the REFERENCE's object code (tests/golden/quant_ref_*.npz, and live against
    oracle/_ref/libref_quant.so when it is present);
  * oracle model wiring vs HuggingFace GPTNeoXForCausalLM goldens (tests/golden/neox_hf_*.npz).
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch

from oracle import gptneox_ref as R
from oracle import quant_ref as Q
from oracle import sampling_ref as S

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_quant.so")


# ------------------------------------------------------------------------------------------------ quantiser KATs
def test_row_permutation_map_kat():
    # th_weight_quant_ops_unit_tests.py:36-37 -- each group of 16 rows is permuted with this map
    perm = [0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15]
    q = np.arange(32, dtype=np.int8).reshape(32, 1).repeat(8, axis=1)
    out = Q.permute_b_rows(q)
    exp = np.concatenate([np.asarray(perm), 16 + np.asarray(perm)]).astype(np.int8)
    assert np.array_equal(out[:, 0], exp)


def test_add_bias_and_interleave_kat():
    # th_weight_quant_ops_unit_tests.py:110-116
    src = np.asarray([-104, -70, -36, 127, 16, 50, 84, 118], dtype=np.int8)
    exp = (np.asarray([-104, -36, -70, 127, 16, 84, 50, 118], dtype=np.int16) + 128).astype(np.uint8)
    assert np.array_equal(Q.add_bias_and_interleave_int8s(src).view(np.uint8), exp)


def test_subbyte_transpose_is_plain_transpose():
    # th_weight_quant_ops_unit_tests.py:133 -- int8 "subbyte" transpose == permute([0, 2, 1])
    g = np.random.default_rng(0)
    q = g.integers(-128, 128, size=(64, 32)).astype(np.int8)
    assert np.array_equal(Q.subbyte_transpose_int8(q).reshape(-1), np.ascontiguousarray(q.T).reshape(-1))   # bytes move, shape kept


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(GOLD, "quant_ref_*.npz"))))
def test_oracle_quantiser_vs_reference_goldens(fixture):
    z = np.load(fixture)
    q, scale = Q.symmetric_quantize_unprocessed(z["w"])
    assert np.array_equal(q, z["unprocessed"]), "unprocessed int8 differs from the reference's"
    assert np.array_equal(scale.astype(np.float16).view(np.uint16), z["scales"].view(np.uint16)), "fp16 scales differ"
    assert np.array_equal(Q.preprocess_weights_ampere(q).reshape(-1), z["processed"].reshape(-1)), "sm80 layout differs"
    k, n = q.shape
    # the loader path for reference-made *.q.bin: sm80 bytes -> plain int8 -> B200 layout
    back = Q.unprocess_weights_ampere(z["processed"], k, n)
    assert np.array_equal(back, q)
    assert np.array_equal(Q.from_b200_layout(Q.to_b200_layout(q), k, n), q)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_quant.so not built (needs /root/reference)")
@pytest.mark.parametrize("k,n,seed", [(64, 64, 1), (128, 192, 2), (256, 64, 3)])
def test_oracle_quantiser_vs_reference_object_code(k, n, seed):
    lib = C.CDLL(REF_SO)
    g = np.random.default_rng(seed)
    w = (g.standard_normal((k, n)) * 0.05).astype(np.float16)
    w[0, 0] = 0.0
    w[:, 1] = 0.0 if n > 1 else w[:, 1]           # an all-zero column: scale 0, q = NaN-path of the reference
    proc, unproc, scales = np.empty((k, n), np.int8), np.empty((k, n), np.int8), np.empty(n, np.float16)
    shape = (C.c_size_t * 2)(k, n)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    assert lib.ref_symmetric_quantize_half(vp(proc), vp(unproc), vp(scales), vp(w), shape, 2) == 0
    q, scale = Q.symmetric_quantize_unprocessed(w)
    assert np.array_equal(q, unproc)
    assert np.array_equal(scale.astype(np.float16).view(np.uint16), scales.view(np.uint16))
    assert np.array_equal(Q.preprocess_weights_ampere(q).reshape(-1), proc.reshape(-1))


# ------------------------------------------------------------------------------------------------ model wiring vs HF
def _oracle_from_fixture(z):
    heads, dh, inter, L, vocab, rot, end_id, parallel = [int(x) for x in z["meta"]]
    cfg = R.RefConfig(head_num=heads, size_per_head=dh, inter_size=inter, layer_num=L, vocab_size=vocab,
                      rotary_embedding_dim=rot, start_id=0, end_id=end_id, tensor_para_size=1, int8_mode=0,
                      use_gptj_residual=bool(parallel))
    w = [torch.from_numpy(z[f"w{i}"]) for i in range(12 * L + 4)]
    return cfg, R.GptNeoXRef(cfg, [R.RankWeights(w=w)])


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(GOLD, "neox_hf_*.npz"))))
def test_oracle_model_vs_huggingface_golden(fixture):
    z = np.load(fixture)
    cfg, ref = _oracle_from_fixture(z)
    out_len = int(z["out_len"])
    lens = z["lens"]
    res = ref.forward(z["ids"], lens, out_len, keep_logits=True)
    B, S = z["ids"].shape
    # While the generated prefix is identical on both sides, the logits of the next step must agree.  Tolerance: the
    # oracle rounds every activation to fp16 where the reference stores half (HF runs fp32); logits here are O(1).
    # Greedy ids must be identical wherever HF's top-2 margin is above twice that tolerance.
    compared = 0
    for b in range(B):
        got = res["output_ids"][b, 0, lens[b]:lens[b] + out_len]
        for s in range(out_len):
            np.testing.assert_allclose(res["logits"][s][b], z["logits"][b, s], rtol=0, atol=4e-2)
            compared += 1
            top2 = np.sort(z["logits"][b, s])[-2:]
            if top2[1] - top2[0] > 0.08:
                assert got[s] == z["gen"][b, s], f"{os.path.basename(fixture)} row {b} step {s}: {got[s]} != {z['gen'][b, s]}"
            elif got[s] != z["gen"][b, s]:
                break                      # a legitimate near-tie: the prefixes differ from here on
    assert compared >= 3, "fixture too ambiguous to pin anything"
    assert np.array_equal(res["sequence_lengths"].reshape(-1), np.full(B, S + out_len))


def test_oracle_tensor_parallel_split_matches_single_rank():
    """TP emulation: splitting the same weights over 2 ranks reproduces the t = 1 tokens (wiring of the column / row
    splits, bias / t, x / t in the residual -- huggingface_convert.py:35-82, add_residual_kernels.cu:116-152)."""
    from fastertransformer4codefuse_b200 import weights as W
    cfg = W.NeoXConfig(head_num=4, size_per_head=16, inter_size=128, layer_num=2, vocab_size=96, rotary_embedding_dim=8,
                       start_id=0, end_id=95, use_gptj_residual=True)
    ids = np.random.default_rng(5).integers(0, 95, size=(2, 6)).astype(np.int32)
    outs = []
    for t in (1, 2):
        ranks = [W.make_synthetic(cfg, t, r, 0, "cpu", seed=3) for r in range(t)]
        rcfg = R.RefConfig(head_num=4, size_per_head=16, inter_size=128, layer_num=2, vocab_size=96, rotary_embedding_dim=8,
                           start_id=0, end_id=95, tensor_para_size=t, int8_mode=0, use_gptj_residual=True)
        ref = R.GptNeoXRef(rcfg, [R.RankWeights(w=list(rw.w)) for rw in ranks])
        outs.append(ref.forward(ids, [6, 4], 4, keep_logits=True))
    np.testing.assert_allclose(outs[0]["logits"][0], outs[1]["logits"][0], atol=3e-2, rtol=0)


# ------------------------------------------------------------------------------------------------ sampling oracle
def test_topk_setup_rules():
    # TopKSamplingLayer.cu:28-78: k = 0 & p = 0 -> k = 1; k > 0 & p = 0 -> p = 1; k clipped to 1024
    ks, ps, _ = S.setup_topk_runtime_args(np.asarray([0, 5, 2000, 3]), np.asarray([0.0, 0.0, 0.5, 0.9], np.float32), 4)
    assert list(ks) == [1, 5, 1024, 3]
    assert np.allclose(ps, [1.0, 1.0, 0.5, 0.9])


def test_topk_sampling_membership_and_greedy():
    # tests/unittests/test_sampling_kernels.cu: a sampled id always belongs to the top-k set; k = 1 is arg-max
    g = np.random.default_rng(7)
    for trial in range(20):
        row = g.standard_normal(300).astype(np.float32)
        k = int(g.integers(1, 9))
        rng = S.CurandXorwow(trial)
        tok, _ = S.topk_sampling_row(row.copy(), k, np.float32(1.0), rng, k, False)
        assert tok in set(np.argsort(-row)[:k])
        tok1, _ = S.topk_sampling_row(row.copy(), 1, np.float32(1.0), S.CurandXorwow(0), 1, False)
        assert tok1 == int(np.argmax(row))


def test_repetition_penalty_and_end_mask():
    logits = np.asarray([[2.0, -2.0, 1.0, 0.5]], np.float32)
    out_ids = np.asarray([[0], [1], [0]], np.int64)          # time-major [step, B]; prompt length 2, one generated
    S.apply_repetition_penalty(logits, np.asarray([2.0], np.float32), out_ids, np.asarray([2]), 2, 3)
    assert np.allclose(logits, [[1.0, -4.0, 1.0, 0.5]])      # each distinct id once: >0 divided, <0 multiplied
    lg = np.zeros((2, 8), np.float32)
    S.add_bias_end_mask(lg, np.asarray([3, 3]), np.asarray([True, False]), 6)
    assert lg[0, 3] == np.finfo(np.float32).max and lg[0, 0] == -np.finfo(np.float32).max
    assert lg[1, 6] == -np.finfo(np.float32).max and lg[1, 0] == 0.0


def test_gather_output_removes_pad_gap():
    # decoding_kernels.cu:519-560: ragged prompt lengths, pad gap [len, S) removed, tail filled with end_id
    S_in, out, end = 4, 3, 9
    step_ids = np.zeros((S_in + out, 2), np.int64)
    step_ids[:, 0] = [1, 2, 3, 4, 5, 6, 7]
    step_ids[:, 1] = [1, 2, end, end, 5, 6, 7]
    ids, lens = R.gather_output(step_ids, np.asarray([6, 6]), np.asarray([4, 2]), S_in, S_in + out, end)
    assert list(ids[0, 0]) == [1, 2, 3, 4, 5, 6, 7]
    assert list(ids[1, 0]) == [1, 2, 5, 6, 7, end, end]
    assert list(lens.reshape(-1)) == [7, 7]
