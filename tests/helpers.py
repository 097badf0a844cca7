"""Shared helpers for the parity tests (oracle on one side, the C ABI on the other)."""
import numpy as np
import torch

from fastertransformer4codefuse_b200 import weights as W
from oracle import gptneox_ref as R


def stream():
    return torch.cuda.current_stream().cuda_stream


def tiny_cfg(**kw):
    base = dict(head_num=4, size_per_head=64, inter_size=512, layer_num=2, vocab_size=512, rotary_embedding_dim=32,
                start_id=0, end_id=511, use_gptj_residual=True)
    base.update(kw)
    return W.NeoXConfig(**base)


def oracle_from_rank_weights(cfg, ranks, int8_mode):
    """Build the oracle model from the very tensors (CPU) the op receives."""
    rcfg = R.RefConfig(head_num=cfg.head_num, size_per_head=cfg.size_per_head, inter_size=cfg.inter_size,
                       layer_num=cfg.layer_num, vocab_size=cfg.vocab_size, rotary_embedding_dim=cfg.rotary_embedding_dim,
                       start_id=cfg.start_id, end_id=cfg.end_id, tensor_para_size=len(ranks), int8_mode=int8_mode,
                       use_gptj_residual=cfg.use_gptj_residual)
    rws = []
    for rw in ranks:
        q = [x.numpy() if x is not None else None for x in rw.plain_q]
        rws.append(R.RankWeights(w=[x for x in rw.w], q=q, scale=list(rw.scale)))
    return R.GptNeoXRef(rcfg, rws)


def to_cuda_lists(rw, dev):
    w, q, s = rw.lists()
    return [x.to(dev) for x in w], [x.to(dev) for x in q], [x.to(dev) for x in s]


def assert_close(name, got, ref, rtol, atol, frac_ok=0.0):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    bad = err > (atol + rtol * np.abs(ref))
    nbad = int(bad.sum())
    assert nbad <= frac_ok * bad.size, (f"{name}: {nbad}/{bad.size} elements out of tolerance (rtol {rtol}, atol {atol}); "
                                        f"max abs err {err.max():.4g} at ref {ref.flat[int(err.argmax())]:.4g}")
