"""world_size-2 `gloo` tests (CPU) of the host-side tensor-parallel logic: the per-rank weight shards reproduce the
unsplit layer when the partial sums are all-reduced where the reference all-reduces (GptNeoXDecoder.cc:348-359), and the
NCCL unique id made on rank 0 reaches every rank through the torch process group (the replacement for
th_op/gptneox/utils/nccl_inherit_utils.cc:8-68)."""
import ctypes as C
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastertransformer4codefuse_b200 import weights as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeLib:
    """Stands in for libftcf on rank 0: writes a recognisable 128-byte id."""

    @staticmethod
    def ftcf_nccl_unique_id(buf):
        C.memmove(buf, bytes(range(128)), 128)
        return 0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp
        uid = GptNeoXOp._exchange_nccl_id(_FakeLib, dist.group.WORLD, rank)
        ok_uid = bytes(uid.raw[:128]) == bytes(range(128))

        cfg = W.NeoXConfig(head_num=4, size_per_head=16, inter_size=128, layer_num=1, vocab_size=64, rotary_embedding_dim=8,
                           start_id=0, end_id=63, use_gptj_residual=True)
        full = {k: v.float() for k, v in W.synthetic_layer(cfg, 0, "cpu", seed=9).items()}
        mine = [t.float() for t in W.split_layer(W.synthetic_layer(cfg, 0, "cpu", seed=9), cfg, world, rank)]
        torch.manual_seed(1)
        x = torch.randn(3, cfg.hidden)
        ctx = torch.randn(3, cfg.hidden)                 # stands for the attention output, heads sharded over ranks
        hl = cfg.hidden // world
        # column-parallel QKV: this rank's [q | k | v] columns are the rank's head slice of each third
        qkv_full = (x @ full["qkv_w"] + full["qkv_b"]).reshape(3, 3, cfg.hidden)[:, :, rank * hl:(rank + 1) * hl].reshape(3, 3 * hl)
        ok_qkv = torch.allclose(x @ mine[W.QKV_W] + mine[W.QKV_B], qkv_full, atol=1e-5)
        # row-parallel O and FFN2 + the pre-divided summed bias + x / t: all-reduce restores the unsplit layer output
        inter = torch.nn.functional.gelu(x @ mine[W.FFN1_W] + mine[W.FFN1_B], approximate="tanh")
        part = ctx[:, rank * hl:(rank + 1) * hl] @ mine[W.O_W] + inter @ mine[W.FFN2_W] + mine[W.FFN2_B] + x / world
        dist.all_reduce(part)
        inter_full = torch.nn.functional.gelu(x @ full["ffn1_w"] + full["ffn1_b"], approximate="tanh")
        exp = ctx @ full["o_w"] + full["o_b"] + inter_full @ full["ffn2_w"] + full["ffn2_b"] + x
        ok_sum = torch.allclose(part, exp, atol=2e-3)
        q.put((rank, ok_uid, ok_qkv, ok_sum))
    finally:
        dist.destroy_process_group()


def test_tensor_parallel_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok_uid, ok_qkv, ok_sum in res:
        assert ok_uid, f"rank {rank}: NCCL id did not arrive"
        assert ok_qkv, f"rank {rank}: QKV column shard mismatch"
        assert ok_sum, f"rank {rank}: all-reduced row-parallel sum != unsplit layer"
