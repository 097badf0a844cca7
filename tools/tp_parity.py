#!/usr/bin/env python
"""Tensor-parallel parity on N real GPUs (torchrun --nproc-per-node N tools/tp_parity.py): every rank builds its weight shard
of a small model, runs GptNeoXOp with tensor_para_size = N over NCCL, and rank 0 compares the output ids with the oracle's
TP emulation (partial sums added where the reference all-reduces).  int8 and fp16, parallel and sequential residual."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from fastertransformer4codefuse_b200 import weights as W
from fastertransformer4codefuse_b200.gptneox_op import GptNeoXOp
from helpers import oracle_from_rank_weights, tiny_cfg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for int8_mode in (1, 0):
    for gptj in (True, False):
        cfg = tiny_cfg(head_num=8, use_gptj_residual=gptj)
        shards = [W.make_synthetic(cfg, world, r, int8_mode, "cpu", seed=4, keep_plain=True) for r in range(world)]
        mine = shards[rank]
        w, q, s = mine.lists()
        op = GptNeoXOp(dist.group.WORLD, rank, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size,
                       cfg.rotary_embedding_dim, cfg.start_id, cfg.end_id, world, 1, int8_mode, 1024, gptj,
                       [x.to(dev) for x in w], [x.to(dev) for x in q], [x.to(dev) for x in s])
        lens = [11, 6, 9]
        g = np.random.default_rng(5)
        ids = g.integers(0, cfg.vocab_size - 1, size=(3, 11)).astype(np.int32)
        for b, n in enumerate(lens):
            ids[b, n:] = cfg.vocab_size - 1
        for graph in (0, 1):
            op.set_option("cuda_graph", graph)
            res = op.forward(torch.from_numpy(ids).to(dev), torch.tensor(lens, dtype=torch.int32, device=dev), 10)
            got = res[0].cpu().numpy()
            if rank == 0:
                ref = oracle_from_rank_weights(cfg, shards, int8_mode)
                exp = ref.forward(ids, lens, 10)
                same = np.array_equal(got, exp["output_ids"]) and np.array_equal(res[1].cpu().numpy(), exp["sequence_lengths"])
                print(f"tp={world} int8={int8_mode} parallel_residual={gptj} graph={graph}: {'OK' if same else 'MISMATCH'}", flush=True)
                ok &= same
        del op
dist.barrier()
if rank == 0:
    print("TP PARITY", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
