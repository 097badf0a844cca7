#!/usr/bin/env python
"""Tensor-parallel parity on N real GPUs (torchrun --nproc-per-node N tools/tp_parity.py): every rank builds its weight shard
of a small model, runs GptNeoXOp with tensor_para_size = N over NCCL / NVLink, and rank 0 compares the output ids with the
oracle's TP emulation (partial sums added where the reference all-reduces, GptNeoXDecoder.cc:348-359).  int8 and fp16,
parallel and sequential residual, graph on / off; plus
  * two same-shape requests with different ragged lengths back to back (a cached decode graph must not see stale offsets);
  * a request where every row hits end_id after a few tokens of a long output_len: all ranks must leave the decode loop at
    the same step (a rank that launched one more step would hang in its collectives), and the next request must still work.
Driven by tests/test_tp_gpu.py; exit code 0 = all match."""
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


def report(name, same):
    global ok
    if rank == 0:
        print(f"tp={world} {name}: {'OK' if same else 'MISMATCH'}", flush=True)
        ok &= bool(same)


def make_op(cfg, shards, int8_mode, gptj):
    w, q, s = shards[rank].lists()
    return GptNeoXOp(dist.group.WORLD, rank, cfg.head_num, cfg.size_per_head, cfg.inter_size, cfg.layer_num, cfg.vocab_size,
                     cfg.rotary_embedding_dim, cfg.start_id, cfg.end_id, world, 1, int8_mode, 1024, gptj,
                     [x.to(dev) for x in w], [x.to(dev) for x in q], [x.to(dev) for x in s])


def prompts(cfg, lens, S, seed):
    g = np.random.default_rng(seed)
    ids = g.integers(0, cfg.vocab_size - 1, size=(len(lens), S)).astype(np.int32)
    for b, n in enumerate(lens):
        ids[b, n:] = cfg.vocab_size - 1
    return ids


def run(op, ids, lens, out_len):
    res = op.forward(torch.from_numpy(ids).to(dev), torch.tensor(lens, dtype=torch.int32, device=dev), out_len)
    return res[0].cpu().numpy(), res[1].cpu().numpy()


# every rank's slice must keep k a multiple of 128 for the int8 GEMMs: hidden / world and inter_size / world >= 128
heads = max(8, 2 * world)
inter = max(512, 128 * world)
for int8_mode in (1, 0):
    for gptj in (True, False):
        cfg = tiny_cfg(head_num=heads, inter_size=inter, use_gptj_residual=gptj)
        shards = [W.make_synthetic(cfg, world, r, int8_mode, "cpu", seed=4, keep_plain=True) for r in range(world)]
        op = make_op(cfg, shards, int8_mode, gptj)
        ref = oracle_from_rank_weights(cfg, shards, int8_mode) if rank == 0 else None
        lens = [11, 6, 9]
        ids = prompts(cfg, lens, 11, 5)
        for graph in (0, 1):
            op.set_option("cuda_graph", graph)
            got, got_len = run(op, ids, lens, 10)
            if rank == 0:
                exp = ref.forward(ids, lens, 10)
                report(f"int8={int8_mode} parallel_residual={gptj} graph={graph}",
                       np.array_equal(got, exp["output_ids"]) and np.array_equal(got_len, exp["sequence_lengths"]))
        # same shape, other ragged lengths, graph replayed from the cache
        lens2 = [4, 11, 7]
        ids2 = prompts(cfg, lens2, 11, 6)
        got, got_len = run(op, ids2, lens2, 10)
        if rank == 0:
            exp = ref.forward(ids2, lens2, 10)
            report(f"int8={int8_mode} parallel_residual={gptj} second ragged request on the cached graph",
                   np.array_equal(got, exp["output_ids"]) and np.array_equal(got_len, exp["sequence_lengths"]))
        if int8_mode == 1 and gptj:
            # more than 4 rows: the decode step leaves the fused path (run_layer: push GEMMs with global split-K partials + stand-alone
            # gather, CTA-share hints) -- batch 6, ragged
            lens6 = [11, 6, 9, 3, 10, 8]
            ids6 = prompts(cfg, lens6, 11, 7)
            for graph in (0, 1):
                op.set_option("cuda_graph", graph)
                got, got_len = run(op, ids6, lens6, 8)
                if rank == 0:
                    exp = ref.forward(ids6, lens6, 8)
                    report(f"int8=1 parallel_residual=True batch 6 graph={graph}",
                           np.array_equal(got, exp["output_ids"]) and np.array_equal(got_len, exp["sequence_lengths"]))
        del op

# ---- every row finishes early: end_id := the third token the oracle generates for a single-row request
cfg = tiny_cfg(head_num=heads, inter_size=inter)
shards = [W.make_synthetic(cfg, world, r, 1, "cpu", seed=9, keep_plain=True) for r in range(world)]
lens, out_len = [7], 40
ids = prompts(cfg, lens, 7, 3)
box = [None]
if rank == 0:
    free = oracle_from_rank_weights(cfg, shards, 1).forward(ids, lens, 6)["output_ids"][0, 0, 7:13]
    firsts = [(i, int(t)) for i, t in enumerate(free) if int(t) not in [int(x) for x in free[:i]]]   # first occurrence of every token
    late = [t for i, t in firsts if i >= 2]
    box[0] = late[0] if late else firsts[-1][1]      # finish after a few tokens if the continuation allows it, else as late as it does
dist.broadcast_object_list(box, src=0)
if box[0] is None:
    report("early end_id: no usable token in the free-running continuation (test not exercised)", False)
else:
    cfg = tiny_cfg(head_num=heads, inter_size=inter, end_id=box[0])
    op = make_op(cfg, shards, 1, True)
    for graph in (1, 0):
        op.set_option("cuda_graph", graph)
        got, got_len = run(op, ids, lens, out_len)
        steps = op.last_stats["steps"]
        all_steps = [None] * world
        dist.all_gather_object(all_steps, steps)
        if rank == 0:
            exp = oracle_from_rank_weights(cfg, shards, 1).forward(ids, lens, out_len)
            report(f"early end_id graph={graph}: ids", np.array_equal(got, exp["output_ids"]) and np.array_equal(got_len, exp["sequence_lengths"]))
            report(f"early end_id graph={graph}: every rank ran {all_steps} steps (< {out_len})", len(set(all_steps)) == 1 and steps < out_len)
    # the engine is still usable afterwards (no rank is stuck in a collective)
    lens3 = [7, 5]
    ids3 = prompts(cfg, lens3, 7, 8)
    got, got_len = run(op, ids3, lens3, 6)
    if rank == 0:
        exp = oracle_from_rank_weights(cfg, shards, 1).forward(ids3, lens3, 6)
        report("request after the early exit", np.array_equal(got, exp["output_ids"]) and np.array_equal(got_len, exp["sequence_lengths"]))
    del op

dist.barrier()
if rank == 0:
    print("TP PARITY", "PASS" if ok else "FAIL", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
