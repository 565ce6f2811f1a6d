"""World-size-2 gloo test (CPU) of the host-side plumbing of the class-sharded head: the collective
wrapper used by PartialFC and the merge of per-rank softmax statistics it feeds."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msml_b200.headers._comm import TorchDistComm
    from oracle import partial_fc as opfc
    comm = TorchDistComm(world, rank)
    B, D, C = 4, 8, 37
    g = torch.Generator().manual_seed(7)
    labels_all = torch.randint(0, C, (world * B,), generator=g)
    feats_all = torch.randn(world * B, D, generator=g)
    # all_gather: labels and features
    tl = torch.zeros(world * B, dtype=torch.long)
    comm.all_gather(tl, labels_all[rank * B:(rank + 1) * B])
    x = torch.zeros(world * B, D)
    comm.all_gather(x, feats_all[rank * B:(rank + 1) * B])
    ok = torch.equal(tl, labels_all) and torch.equal(x, feats_all)
    # shard geometry + remap as each rank's PartialFC computes them
    num_local, class_start, _ = opfc.shard_geometry(C, world, rank)
    remapped = opfc.remap_labels(tl.numpy(), class_start, num_local)
    # stats all_gather layout (W*3, B_tot) consumed by msml_head_merge_stats
    stats = torch.full((3, world * B), float(rank))
    gathered = torch.zeros(world * 3, world * B)
    comm.all_gather(gathered, stats)
    ok = ok and all(float(gathered[r * 3 + j, 0]) == r for r in range(world) for j in range(3))
    # reduce_scatter: sum over ranks, own slice
    dx_full = torch.arange(world * B * D, dtype=torch.float32).reshape(world * B, D) * (rank + 1)
    x_grad = torch.zeros(B, D)
    comm.reduce_scatter(x_grad, dx_full)
    want = torch.arange(world * B * D, dtype=torch.float32).reshape(world * B, D)[rank * B:(rank + 1) * B] * sum(range(1, world + 1))
    ok = ok and torch.equal(x_grad, want)
    q.put((rank, bool(ok), remapped.tolist(), num_local, class_start))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    world, port = 2, 29733
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(g[1] for g in got)
    # every label is owned by exactly one rank after the remap
    owned = np.array([[v != -1 for v in g[2]] for g in got])
    assert (owned.sum(axis=0) == 1).all()
    assert got[0][3] + got[1][3] == 37 and got[1][4] == got[0][3]
