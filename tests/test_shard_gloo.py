"""CPU, world_size 2, gloo: the row-exchange plan of the sharded tables (sml_b200/shard.py).  The local
gather / scatter operators are torch-indexing test doubles; on the GPU they are the CUDA kernels
sml_gather_pairs / sml_scatter_grads and the collectives run over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sml_b200.shard import RowExchange, shard_rows


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        N, W = 101, 8                                             # ragged: 101 rows over 2 ranks
        full = torch.randn(N, W, generator=g)
        shard = shard_rows(full, world, rank)
        ex = RowExchange(world, rank)
        gi = torch.Generator().manual_seed(10 + rank)
        ids = torch.randint(0, N, (37 + 5 * rank,), generator=gi)   # different counts per rank, duplicates inside
        ids[:4] = ids[0]
        plan = ex.plan(ids)
        got = ex.fetch(plan, lambda loc: shard[loc])
        assert torch.equal(got, full[ids]), "fetched rows differ"
        # gradients: every occurrence returns one row; owners accumulate
        grads = torch.randn(ids.numel(), W, generator=gi)
        acc = torch.zeros_like(shard)
        ex.push(plan, grads, lambda loc, rows: acc.index_add_(0, loc, rows))
        torch.save(dict(ids=ids, grads=grads, acc=acc), os.path.join(out_dir, "r%d.pt" % rank))
        # whole-epoch planning: per-step plans identical to planning every step on its own (ragged last step)
        sizes = [11, 11, 11, 4 + rank]
        eids = torch.randint(0, N, (sum(sizes),), generator=gi)
        eids[11:15] = eids[11]
        plans = ex.plan_epoch(eids, sizes)
        o = 0
        for s_, p_ in zip(sizes, plans):
            one = ex.plan(eids[o:o + s_])
            assert torch.equal(p_.order, one.order) and torch.equal(p_.inverse, one.inverse)
            assert p_.send_counts == one.send_counts and p_.recv_counts == one.recv_counts
            assert torch.equal(p_.recv_loc, one.recv_loc) and p_.n == one.n
            assert torch.equal(ex.fetch(p_, lambda loc: shard[loc]), full[eids[o:o + s_]])
            o += s_
        # empty request from one rank must not dead-lock
        plan2 = ex.plan(ids[:0] if rank == 0 else ids[:3])
        got2 = ex.fetch(plan2, lambda loc: shard[loc])
        assert got2.shape[0] == (0 if rank == 0 else 3)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_row_exchange_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    N, W = 101, 8
    ref = torch.zeros(N, W)
    for p in parts:
        ref.index_add_(0, p["ids"], p["grads"])
    for r, p in enumerate(parts):
        assert torch.allclose(p["acc"], ref[r::world], atol=1e-6), "accumulated gradients differ on rank %d" % r


def test_row_exchange_world1_identity():
    ex = RowExchange(1, 0)
    full = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    ids = torch.tensor([3, 3, 9, 0])
    plan = ex.plan(ids)
    assert torch.equal(ex.fetch(plan, lambda loc: full[loc]), full[ids])
    acc = torch.zeros_like(full)
    ex.push(plan, torch.ones(4, 4), lambda loc, rows: acc.index_add_(0, loc, rows))
    assert acc[3, 0] == 2 and acc[9, 0] == 1 and acc[1, 0] == 0
    plans = ex.plan_epoch(torch.tensor([3, 3, 9, 0, 7]), [2, 2, 1])
    assert [p.n for p in plans] == [2, 2, 1]
    assert torch.equal(ex.fetch(plans[1], lambda loc: full[loc]), full[torch.tensor([9, 0])])
