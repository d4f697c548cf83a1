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


# ---------------------------------------------------------------------------------------------------------
# the whole sharded MF step (exchange + row-lazy Adam bookkeeping + global-batch scaling) on CPU: ShardedSML with the numpy
# stand-in operators of tests/fake_ops.py, two gloo ranks, against the oracle's dense single-process step
def _sharded_worker(rank, world, port, out_dir):
    import types
    from oracle import sml_oracle as O
    from sml_b200.shard import ShardedSML
    from tests import fake_ops
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = _sharded_case()
        fake_ops.THETA[:] = [cfg["tu"], cfg["ti"]]
        tr = types.SimpleNamespace(theta=torch.zeros(8), theta_grad=torch.zeros(8), variant=0)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        s = ShardedSML(shard_rows(T(cfg["ut"]), world, rank), shard_rows(T(cfg["it"]), world, rank), tr, world=world, rank=rank,
                       mf_lr=cfg["lr"], l2=cfg["l2"], ops=fake_ops)
        s.last_user.mul_(0.9); s.last_item.mul_(0.9)                # w_{t-1} != w_hat
        u, i, j = (T(x[rank]) for x in (cfg["u"], cfg["i"], cfg["j"]))
        loss = s.mf_epoch(u, i, j, cfg["B"])
        s.flush()
        res = dict(user=s.user.clone(), item=s.item.clone(), m_user=s.m_user.clone(), stamp=s.stamp_user.clone(), loss=float(loss))
        # transfer epoch on the snapshots: theta gradients all-reduced, replicated Adam with coupled L2
        flat, views = fake_ops.flat_theta(cfg["tu"], cfg["ti"])
        fake_ops.THETA[:] = views
        tr.theta, tr.theta_grad = flat, torch.zeros_like(flat)
        s.m_theta, s.v_theta = torch.zeros_like(flat), torch.zeros_like(flat)
        s.save_hat()
        res["tr_loss"] = float(s.tr_epoch(u, i, j, cfg["B"]))
        res["theta"] = flat.clone()
        torch.save(res, os.path.join(out_dir, "s%d.pt" % rank))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _sharded_case():
    from oracle import sml_oracle as O
    rng = np.random.default_rng(21)
    U, I, B, n = 37, 53, 12, 12 * 2 + 5                                 # per rank: two full steps and a ragged third
    return dict(ut=(rng.standard_normal((U, 64)) * 0.3).astype(np.float32), it=(rng.standard_normal((I, 64)) * 0.3).astype(np.float32),
                tu=O.init_theta(np.random.default_rng(1)), ti=O.init_theta(np.random.default_rng(2)), lr=0.01, l2=1e-4, B=B, n=n,
                u=[rng.integers(0, 9, n), rng.integers(0, U, n)],         # rank 0 hammers a few users: duplicates + lazy catch-up
                i=[rng.integers(0, I, n), rng.integers(0, I, n)], j=[rng.integers(0, I, n), rng.integers(0, I, n)])


def test_sharded_mf_epoch_world2_matches_dense_oracle(tmp_path):
    from oracle import sml_oracle as O
    world = 2
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cfg = _sharded_case()
    st = dict(last_user=cfg["ut"] * np.float32(0.9), last_item=cfg["it"] * np.float32(0.9), user_tab=cfg["ut"].copy(), item_tab=cfg["it"].copy())
    for k, ref in (("m_user", "user_tab"), ("v_user", "user_tab"), ("m_item", "item_tab"), ("v_item", "item_tab")):
        st[k] = np.zeros_like(st[ref])
    B, n = cfg["B"], cfg["n"]
    total = 0.0
    for step, o in enumerate(range(0, n, B), 1):                      # the global batch of a step = both ranks' slices
        cat = lambda x: np.concatenate([x[0][o:o + B], x[1][o:o + B]])
        gu, gi, gj = cat(cfg["u"]), cat(cfg["i"]), cat(cfg["j"])
        l2_term = cfg["l2"] * 0.5 * float((st["user_tab"][gu] ** 2).sum() + (st["item_tab"][gi] ** 2).sum() + (st["item_tab"][gj] ** 2).sum())
        loss, _, _ = O.sml_mf_step(st, cfg["tu"], cfg["ti"], gu, gi, gj, cfg["lr"], cfg["l2"], step)
        total += float(loss) - l2_term                                # the sharded step reports the BCE part (like run_MF)
    parts = [torch.load(os.path.join(str(tmp_path), "s%d.pt" % r)) for r in range(world)]
    for r, p in enumerate(parts):
        assert np.abs(p["user"].numpy() - st["user_tab"][r::world]).max() < 2e-6, "user shard %d" % r
        assert np.abs(p["item"].numpy() - st["item_tab"][r::world]).max() < 2e-6, "item shard %d" % r
        assert np.abs(p["m_user"].numpy() - st["m_user"][r::world]).max() < 1e-7
        assert int(p["stamp"].min()) == 3 and int(p["stamp"].max()) == 3
    # each rank returns the sum over steps of its share of the global-mean BCE loss; the shares add up
    assert abs(sum(p["loss"] for p in parts) - total) < 1e-4 * abs(total)
    # transfer epoch: the oracle's dense single-process steps on the same global batches, snapshots = the tables after the MF epoch
    tu = {k: v.copy() for k, v in cfg["tu"].items()}; ti = {k: v.copy() for k, v in cfg["ti"].items()}
    opt = {nm: {k: (np.zeros_like(th[k]), np.zeros_like(th[k])) for k in O.THETA_KEYS} for nm, th in (("user", tu), ("item", ti))}
    tabs = dict(last_user=st["last_user"], last_item=st["last_item"], user_hat=st["user_tab"], item_hat=st["item_tab"])
    tr_total = 0.0
    for step, o in enumerate(range(0, n, B), 1):
        cat = lambda x: np.concatenate([x[0][o:o + B], x[1][o:o + B]])
        loss, _, _ = O.sml_tr_step(tu, ti, opt, tabs, cat(cfg["u"]), cat(cfg["i"]), cat(cfg["j"]), 0.001, 1e-4, step)
        tr_total += float(loss)
    from tests import fake_ops
    ref_flat, _ = fake_ops.flat_theta(tu, ti)
    for p in parts:
        assert (p["theta"] - ref_flat).abs().max().item() < 2e-5      # Adam on theta: lr 1e-3, noise-sensitive where |g| ~ eps
    assert torch.equal(parts[0]["theta"], parts[1]["theta"])           # replicas stay identical
    assert abs(sum(p["tr_loss"] for p in parts) - tr_total) < 1e-4 * abs(tr_total)


# ---------------------------------------------------------------------------------------------------------
# sharded full-catalog evaluation (padding, positive-row mapping, count all-reduce, per-rank slices) on CPU
def _fullcat_case():
    rng = np.random.default_rng(33)
    U, I = 41, 203
    it = rng.standard_normal((I, 64)).astype(np.float32)
    pairs = [np.stack([rng.integers(0, U, n), rng.integers(0, I, n)], 1) for n in (23, 9)]     # ragged: 23 pairs vs 9
    ut = rng.standard_normal((U, 64)).astype(np.float32) * 0.2
    for p in pairs:                                   # users look like their positives: small ranks, many hits
        ut[p[:, 0]] += it[p[:, 1]]
    return ut, it, pairs


def _fullcat_worker(rank, world, port, out_dir):
    import types
    from sml_b200.shard import ShardedSML
    from tests import fake_ops
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ut, it, pairs = _fullcat_case()
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        tr = types.SimpleNamespace(theta=torch.zeros(8), theta_grad=torch.zeros(8), variant=0)
        s = ShardedSML(shard_rows(T(ut), world, rank), shard_rows(T(it), world, rank), tr, world=world, rank=rank, ops=fake_ops)
        out = s.eval_fullcat(T(pairs[rank]), 5, chunk=8)             # several chunks, the last ones empty on rank 1
        torch.save(out, os.path.join(out_dir, "f%d.pt" % rank))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_fullcat_world2_matches_bruteforce(tmp_path):
    world = 2
    mp.spawn(_fullcat_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ut, it, pairs = _fullcat_case()
    allp = np.concatenate(pairs)
    s = ut[allp[:, 0]] @ it.T
    sp = (ut[allp[:, 0]] * it[allp[:, 1]]).sum(-1)
    mask = np.ones_like(s, dtype=bool); mask[np.arange(len(allp)), allp[:, 1]] = False
    rank = ((s >= sp[:, None]) & mask).sum(1)                        # ties: the positive loses
    hits = int((rank < 5).sum()); ndcg = float((1.0 / np.log2(rank[rank < 5] + 2.0)).sum())
    assert hits > 10                                                  # the case is built to have many hits
    for r in range(world):
        out = torch.load(os.path.join(str(tmp_path), "f%d.pt" % r))
        assert int(out[0]) == hits and int(out[2]) == len(allp) and abs(float(out[1]) - ndcg) < 1e-4
