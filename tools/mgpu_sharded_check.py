"""torchrun --nproc-per-node N tools/mgpu_sharded_check.py : the N-rank sharded MF / transfer steps (NCCL
all-to-all row exchange, theta all-reduce) against the same global batch run on one GPU with the fused steps.
Also times sharded steps on scaled synthetic tables (rows per GPU fixed: weak scaling of the table size)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sml_oracle as O  # noqa: E402  (checker only)
from sml_b200 import ops  # noqa: E402
from sml_b200.model.conv_transfer import ConvTransfer_com  # noqa: E402
from sml_b200.shard import ShardedSML, shard_rows  # noqa: E402


def module(tu, ti, dev):
    import contextlib, io
    with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
        m = ConvTransfer_com(64, 64).to(dev)
    sd = {("user_transfer." + k): torch.from_numpy(v) for k, v in tu.items()}
    sd.update({("item_transfer." + k): torch.from_numpy(v) for k, v in ti.items()})
    m.load_state_dict(sd)
    return m


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rng = np.random.default_rng(0)
    U, I, B = 4000, 6000, 256 * world
    ut = rng.standard_normal((U, 64)).astype(np.float32); it = rng.standard_normal((I, 64)).astype(np.float32)
    tu, ti = O.init_theta(np.random.default_rng(1)), O.init_theta(np.random.default_rng(2))
    ids = [rng.integers(0, n, B).astype(np.int64) for n in (U, I, I)]
    res = {}
    # single-GPU fused reference on rank 0's device (every rank computes it: cheap)
    m1 = module(tu, ti, dev)
    u1, i1 = T(ut), T(it)
    z = {k: torch.zeros_like(u1 if "user" in k else i1) for k in ("m_user", "v_user", "m_item", "v_item", "g_user", "g_item")}
    loss = torch.zeros(2, device=dev)
    a = ops.make_step_args(user=T(ids[0]), item=T(ids[1]), neg=T(ids[2]), last_user=T(ut), last_item=T(it), hat_user=u1, hat_item=i1,
                           theta=m1.theta, adam_state=ops.new_adam_state(dev), lr=0.01, l2=1e-6, loss_out=loss, **z)
    ops.mf_step(a)
    m2 = module(tu, ti, dev)
    s = ShardedSML(shard_rows(T(ut), world, rank), shard_rows(T(it), world, rank), m2, world=world, rank=rank, mf_lr=0.01, l2=1e-6)
    sl = slice(rank * (B // world), (rank + 1) * (B // world))
    # full-catalog ranks of 512 pairs (split over the ranks) on the untouched tables: exact integer agreement expected
    pairs = np.stack([rng.integers(0, U, 512), rng.integers(0, I, 512)], 1).astype(np.int64)
    g1, e1 = ops.fullcat_ranks(T(ut), T(it), T(pairs[:, 0]), T(pairs[:, 1]))
    h1, n1 = ops.eval_reduce(g1, e1, 20, batch=512)
    psl = slice(rank * 512 // world, (rank + 1) * 512 // world)
    fc = s.eval_fullcat(T(pairs[psl]), 20, chunk=100)
    res["fullcat_hits"] = [int(fc[0]), int(h1.sum())]
    res["fullcat_ndcg_diff"] = abs(float(fc[1]) - float(n1.sum()))
    s.mf_step(T(ids[0][sl]), T(ids[1][sl]), T(ids[2][sl]))
    s.flush()
    res["mf_user_maxdiff"] = float((s.user - u1[rank::world]).abs().max())
    res["mf_item_maxdiff"] = float((s.item - i1[rank::world]).abs().max())
    s.save_hat()
    mm, vv = torch.zeros_like(m1.theta), torch.zeros_like(m1.theta)
    a = ops.make_step_args(user=T(ids[0]), item=T(ids[1]), neg=T(ids[2]), last_user=T(ut), last_item=T(it), hat_user=u1, hat_item=i1,
                           theta=m1.theta, adam_state=ops.new_adam_state(dev), lr=0.001, l2=1e-4, g_theta=m1.theta_grad, m_theta=mm,
                           v_theta=vv, loss_out=loss)
    ops.tr_step(a)
    s.tr_step(T(ids[0][sl]), T(ids[1][sl]), T(ids[2][sl]))
    res["tr_theta_maxdiff"] = float((m2.theta - m1.theta).abs().max())
    ok = res["fullcat_hits"][0] == res["fullcat_hits"][1] and int(fc[2]) == 512 and res["fullcat_ndcg_diff"] < 1e-3 and \
        res["mf_user_maxdiff"] < 1e-5 and res["mf_item_maxdiff"] < 1e-5 and res["tr_theta_maxdiff"] < 1e-4   # theta after Adam: step tolerance (1e-4)
    # ---- throughput on scaled tables: rows per GPU fixed ----
    Ul, Il, Bl = int(os.environ.get("SML_ROWS_PER_GPU", 4_000_000)), int(os.environ.get("SML_ITEMS_PER_GPU", 1_000_000)), 8192
    g = torch.Generator(device=dev).manual_seed(rank)
    big = ShardedSML(torch.randn(Ul, 64, device=dev, generator=g), torch.randn(Il, 64, device=dev, generator=g), m2, world=world, rank=rank)
    gu = lambda n, hi: torch.randint(0, hi, (n,), device=dev, generator=g)
    steps = 12
    for kind, fn in (("mf", big.mf_epoch), ("tr", big.tr_epoch)):
        # whole epochs through the epoch API: exchange planned once (one host sync), steps enqueue asynchronously;
        # the timed region includes the planning
        tri = lambda: (gu(Bl * steps, Ul * world), gu(Bl * steps, Il * world), gu(Bl * steps, Il * world))
        fn(*tri(), Bl)
        torch.cuda.synchronize(); dist.barrier()
        u_, i_, j_ = tri()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(u_, i_, j_, Bl)
        if kind == "mf":
            big.flush()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["%s_step_ms" % kind] = float(t)
        res["%s_triples_per_s" % kind] = Bl * world / float(t) * 1e3
    # full-catalog evaluation: 16384 pairs per GPU against all Il * world items
    npairs = 16384
    pr = torch.stack([gu(npairs, Ul * world), gu(npairs, Il * world)], 1)
    big.eval_fullcat(pr, 20)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); big.eval_fullcat(pr, 20); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["fullcat_ms"] = float(t)
    res["fullcat_pairs_per_s"] = npairs * world / float(t) * 1e3
    res["fullcat_tflops"] = 2.0 * 64 * npairs * world * Il * world / (float(t) * 1e-3) / 1e12
    res.update(world=world, ok=bool(ok), rows_per_gpu=Ul + Il, batch_per_gpu=Bl)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
