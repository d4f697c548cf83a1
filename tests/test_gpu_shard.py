"""GPU: the sharded step path (exchange + pitch-128 views + owner-side scatter/Adam) at world_size 1 must
reproduce the fused single-GPU steps; the multi-rank equivalence is checked by tools/mgpu_sharded_check.py
under torchrun (NCCL)."""
import numpy as np
import pytest
import torch

from oracle import sml_oracle as O

pytestmark = pytest.mark.gpu


def test_sharded_steps_world1_match_fused_steps():
    from sml_b200 import ops
    from sml_b200.shard import ShardedSML
    from tests.test_gpu_parity import make_module, T
    from sml_b200.model.conv_transfer import ConvTransfer_com
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    U, I, B = 300, 500, 200                                    # B not a multiple of 128: padded row layout
    ut = rng.standard_normal((U, 64)).astype(np.float32); it = rng.standard_normal((I, 64)).astype(np.float32)
    tu, ti = O.init_theta(np.random.default_rng(1)), O.init_theta(np.random.default_rng(2))
    ids = [rng.integers(0, n, B).astype(np.int64) for n in (40, 60, 60)]     # duplicates
    # fused single-GPU reference
    m1 = make_module(ConvTransfer_com, tu, ti, dev)
    u1, i1 = T(ut, dev), T(it, dev)
    z = {k: torch.zeros_like(u1 if "user" in k else i1) for k in ("m_user", "v_user", "m_item", "v_item", "g_user", "g_item")}
    loss = torch.zeros(2, device=dev)
    a = ops.make_step_args(user=T(ids[0], dev), item=T(ids[1], dev), neg=T(ids[2], dev), last_user=T(ut, dev), last_item=T(it, dev),
                           hat_user=u1, hat_item=i1, theta=m1.theta, adam_state=ops.new_adam_state(dev), lr=0.01, l2=1e-6,
                           loss_out=loss, **z)
    ops.mf_step(a)
    # sharded path, one rank
    m2 = make_module(ConvTransfer_com, tu, ti, dev)
    s = ShardedSML(T(ut, dev), T(it, dev), m2, world=1, rank=0, mf_lr=0.01, l2=1e-6, tr_lr=0.001, tr_l2=1e-4)
    l2 = s.mf_step(T(ids[0], dev), T(ids[1], dev), T(ids[2], dev))
    assert abs(float(l2) + 0.0 - (loss[0].item() - 1e-6 * 0.5 * float((ut[ids[0]] ** 2).sum() + (it[ids[1]] ** 2).sum() + (it[ids[2]] ** 2).sum()))) < 2e-5
    s.flush()
    # (Adam amplifies summation-order noise on elements with |g| <~ eps by lr / eps: up to ~1e-6 here)
    assert (s.user - u1).abs().max().item() < 1e-5 and (s.item - i1).abs().max().item() < 1e-5
    # transfer step
    s.save_hat()
    mm, vv = torch.zeros_like(m1.theta), torch.zeros_like(m1.theta)
    loss.zero_()
    a = ops.make_step_args(user=T(ids[0], dev), item=T(ids[1], dev), neg=T(ids[2], dev), last_user=T(ut, dev), last_item=T(it, dev),
                           hat_user=u1, hat_item=i1, theta=m1.theta, adam_state=ops.new_adam_state(dev), lr=0.001, l2=1e-4,
                           g_theta=m1.theta_grad, m_theta=mm, v_theta=vv, loss_out=loss)
    ops.tr_step(a)
    lt = s.tr_step(T(ids[0], dev), T(ids[1], dev), T(ids[2], dev))
    assert abs(float(lt) - loss[0].item()) < 2e-5
    assert (m2.theta - m1.theta).abs().max().item() < 1e-4        # after Adam: step tolerance (atomic summation order varies)
    # updata + evaluation on the shard
    s.updata()
    rows = T(np.concatenate([rng.integers(0, U, (64, 1)), rng.integers(0, I, (64, 30))], 1).astype(np.int64), dev)
    out = s.eval_candidates(rows, 5)
    gt, eq = ops.eval_candidates(s.user, s.item, rows)
    assert int(out[0]) == int(((gt + eq) < 5).sum()) and int(out[2]) == 64
    # full-catalog ranks through the sharded path == the single-GPU full-catalog kernel
    pairs = T(np.stack([rng.integers(0, U, 90), rng.integers(0, I, 90)], 1).astype(np.int64), dev)
    fc = s.eval_fullcat(pairs, 20, chunk=64)                     # two chunks
    g1, e1 = ops.fullcat_ranks(s.user, s.item, pairs[:, 0].contiguous(), pairs[:, 1].contiguous())
    h1, n1 = ops.eval_reduce(g1, e1, 20, batch=90)
    assert int(fc[0]) == int(h1.sum()) and int(fc[2]) == 90 and abs(float(fc[1]) - float(n1.sum())) < 1e-4


def test_sharded_epoch_api_matches_step_api():
    """mf_epoch / tr_epoch (exchange planned once per epoch) == the same steps planned one by one."""
    from sml_b200.shard import ShardedSML
    from tests.test_gpu_parity import make_module, T
    from sml_b200.model.conv_transfer import ConvTransfer_com
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    U, I, B, n = 400, 700, 96, 96 * 3 + 17
    ut = rng.standard_normal((U, 64)).astype(np.float32) * 0.3; it = rng.standard_normal((I, 64)).astype(np.float32) * 0.3
    tu, ti = O.init_theta(np.random.default_rng(1)), O.init_theta(np.random.default_rng(2))
    u, i, j = (T(rng.integers(0, k, n).astype(np.int64), dev) for k in (U, I, I))
    res = []
    for epoch_api in (False, True):
        s = ShardedSML(T(ut, dev), T(it, dev), make_module(ConvTransfer_com, tu, ti, dev), world=1, rank=0)
        s.user.add_(0.05); s.save_hat(); s.user.sub_(0.05)
        if epoch_api:
            lm = s.mf_epoch(u, i, j, B); lt = s.tr_epoch(u, i, j, B)
        else:
            lm = sum(s.mf_step(u[o:o + B], i[o:o + B], j[o:o + B]) for o in range(0, n, B))
            lt = sum(s.tr_step(u[o:o + B], i[o:o + B], j[o:o + B]) for o in range(0, n, B))
        s.flush()
        res.append((s.user.clone(), s.item.clone(), s.transfer.theta.clone(), float(lm), float(lt)))
    a, b = res
    du, di, dt = ((a[k] - b[k]).abs().max().item() for k in range(3))
    # same kernels, same inputs: the two runs differ only through the order of fp32 atomics (row-gradient scatter, split-K
    # slices, theta-gradient reductions).  Adam amplifies an absolute gradient noise d on elements with |g| <~ eps = 1e-8 by
    # lr / eps per step: tables (lr 0.01, d ~ 1e-12) up to ~1e-6 per step, theta (lr 0.001, d ~ 1e-10) up to ~1e-5 per step;
    # four steps each here
    assert du < 2e-5 and di < 2e-5, (du, di)
    assert dt < 1e-4, dt
    assert abs(a[3] - b[3]) < 1e-4 * abs(a[3]) and abs(a[4] - b[4]) < 1e-4 * abs(a[4]), (a[3], b[3], a[4], b[4])
