"""GPU: full-catalog rank counts (tcgen05 score GEMM + compare-and-count epilogue) against numpy fp64 scores and
against the candidate-list kernel when the candidate list IS the whole catalog."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_users,n_items", [(1, 7), (130, 1000), (300, 5000 + 37), (257, 128 * 40)])
def test_fullcat_rank_counts(n_users, n_items):
    from sml_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(n_users + n_items)
    U = rng.standard_normal((400, 64)).astype(np.float32); I = rng.standard_normal((n_items, 64)).astype(np.float32)
    users = rng.integers(0, 400, n_users).astype(np.int64); pos = rng.integers(0, n_items, n_items and n_users).astype(np.int64)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    gt, eq = ops.fullcat_ranks(T(U), T(I), T(users), T(pos))
    s = U[users].astype(np.float64) @ I.astype(np.float64).T
    sp = s[np.arange(n_users), pos]
    mask = np.ones_like(s, dtype=bool); mask[np.arange(n_users), pos] = False
    ref_gt = ((s > sp[:, None]) & mask).sum(1)
    margin = np.where(mask, np.abs(s - sp[:, None]), np.inf).min(1)
    safe = margin > 1e-4                                              # 3xTF32 vs fp32 FFMA: a few 1e-6 relative
    assert safe.mean() > 0.9
    assert np.array_equal(gt.cpu().numpy()[safe], ref_gt[safe])
    assert np.abs(gt.cpu().numpy() - ref_gt).max() <= 2
    assert int(eq.sum()) <= int((~safe).sum())
    if n_items <= 1000:
        # the candidate-list kernel over [pos, every other item] must agree
        rows = np.stack([np.concatenate([[u, p], np.delete(np.arange(n_items), p)]) for u, p in zip(users, pos)]).astype(np.int64)
        g2, e2 = ops.eval_candidates(T(U), T(I), T(rows))
        assert np.array_equal(g2.cpu().numpy()[safe], gt.cpu().numpy()[safe])


def test_fullcat_item_shards_add_up():
    from sml_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    U = rng.standard_normal((200, 64)).astype(np.float32); I = rng.standard_normal((3000, 64)).astype(np.float32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    users = T(rng.integers(0, 200, 150).astype(np.int64)); pos = T(rng.integers(0, 3000, 150).astype(np.int64))
    gt, eq = ops.fullcat_ranks(T(U), T(I), users, pos)
    # two contiguous item shards, as two ranks would hold them
    It = T(I)
    g2 = torch.zeros_like(gt); e2 = torch.zeros_like(eq)
    for lo, hi in ((0, 1700), (1700, 3000)):
        ops.fullcat_ranks(T(U), It, users, pos, items_packed=ops.pack_rows(It[lo:hi].contiguous()), item_id0=lo, n_items=hi - lo, gt=g2, eq=e2)
    assert torch.equal(gt, g2) and torch.equal(eq, e2)
