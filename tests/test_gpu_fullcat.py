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


def test_fullcat_counts_are_self_consistent_on_exact_ties():
    """The positive's score comes from the same tensor-core arithmetic as the catalog scores (sml_fullcat_pos_scores): copies of
    the positive's row elsewhere in the catalog must tie EXACTLY (eq counts them, gt does not), whatever the 3xTF32 rounding."""
    from sml_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(8)
    n_items, n_users = 4000, 300
    U = rng.standard_normal((n_users, 64)).astype(np.float32); I = rng.standard_normal((n_items, 64)).astype(np.float32)
    users = np.arange(n_users, dtype=np.int64); pos = rng.integers(0, n_items // 2, n_users).astype(np.int64)
    copies = n_items // 2 + np.arange(n_users)                       # item n/2 + u := the positive of user u, twice for every third user
    I[copies] = I[pos]
    I[copies[::3] + n_users] = I[pos[::3]]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    gt, eq = ops.fullcat_ranks(T(U), T(I), T(users), T(pos))
    s = U.astype(np.float64) @ I.astype(np.float64).T
    dup = np.array([(np.abs(I - I[p]).max(1) == 0).sum() - 1 for p in pos])          # exact copies of each positive (other users may share it)
    assert np.array_equal(eq.cpu().numpy() >= dup, np.ones(n_users, bool)) and int((eq.cpu().numpy() - dup).max()) <= 1
    sp = s[np.arange(n_users), pos]
    strictly = ((s > sp[:, None] + 1e-4)).sum(1)
    assert np.all(gt.cpu().numpy() >= strictly) and np.all(gt.cpu().numpy() <= ((s > sp[:, None] - 1e-4)).sum(1) - 1 - dup + 1)


@pytest.mark.parametrize("n_users,n_items,k", [(5, 7, 20), (130, 1000, 20), (300, 122816, 64), (1, 129, 1)])
def test_fullcat_topk_vs_fp64(n_users, n_items, k):
    """Full-catalog top-k (fused epilogue + merge) against a full fp64 sort."""
    from sml_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(n_users * 7 + n_items)
    U = rng.standard_normal((500, 64)).astype(np.float32); I = rng.standard_normal((n_items, 64)).astype(np.float32)
    users = rng.integers(0, 500, n_users).astype(np.int64)
    excl = rng.integers(0, n_items, n_users).astype(np.int64)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for exclude in (None, excl):
        sc, ids = ops.fullcat_topk(T(U), T(I), T(users), k, exclude=None if exclude is None else T(exclude))
        sc, ids = sc.cpu().numpy(), ids.cpu().numpy()
        s = U[users].astype(np.float64) @ I.astype(np.float64).T
        if exclude is not None:
            s[np.arange(n_users), exclude] = -np.inf
        kk = min(k, n_items - (exclude is not None))
        order = np.argsort(-s, axis=1, kind="stable")[:, :kk]
        ref = np.take_along_axis(s, order, 1)
        assert np.all(np.diff(sc[:, :kk], axis=1) <= 0)                                            # descending
        assert np.abs(sc[:, :kk] - ref).max() < 2e-5 * np.abs(ref).max()                           # the k best scores
        assert np.all(ids[:, kk:] == -1) and np.all(np.isneginf(sc[:, kk:]))                       # padding beyond the catalog
        got = np.take_along_axis(s, np.maximum(ids[:, :kk], 0), 1)                                 # ids really carry those scores
        assert np.abs(got - sc[:, :kk]).max() < 2e-5 * np.abs(ref).max()
        assert all(len(set(r[:kk].tolist())) == kk for r in ids)                                   # no item twice
        # ids agree wherever the k-th / (k+1)-th gap and the in-list gaps are safe
        gaps = np.abs(np.diff(np.take_along_axis(s, np.argsort(-s, axis=1, kind="stable")[:, :kk + 1], 1), axis=1)).min(1) if n_items > kk + 1 else np.ones(n_users)
        safe = gaps > 1e-4
        assert np.array_equal(ids[safe][:, :kk], order[safe])
        if exclude is not None:
            assert not np.any(ids[:, :kk] == exclude[:, None])


def test_fullcat_topk_five_million_items():
    """config 5 scale on one GPU: 32 users x 5 M items, K = 20, against numpy fp64."""
    from sml_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(4)
    n_items, n_users, k = 5_000_000, 32, 20
    I = torch.randn(n_items, 64, device=dev, generator=g); U = torch.randn(n_users, 64, device=dev, generator=g)
    users = torch.arange(n_users, device=dev)
    sc, ids = ops.fullcat_topk(U, I, users, k)
    s = U.cpu().numpy().astype(np.float64) @ I.cpu().numpy().astype(np.float64).T
    order = np.argsort(-s, axis=1)[:, :k]
    ref = np.take_along_axis(s, order, 1)
    assert np.abs(sc.cpu().numpy() - ref).max() < 2e-5 * np.abs(ref).max()
    agree = (ids.cpu().numpy() == order).mean()
    assert agree > 0.98, agree                                        # near-ties among 5 M scores may swap neighbours
    # and the rank kernel agrees with the list: the k-th item of a user has exactly k - 1 items above it
    kth = torch.from_numpy(order[:, -1].copy()).to(dev)
    gt, eq = ops.fullcat_ranks(U, I, users, kth)
    assert int((gt - (k - 1)).abs().max()) <= 1
