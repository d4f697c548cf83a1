"""GPU: size-independent properties at the BASELINE.json full sizes (Yelp shape: 59 082 users, 122 816 items, 75 000-row
files with 1 + 1000 candidates), where the numpy oracle is too slow to be the checker."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

U, I, N, C = 59082, 122816, 75000, 1000


@pytest.fixture(scope="module")
def world():
    from sml_b200 import ops
    from sml_b200.model.conv_transfer import ConvTransfer_com
    import contextlib, io
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    R = lambda *s: torch.randn(*s, generator=g).to(dev)
    with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
        tr = ConvTransfer_com(64, 64).to(dev)
    rows = torch.cat([torch.randint(0, U, (N, 1), generator=g), torch.randint(0, I, (N, C), generator=g)], 1).to(dev)
    return dict(ops=ops, dev=dev, tr=tr, ut=R(U, 64), it=R(I, 64), lu=R(U, 64), li=R(I, 64), rows=rows, g=g)


def test_eval_full_file_properties(world):
    ops, rows, ut, it = world["ops"], world["rows"], world["ut"], world["it"]
    gt, eq = ops.eval_candidates(ut, it, rows)
    assert int(gt.min()) >= 0 and int(gt.max()) <= C - 1
    # (1) permuting the negatives of a row does not change its rank
    perm = torch.randperm(C - 1, generator=world["g"]).to(world["dev"]) + 2
    rows2 = torch.cat([rows[:, :2], rows[:, perm]], 1).contiguous()
    gt2, eq2 = ops.eval_candidates(ut, it, rows2)
    assert torch.equal(gt, gt2) and torch.equal(eq, eq2)
    # (2) idempotent, and a strided view of a wider file gives the same counts (row_stride > 1 + C)
    wide = torch.cat([rows, torch.zeros(N, 7, dtype=torch.int64, device=world["dev"])], 1)
    from sml_b200._lib import lib, check, ptr, stream
    g3 = torch.empty_like(gt); e3 = torch.empty_like(eq)
    check(lib().sml_eval_candidates(ptr(ut), ptr(it), 64, ptr(wide), N, wide.stride(0), C, ptr(g3), ptr(e3), stream()))
    assert torch.equal(gt, g3)
    # (3) hits@K is monotone in K and equals the rank histogram; sum of per-batch hits = total
    h = [int(ops.eval_reduce(gt, eq, K)[0].sum()) for K in (5, 10, 20, 1000)]
    assert h[0] <= h[1] <= h[2] <= h[3] == N
    assert h[2] == int(((gt + eq) < 20).sum())
    # (4) boosting a positive's score can only improve its rank: copy the user row into the positive item row
    it2 = it.clone()
    sel = torch.arange(0, 2000, device=world["dev"])
    it2[rows[sel, 1]] = ut[rows[sel, 0]] * 3.0
    gt4, _ = ops.eval_candidates(ut, it2, rows[:2000].contiguous())
    assert int((gt4 <= gt[:2000]).float().mean() > 0.99)
    # (5) random rows agree with scores recomputed pair by pair (same fp32 kernel family)
    pick = torch.randint(0, N, (64,), generator=world["g"]).to(world["dev"])
    uu = rows[pick, :1].expand(-1, C).reshape(-1).contiguous(); ii = rows[pick, 1:].reshape(-1).contiguous()
    sc = ops.pair_scores(ut, it, uu, ii).reshape(64, C)
    assert torch.equal((sc[:, 1:] > sc[:, :1]).sum(1).int(), gt[pick])


def test_transfer_full_table_properties(world):
    ops, tr, lu, ut = world["ops"], world["tr"], world["lu"], world["ut"]
    th = tr.theta[:ops.NET_STRIDE]
    full = ops.transfer_forward(lu, ut, th)
    assert torch.isfinite(full).all()
    # chunk / tile independence: any sub-range computed alone is bit-identical to the same rows of the full pass
    for lo, hi in ((0, 1), (5000, 5000 + 129), (16384 - 3, 16384 + 200), (U - 77, U)):
        part = ops.transfer_forward(lu[lo:hi].contiguous(), ut[lo:hi].contiguous(), th)
        assert torch.equal(part, full[lo:hi]), (lo, hi)
    # gather path == direct path on a permutation (row-permutation equivariance)
    perm = torch.randperm(U, generator=world["g"]).to(world["dev"])
    assert torch.equal(ops.transfer_forward(lu, ut, th, ids=perm), full[perm])
    # tensor-core path vs SIMT path is covered at small sizes; here: user and item nets differ
    other = ops.transfer_forward(lu, ut, tr.theta[ops.NET_STRIDE:])
    assert not torch.equal(other, full)


def test_mf_step_full_tables_touches_only_batch_rows(world):
    """Dense Adam with all-zero state: after the first step exactly the rows of the batch move (model/transfer.py:392)."""
    ops, tr, dev = world["ops"], world["tr"], world["dev"]
    ut, it = world["ut"].clone(), world["it"].clone()
    z = torch.zeros_like
    st = dict(g_user=z(ut), g_item=z(it), m_user=z(ut), v_user=z(ut), m_item=z(it), v_item=z(it))
    B = 1024
    g = world["g"]
    u = torch.randint(0, U, (B,), generator=g).to(dev); i = torch.randint(0, I, (B,), generator=g).to(dev); j = torch.randint(0, I, (B,), generator=g).to(dev)
    loss = torch.zeros(2, device=dev)
    a = ops.make_step_args(user=u, item=i, neg=j, last_user=world["lu"], last_item=world["li"], hat_user=ut, hat_item=it, theta=tr.theta,
                           adam_state=ops.new_adam_state(dev), lr=0.01, l2=1e-6, loss_out=loss, **st)
    ops.mf_step(a)
    moved_u = (ut != world["ut"]).any(1); moved_i = (it != world["it"]).any(1)
    exp_u = torch.zeros(U, dtype=torch.bool, device=dev); exp_u[u] = True
    exp_i = torch.zeros(I, dtype=torch.bool, device=dev); exp_i[i] = True; exp_i[j] = True
    assert torch.equal(moved_u, exp_u) and torch.equal(moved_i, exp_i)
    assert float(st["g_user"].abs().max()) == 0.0 and float(st["g_item"].abs().max()) == 0.0
    # first Adam update: lr * g / (|g| + eps) -> never above lr, close to lr unless |g| is comparable to eps = 1e-8
    d = (ut - world["ut"])[exp_u].abs()
    assert float(d.max()) <= 0.01 * (1 + 1e-5) and float(d.median()) > 0.009
    assert torch.isfinite(loss).all() and loss[0].item() > 0


def test_fullcat_matches_candidate_kernel_at_full_catalog(world):
    ops, ut, it, dev = world["ops"], world["ut"], world["it"], world["dev"]
    g = world["g"]
    users = torch.randint(0, U, (300,), generator=g).to(dev); pos = torch.randint(0, I, (300,), generator=g).to(dev)
    gt, eq = ops.fullcat_ranks(ut, it, users, pos)
    # candidate kernel over [pos, 999 random other items]: its count is a lower bound of the full-catalog count
    neg = torch.randint(0, I, (300, 999), generator=g).to(dev)
    rows = torch.cat([users[:, None], pos[:, None], neg], 1).contiguous()
    g2, _ = ops.eval_candidates(ut, it, rows)
    assert bool((g2 <= gt + 2).all())
    # exact check on a few users with fp64 scores
    s = ut[users[:16]].double() @ it.double().t()
    sp = s[torch.arange(16), pos[:16]]
    ref = (s > sp[:, None]).sum(1)
    assert int((gt[:16].long() - ref).abs().max()) <= 2
