"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures generated from the
UNMODIFIED reference and against the numpy oracle on seeded inputs.

Bars (BASELINE.json north_star): integer / index outputs (ranks, hit counts, hit rows, sampled
triples) bit-exact; fp32 forward quantities rel 1e-5; after optimizer updates 1e-4."""
import numpy as np
import pytest
import torch

from oracle import sml_oracle as O
from tests.helpers import theta_from_chk, sample, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 5e-5
STEP_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from sml_b200 import _lib
    assert _lib.lib().sml_device_check() == 0, _lib.lib().sml_last_error()
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def make_module(cls, tu, ti, dev):
    with torch.random.fork_rng(devices=[]):
        m = cls(64, 64).to(dev)
    sd = {("user_transfer." + k): torch.from_numpy(v) for k, v in tu.items()}
    sd.update({("item_transfer." + k): torch.from_numpy(v) for k, v in ti.items()})
    m.load_state_dict(sd)
    return m


def theta_np(m, net):
    return {k: getattr(getattr(getattr(m, net + "_transfer"), k.split(".")[0]), k.split(".")[1]).detach().cpu().numpy() for k in O.THETA_KEYS}


# ------------------------------------------------------------------------------- evaluation
def test_eval_golden(golden, dev):
    from sml_b200.model.MF import MFbasemode
    from sml_b200.evalution.evaluation2 import test_model, DeviceTestSet
    g = golden("eval")
    with torch.random.fork_rng(devices=[]):
        mf = MFbasemode(30, 200, 64).to(dev)
    mf.user_laten.weight.data.copy_(T(g["user"], dev)); mf.item_laten.weight.data.copy_(T(g["item"], dev))
    rows = T(g["rows"], dev)
    for K in (20, 10, 5):
        h, nd, idx = mf.test(rows, topK=K)
        assert h == float(g["hits@%d" % K])                               # exact
        assert np.array_equal(idx.cpu().numpy(), g["idx@%d" % K])         # exact
        assert abs(float(nd) - float(g["ndcg@%d" % K])) < 1e-5
        r, n = test_model(mf, DeviceTestSet(rows, batch=16), topK=K)
        assert abs(r - float(g["recall@%d" % K])) < 1e-12 and abs(float(n) - float(g["tm_ndcg@%d" % K])) < 1e-6
        loader = [rows[s:s + 16] for s in range(0, rows.shape[0], 16)]   # the reference's calling convention
        r2, n2 = test_model(mf, loader, topK=K)
        assert abs(r2 - r) < 1e-12 and abs(float(n2) - float(n)) < 1e-6
    # forward scores
    u = rows[:, :1].expand(-1, rows.shape[1] - 1).reshape(-1).contiguous()
    i = rows[:, 1:].reshape(-1).contiguous()
    _, _, sc = mf(u, i)
    assert rel_err(sc.cpu().numpy().reshape(g["scores"].shape), g["scores"]) < FWD_TOL


def test_legacy_eval_golden(golden, dev):
    """evalution/evaluation.py:7-60 (per-user lists, positives first) through the CUDA pair-score kernel against the
    reference's own output on the same tables (tests/golden/eval.npz: legacy)."""
    from sml_b200.model.MF import MFbasemode
    from sml_b200.evalution import evaluation
    g = golden("eval")
    with torch.random.fork_rng(devices=[]):
        mf = MFbasemode(30, 200, 64).to(dev)
    mf.user_laten.weight.data.copy_(T(g["user"], dev)); mf.item_laten.weight.data.copy_(T(g["legacy_item"], dev))
    users, pos, negs = [1, 2], [[3, 4], [5]], [list(range(10, 40)), list(range(50, 80))]
    res = [float(x) for x in evaluation.test_model(mf, (users, pos, negs), topK=5)]
    ref = g["legacy"]
    assert res[0] == 0 and res[2] == 0 and res[4] == 0 and ref[0] == 0 and ref[2] == 0 and ref[4] == 0
    assert abs(res[1] - ref[1]) < 1e-7 and abs(res[3] - ref[3]) < 1e-6, (res, ref)
    # and against the oracle's restatement user by user, at a K that cuts into the positives
    for u, p, n in zip(users, pos, negs):
        s = g["legacy_item"][np.array(p + n)] @ g["user"][u]
        for K in (1, 3, 5, 10):
            ro, do = O.legacy_rec_ndcg(s, len(p), K)
            _, rg, _, dg, _ = evaluation.evalution_for_user(mf, u, p, n, K)
            assert abs(float(rg) - ro) < 1e-7 and abs(float(dg) - do) < 1e-6, (u, K)


def test_eval_vs_oracle_c1000_and_edges(dev):
    from sml_b200 import ops
    rng = np.random.default_rng(1)
    U, I, N, C = 400, 3000, 777, 1000                                    # ragged N, the reference's C
    ut = rng.standard_normal((U, 64)).astype(np.float32); it = rng.standard_normal((I, 64)).astype(np.float32)
    rows = np.concatenate([rng.integers(0, U, (N, 1)), rng.integers(0, I, (N, C))], 1).astype(np.int64)
    rows[5, 2:6] = rows[5, 1]                                             # exact ties with the positive
    it[7] = np.nan; rows[9, 3] = 7                                        # a NaN candidate outranks everything
    rows[11, 1] = 7                                                       # NaN positive
    gt, eq = ops.eval_candidates(T(ut, dev), T(it, dev), T(rows, dev))
    s64 = O.candidate_scores(ut.astype(np.float64), it.astype(np.float64), rows)
    ogt, oeq = O.candidate_ranks(s64.astype(np.float32))
    # rows whose positive is separated from every negative by > 1e-4 must be bit-exact (summation order only
    # matters inside a few ulp); count the rest
    with np.errstate(invalid="ignore"):
        margin = np.nanmin(np.where(s64[:, 1:] == s64[:, :1], np.inf, np.abs(s64[:, 1:] - s64[:, :1])), axis=1)
    safe = margin > 1e-4
    assert safe.mean() > 0.95
    assert np.array_equal(gt.cpu().numpy()[safe], ogt[safe])
    assert int(eq[5]) >= 4 and int(gt[9]) >= 1
    assert int(gt[11]) == 0                                               # nothing beats a NaN positive
    # empty input and C = 1
    g0, e0 = ops.eval_candidates(T(ut, dev), T(it, dev), torch.zeros(0, 1 + C, dtype=torch.int64, device=dev))
    assert g0.numel() == 0
    g1, e1 = ops.eval_candidates(T(ut, dev), T(it, dev), T(rows[:, :2].copy(), dev))
    assert int(g1.sum()) == 0 and int(e1.sum()) == 0
    hits, nd = ops.eval_reduce(gt, eq, 20, batch=1024)
    r = (ogt + oeq)
    assert abs(int(hits.sum()) - int((r < 20).sum())) <= int((~safe).sum())


def test_eval_rejects_bad_args(dev):
    from sml_b200 import ops
    t = torch.zeros(4, 32, device=dev)
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.eval_candidates(t, t, torch.zeros(2, 3, dtype=torch.int64, device=dev))
    with pytest.raises(RuntimeError):
        ops.eval_candidates(torch.zeros(4, 64), torch.zeros(4, 64), torch.zeros(2, 3, dtype=torch.int64))   # CPU tensors


# ------------------------------------------------------------------------------- transfer forward
def test_transfer_forward_golden(golden, dev):
    from sml_b200.model.conv_transfer import ConvTransfer_com, ConvTransfer
    g = golden("transfer_fwd")
    x_t, x_hat = T(g["x_t"], dev), T(g["x_hat"], dev)
    com = make_module(ConvTransfer_com, *theta_from_chk(g["theta_com"]), dev)
    assert rel_err(com(x_t, x_hat, "user").cpu().numpy(), g["com_user"]) < FWD_TOL
    assert rel_err(com(x_t, x_hat, "item").cpu().numpy(), g["com_item"]) < FWD_TOL
    conv = make_module(ConvTransfer, *theta_from_chk(g["theta_conv"]), dev)
    assert rel_err(conv(x_t, x_hat, "user").cpu().numpy(), g["conv_user"]) < FWD_TOL
    assert rel_err(conv(x_t, x_hat, "item").cpu().numpy(), g["conv_item"]) < FWD_TOL
    with pytest.raises(TypeError):
        com(x_t, x_hat, "nobody")


@pytest.mark.parametrize("n", [1, 63, 129, 8192 + 77])
def test_transfer_forward_vs_oracle(dev, n):
    from sml_b200 import ops
    rng = np.random.default_rng(n)
    th = O.init_theta(np.random.default_rng(3))
    flat = flat_theta(th, dev)
    xt = rng.standard_normal((n, 64)).astype(np.float32); xh = (0.5 * rng.standard_normal((n, 64))).astype(np.float32)
    y = ops.transfer_forward(T(xt, dev), T(xh, dev), flat).cpu().numpy()
    assert rel_err(y, O.conv_transfer_com_forward(th, xt, xh)) < FWD_TOL
    ids = rng.integers(0, n, size=max(1, n // 2)).astype(np.int64)
    y = ops.transfer_forward(T(xt, dev), T(xh, dev), flat, ids=T(ids, dev)).cpu().numpy()
    assert rel_err(y, O.conv_transfer_com_forward(th, xt[ids], xh[ids])) < FWD_TOL


def flat_theta(th, dev):
    from sml_b200 import _lib
    f = torch.zeros(_lib.NET_STRIDE, dtype=torch.float32)
    for k, off in (("conv1.weight", _lib.OFF_C1W), ("conv1.bias", _lib.OFF_C1B), ("conv2.weight", _lib.OFF_C2W), ("conv2.bias", _lib.OFF_C2B),
                   ("fc1.weight", _lib.OFF_F1W), ("fc1.bias", _lib.OFF_F1B), ("fc2.weight", _lib.OFF_F2W), ("fc2.bias", _lib.OFF_F2B)):
        f[off:off + th[k].size] = torch.from_numpy(th[k].reshape(-1))
    return f.to(dev)


def test_transfer_zero_row_is_nan_like_reference(dev):
    """x_com = x_t*x_hat/||x_t|| has no eps (conv_transfer.py:94-98): an all-zero w_{t-1} row gives NaN."""
    from sml_b200 import ops
    th = O.init_theta(np.random.default_rng(3))
    xt = np.zeros((3, 64), np.float32); xt[1] = 1.0
    xh = np.ones((3, 64), np.float32)
    y = ops.transfer_forward(T(xt, dev), T(xh, dev), flat_theta(th, dev)).cpu().numpy()
    assert np.isnan(y[0]).all() and np.isfinite(y[1]).all() and np.isnan(y[2]).all()


# ------------------------------------------------------------------------------- run_MF + gradients
@pytest.mark.parametrize("tag", ["com_bce", "com_bpr", "conv_bpr"])
def test_run_mf_golden(golden, dev, tag):
    from sml_b200.model.conv_transfer import ConvTransfer_com, ConvTransfer
    g = golden("run_mf")
    cls = ConvTransfer if tag == "conv_bpr" else ConvTransfer_com
    m = make_module(cls, *theta_from_chk(g["theta_conv" if tag == "conv_bpr" else "theta_com"]), dev)
    rows = {k: T(g[k], dev).requires_grad_("hat" in k) for k in ("u_last", "u_hat", "i_last", "i_hat", "j_last", "j_hat")}
    kw = {} if tag == "conv_bpr" else dict(BCE=(tag == "com_bce"))
    loss = m.run_MF(rows["u_last"], rows["u_hat"], rows["i_last"], rows["i_hat"], rows["j_last"], rows["j_hat"], **kw)
    loss.backward()
    ref = float(g[tag + ".loss"])
    assert abs(loss.item() - ref) < FWD_TOL * max(1.0, abs(ref))
    for k in ("u_hat", "i_hat", "j_hat"):
        assert rel_err(rows[k].grad.cpu().numpy(), g[tag + ".d_" + k]) < GRAD_TOL, k
    for net in ("user", "item"):
        mod = getattr(m, net + "_transfer")
        for k in O.THETA_KEYS:
            a, b = k.split(".")
            got = getattr(getattr(mod, a), b).grad.cpu().numpy()
            ref = g["%s.g_%s.%s" % (tag, net, k)]
            assert rel_err(sample(got).reshape(ref.shape), ref) < GRAD_TOL, (net, k)
    # no autograd requested -> loss only
    with torch.no_grad():
        l2 = m.run_MF(*(rows[k].detach() for k in ("u_last", "u_hat", "i_last", "i_hat", "j_last", "j_hat")), **kw)
    assert abs(l2.item() - loss.item()) < 1e-6


# ------------------------------------------------------------------------------- the two hot loops
def test_mf_steps_golden(golden, dev):
    """3 MF steps with duplicate ids in a batch; dense Adam: rows touched earlier keep moving."""
    from sml_b200 import ops
    from sml_b200.model.conv_transfer import ConvTransfer_com
    g = golden("mf_steps")
    m = make_module(ConvTransfer_com, *theta_from_chk(g["theta_com"]), dev)
    ut, it = T(g["user0"], dev), T(g["item0"], dev)
    z = {k: torch.zeros_like(ut if "user" in k else it) for k in ("m_user", "v_user", "m_item", "v_item", "g_user", "g_item")}
    loss = torch.zeros(2, device=dev)
    state = ops.new_adam_state(dev)
    lu, li = T(g["last_user"], dev), T(g["last_item"], dev)
    for s in range(3):
        u, i, j = (T(g["ids"][s, k], dev) for k in range(3))
        a = ops.make_step_args(user=u, item=i, neg=j, last_user=lu, last_item=li, hat_user=ut, hat_item=it, theta=m.theta,
                               adam_state=state, lr=float(g["lr"]), l2=float(g["l2"]), loss_out=loss, **z)
        ops.mf_step(a)
        assert abs(loss[0].item() - float(g["loss%d" % s])) < 2e-5
        assert np.abs(ut.cpu().numpy() - g["user%d" % (s + 1)]).max() < STEP_TOL
        assert np.abs(it.cpu().numpy() - g["item%d" % (s + 1)]).max() < STEP_TOL
        assert float(z["g_user"].abs().max()) == 0.0 and float(z["g_item"].abs().max()) == 0.0   # re-zeroed
    assert int(state[0]) == 3
    assert abs(loss[1].item() - sum(float(g["loss%d" % s]) for s in range(3))) < 1e-4
    assert rel_err(z["m_user"].cpu().numpy(), g["m_user"]) < 1e-4 and rel_err(z["v_item"].cpu().numpy(), g["v_item"]) < 1e-4
    untouched = np.setdiff1d(np.arange(50), np.unique(g["ids"][:, 0]))
    assert np.array_equal(ut.cpu().numpy()[untouched], g["user0"][untouched])                     # exact


def test_tr_steps_golden(golden, dev):
    from sml_b200 import ops
    from sml_b200.model.conv_transfer import ConvTransfer_com
    g = golden("tr_steps")
    m = make_module(ConvTransfer_com, *theta_from_chk(g["theta_com"]), dev)
    tabs = {k: T(g[k], dev) for k in ("last_user", "user_hat", "last_item", "item_hat")}
    mm, vv = torch.zeros_like(m.theta), torch.zeros_like(m.theta)
    loss = torch.zeros(2, device=dev)
    state = ops.new_adam_state(dev)
    for s in range(3):
        u, i, j = (T(g["ids"][s, k], dev) for k in range(3))
        a = ops.make_step_args(user=u, item=i, neg=j, last_user=tabs["last_user"], last_item=tabs["last_item"],
                               hat_user=tabs["user_hat"], hat_item=tabs["item_hat"], theta=m.theta, adam_state=state,
                               lr=float(g["lr"]), l2=float(g["wd"]), g_theta=m.theta_grad, m_theta=mm, v_theta=vv, loss_out=loss)
        ops.tr_step(a)
        assert abs(loss[0].item() - float(g["loss%d" % s])) < 2e-5
    for net in ("user", "item"):
        th = theta_np(m, net)
        for k in O.THETA_KEYS:
            ref = g["t3.%s.%s" % (net, k)]
            assert np.abs(sample(th[k]).reshape(ref.shape) - ref).max() < STEP_TOL, (net, k)
    assert float(m.theta_grad.abs().max()) == 0.0
    # snapshot tables are read-only in the transfer step
    assert np.array_equal(tabs["user_hat"].cpu().numpy(), g["user_hat"])


def test_plain_mf_vs_oracle(dev):
    from sml_b200 import ops
    rng = np.random.default_rng(5)
    U, I, B = 200, 300, 257
    ut = (0.3 * rng.standard_normal((U, 64))).astype(np.float32); it = (0.3 * rng.standard_normal((I, 64))).astype(np.float32)
    ib = rng.standard_normal((I, 1)).astype(np.float32); ub = rng.standard_normal((U, 1)).astype(np.float32)
    u, i, j = (rng.integers(0, n, B).astype(np.int64) for n in (40, 60, 60))
    lo, gu, gi = O.plain_mf_bce_grads(ut, it, u, i, j, 1e-3, 2e-3)
    g_u, g_i = torch.zeros(U, 64, device=dev), torch.zeros(I, 64, device=dev)
    loss = torch.zeros(2, device=dev)
    ops.plain_mf_grads(T(ut, dev), T(it, dev), T(u, dev), T(i, dev), T(j, dev), g_u, g_i, loss, loss=ops.LOSS_BCE, l2_u=1e-3, l2_i=2e-3)
    assert abs(loss[0].item() - float(lo)) < 2e-5
    assert rel_err(g_u.cpu().numpy(), gu) < GRAD_TOL and rel_err(g_i.cpu().numpy(), gi) < GRAD_TOL
    lo, gu, gi, gb = O.plain_mf_bpr_grads(ut, it, ub, ib, u, i, j)
    g_u.zero_(); g_i.zero_(); g_b = torch.zeros(I, device=dev); loss.zero_()
    ops.plain_mf_grads(T(ut, dev), T(it, dev), T(u, dev), T(i, dev), T(j, dev), g_u, g_i, loss, loss=ops.LOSS_BPR,
                       item_bias=T(ib[:, 0].copy(), dev), g_item_bias=g_b)
    assert abs(loss[0].item() - float(lo)) < 1e-4 * max(1.0, abs(float(lo)))
    assert rel_err(g_u.cpu().numpy(), gu) < GRAD_TOL and rel_err(g_i.cpu().numpy(), gi) < GRAD_TOL
    assert rel_err(g_b.cpu().numpy(), gb[:, 0]) < GRAD_TOL
    # dense Adam on top (fused zero_grad)
    m, v = torch.zeros_like(g_u), torch.zeros_like(g_u)
    p = T(ut, dev)
    st = ops.new_adam_state(dev)
    ops.adam_tick(st, 0.01)
    ops.adam_dense(p, m, v, g_u, st)
    pn, mn, vn = ut.copy(), np.zeros_like(ut), np.zeros_like(ut)
    O.adam_step(pn, gu, mn, vn, 1, 0.01)
    assert np.abs(p.cpu().numpy() - pn).max() < 1e-5 and float(g_u.abs().max()) == 0.0


@pytest.mark.parametrize("loss_kind", ["bce", "bpr"])
def test_plain_mf_fused_step_vs_oracle(dev, loss_kind):
    """sml_plain_mf_step (gather - dot - loss - row gradients + L2 - Adam in one kernel) against the oracle's
    plain-MF gradients followed by DENSE Adam on whole tables (model/baseline.py:111,188-201), over several steps with
    duplicate ids inside a batch, rows that skip steps (row-lazy replay) and rows never touched."""
    from sml_b200 import ops
    rng = np.random.default_rng(11)
    U, I, B, steps = 300, 400, 333, 9
    ut = (0.3 * rng.standard_normal((U, 64))).astype(np.float32); it = (0.3 * rng.standard_normal((I, 64))).astype(np.float32)
    l2u, l2i, lr = (1e-3, 2e-3, 0.01) if loss_kind == "bce" else (0.0, 0.0, 0.01)
    pu, pi = T(ut, dev), T(it, dev)
    z = torch.zeros_like
    mu, vu, mi, vi = z(pu), z(pu), z(pi), z(pi)
    st = ops.new_adam_state(dev, history=True)
    su, si = ops.new_row_stamps(U, st), ops.new_row_stamps(I, st)
    hu, hi = ops.new_list_heads(U, dev), ops.new_list_heads(I, dev)
    loss = torch.zeros(2, device=dev)
    nu, ni = ut.copy(), it.copy()
    nmu, nvu, nmi, nvi = (np.zeros_like(a) for a in (ut, ut, it, it))
    zb_u, zb_i = np.zeros((U, 1), np.float32), np.zeros((I, 1), np.float32)
    for s in range(steps):
        hot = 30 if s % 2 == 0 else 250                     # small id ranges: many duplicates; rows come and go
        u, i, j = (rng.integers(0, n, B).astype(np.int64) for n in (hot, hot + 20, hot + 20))
        if loss_kind == "bce":
            lo, gu, gi = O.plain_mf_bce_grads(nu, ni, u, i, j, l2u, l2i)
        else:
            lo, gu, gi, _ = O.plain_mf_bpr_grads(nu, ni, zb_u, zb_i, u, i, j)
        O.adam_step(nu, gu, nmu, nvu, s + 1, lr); O.adam_step(ni, gi, nmi, nvi, s + 1, lr)
        ops.plain_mf_step(pu, pi, mu, vu, mi, vi, hu, hi, T(u, dev), T(i, dev), T(j, dev), st, lr, loss,
                          loss=ops.LOSS_BCE if loss_kind == "bce" else ops.LOSS_BPR, l2_u=l2u, l2_i=l2i,
                          optimizer=ops.OPT_ADAM_DENSE_EXACT, stamp_user=su, stamp_item=si)
        assert abs(loss[0].item() - float(lo)) < 1e-4 * max(1.0, abs(float(lo))), (s, loss[0].item(), float(lo))
        assert int(hu.max()) == -1 and int(hi.max()) == -1 and int(hu.min()) == -1      # list heads re-armed
    assert int(st[0]) == steps
    ops.adam_flush(pu, mu, vu, su, st); ops.adam_flush(pi, mi, vi, si, st)
    # (fp32 vs fp64 oracle runs of this scenario differ by < 1e-6: a 2e-5 bound catches a single skipped or doubled
    # zero-gradient replay, which moves a row by ~0.6 lr)
    assert np.abs(pu.cpu().numpy() - nu).max() < 2e-5 and np.abs(pi.cpu().numpy() - ni).max() < 2e-5
    assert rel_err(mu.cpu().numpy(), nmu) < 1e-4 and rel_err(vi.cpu().numpy(), nvi) < 1e-4
    untouched = np.setdiff1d(np.arange(U), np.arange(250))
    assert np.array_equal(pu.cpu().numpy()[untouched], ut[untouched])            # rows no gradient ever reached


def test_plain_mf_fused_step_workspace_outlives_optimizer_state(dev):
    """The scratch of sml_plain_mf_step is shared by every caller in the process.  A run that ends on an odd step with
    duplicate ids in its last batch must not leak its duplicate-row work list into the next run, which starts a fresh Adam state
    at step 1 with a smaller batch (the list counters are double-buffered by LAUNCH parity kept in the scratch itself)."""
    from sml_b200 import ops
    rng = np.random.default_rng(13)
    z = torch.zeros_like

    def fresh(U, I):
        ut = (0.3 * rng.standard_normal((U, 64))).astype(np.float32); it = (0.3 * rng.standard_normal((I, 64))).astype(np.float32)
        pu, pi = T(ut, dev), T(it, dev)
        st = ops.new_adam_state(dev, history=True)
        return dict(ut=ut, it=it, pu=pu, pi=pi, mu=z(pu), vu=z(pu), mi=z(pi), vi=z(pi), st=st, su=ops.new_row_stamps(U, st),
                    si=ops.new_row_stamps(I, st), hu=ops.new_list_heads(U, dev), hi=ops.new_list_heads(I, dev), loss=torch.zeros(2, device=dev))

    def step(s, u, i, j):
        ops.plain_mf_step(s["pu"], s["pi"], s["mu"], s["vu"], s["mi"], s["vi"], s["hu"], s["hi"], T(u, dev), T(i, dev), T(j, dev), s["st"],
                          0.01, s["loss"], loss=ops.LOSS_BCE, l2_u=1e-3, l2_i=1e-3, optimizer=ops.OPT_ADAM_DENSE_EXACT,
                          stamp_user=s["su"], stamp_item=s["si"])
    a = fresh(64, 64)
    for _ in range(3):                                        # odd number of steps, 4 096 triples on 64 rows: every row is a duplicate
        step(a, *(rng.integers(0, 64, 4096).astype(np.int64) for _ in range(3)))
    b = fresh(300, 400)
    nu, ni = b["ut"].copy(), b["it"].copy()
    nmu, nvu, nmi, nvi = (np.zeros_like(x) for x in (nu, nu, ni, ni))
    for s in range(4):
        u, i, j = (rng.integers(0, n, 97).astype(np.int64) for n in (40, 60, 60))
        lo, gu, gi = O.plain_mf_bce_grads(nu, ni, u, i, j, 1e-3, 1e-3)
        O.adam_step(nu, gu, nmu, nvu, s + 1, 0.01); O.adam_step(ni, gi, nmi, nvi, s + 1, 0.01)
        step(b, u, i, j)
        assert abs(b["loss"][0].item() - float(lo)) < 1e-4 * max(1.0, abs(float(lo)))
    ops.adam_flush(b["pu"], b["mu"], b["vu"], b["su"], b["st"]); ops.adam_flush(b["pi"], b["mi"], b["vi"], b["si"], b["st"])
    assert np.abs(b["pu"].cpu().numpy() - nu).max() < 2e-5 and np.abs(b["pi"].cpu().numpy() - ni).max() < 2e-5
    assert rel_err(b["mu"].cpu().numpy(), nmu) < 1e-4 and rel_err(b["vi"].cpu().numpy(), nvi) < 1e-4


def test_plain_mf_fused_step_modes(dev):
    """(a) the fused step against the two-kernel path (sml_plain_mf_grads + dense Adam sweep); (b) when the same rows are
    touched at every step the exact and the sparse mode are the same arithmetic: bit-identical; (c) SML_OPT_ADAM_SPARSE
    moves only the rows of the batch."""
    from sml_b200 import ops
    rng = np.random.default_rng(12)
    U, I, B = 500, 1200, 400
    ut = (0.3 * rng.standard_normal((U, 64))).astype(np.float32); it = (0.3 * rng.standard_normal((I, 64))).astype(np.float32)
    u = rng.permutation(U)[:B].astype(np.int64)
    ij = rng.permutation(I)[:2 * B].astype(np.int64)
    i, j = ij[:B].copy(), ij[B:].copy()
    z = torch.zeros_like
    # two-kernel path
    pu0, pi0 = T(ut, dev), T(it, dev)
    gu, gi, loss0 = z(pu0), z(pi0), torch.zeros(2, device=dev)
    mu0, vu0, mi0, vi0 = z(pu0), z(pu0), z(pi0), z(pi0)
    st0 = ops.new_adam_state(dev)
    for _ in range(2):
        ops.plain_mf_grads(pu0, pi0, T(u, dev), T(i, dev), T(j, dev), gu, gi, loss0, loss=ops.LOSS_BCE, l2_u=1e-3, l2_i=1e-3)
        ops.adam_tick(st0, 0.01)
        ops.adam_dense(pu0, mu0, vu0, gu, st0); ops.adam_dense(pi0, mi0, vi0, gi, st0)
    got = {}
    for mode in (ops.OPT_ADAM_DENSE_EXACT, ops.OPT_ADAM_SPARSE):
        pu, pi, loss = T(ut, dev), T(it, dev), torch.zeros(2, device=dev)
        mu, vu, mi, vi = z(pu), z(pu), z(pi), z(pi)
        st = ops.new_adam_state(dev, history=True)
        su, si = ops.new_row_stamps(U, st), ops.new_row_stamps(I, st)
        hu, hi = ops.new_list_heads(U, dev), ops.new_list_heads(I, dev)
        for _ in range(2):
            ops.plain_mf_step(pu, pi, mu, vu, mi, vi, hu, hi, T(u, dev), T(i, dev), T(j, dev), st, 0.01, loss, loss=ops.LOSS_BCE,
                              l2_u=1e-3, l2_i=1e-3, optimizer=mode, stamp_user=su, stamp_item=si)
        if mode == ops.OPT_ADAM_DENSE_EXACT:
            ops.adam_flush(pu, mu, vu, su, st); ops.adam_flush(pi, mi, vi, si, st)
        # (the dot products are summed in a different lane order than in the two-kernel path: last-bit differences)
        assert (pu - pu0).abs().max().item() < 2e-6 and (pi - pi0).abs().max().item() < 2e-6
        assert rel_err(mu.cpu().numpy(), mu0.cpu().numpy()) < 1e-5 and rel_err(vi.cpu().numpy(), vi0.cpu().numpy()) < 1e-5
        assert abs(loss[1].item() - loss0[1].item()) < 1e-5
        got[mode] = (pu, pi, mu, vi)
    for a, b in zip(got[ops.OPT_ADAM_DENSE_EXACT], got[ops.OPT_ADAM_SPARSE]):
        assert torch.equal(a, b)
    # sparse mode: a row touched at step 1 only does NOT move at step 2 (dense Adam would move it through its momentum)
    pu, pi, loss = T(ut, dev), T(it, dev), torch.zeros(2, device=dev)
    mu, vu, mi, vi = z(pu), z(pu), z(pi), z(pi)
    st = ops.new_adam_state(dev, history=True)
    hu, hi = ops.new_list_heads(U, dev), ops.new_list_heads(I, dev)
    args = dict(loss=ops.LOSS_BCE, l2_u=1e-3, l2_i=1e-3, optimizer=ops.OPT_ADAM_SPARSE)
    ops.plain_mf_step(pu, pi, mu, vu, mi, vi, hu, hi, T(u[:10], dev), T(i[:10], dev), T(j[:10], dev), st, 0.01, loss, **args)
    snap = pu[u[:10]].clone()
    ops.plain_mf_step(pu, pi, mu, vu, mi, vi, hu, hi, T(u[10:20], dev), T(i[10:20], dev), T(j[10:20], dev), st, 0.01, loss, **args)
    assert torch.equal(pu[u[:10]], snap) and not torch.equal(pu[u[10:20]], T(ut, dev)[u[10:20]])


def test_eval_sees_kernel_updates(dev):
    """The step kernels write the tables through raw pointers (no tensor version bump): a test set
    evaluated before and after a training step must not return a stale scoring pass."""
    from sml_b200 import ops
    from sml_b200.model.MF import MFbasemode
    from sml_b200.evalution.evaluation2 import DeviceTestSet, test_model
    rng = np.random.default_rng(2)
    with torch.random.fork_rng(devices=[]):
        mf = MFbasemode(64, 128, 64).to(dev)
    rows = T(np.concatenate([rng.integers(0, 64, (200, 1)), rng.integers(0, 128, (200, 50))], 1).astype(np.int64), dev)
    ts = DeviceTestSet(rows)
    r0, n0 = test_model(mf, ts, topK=5)
    uw, iw = mf.user_laten.weight.data, mf.item_laten.weight.data
    g_u, g_i = torch.zeros_like(uw), torch.zeros_like(iw)
    loss = torch.zeros(2, device=dev)
    u = rows[:, 0].contiguous(); i = rows[:, 1].contiguous(); j = rows[:, 2].contiguous()
    st = ops.new_adam_state(dev)
    for _ in range(30):
        ops.plain_mf_grads(uw, iw, u, i, j, g_u, g_i, loss)
        ops.adam_tick(st, 0.05)
        ops.adam_dense(uw, torch.zeros_like(uw), torch.zeros_like(uw), g_u, st)
        ops.adam_dense(iw, torch.zeros_like(iw), torch.zeros_like(iw), g_i, st)
    r1, n1 = test_model(mf, ts, topK=5)
    r2, n2 = test_model(mf, DeviceTestSet(rows), topK=5)
    assert r1 == r2 and float(n1) == float(n2)
    assert r1 > r0                                   # training on the positives must move them up
