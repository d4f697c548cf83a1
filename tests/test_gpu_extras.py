"""GPU: device-side sampler, checkpoint / resume of the whole SML state, CLI-level device_sampler run."""
import numpy as np
import pytest
import torch

from tests.test_host_logic import make_args, write_fixture_stream

pytestmark = pytest.mark.gpu


def test_philox_negatives_semantics():
    """Same exclusion semantics as offlineDataset_withsample (data/dataset.py:62-71): the negative is an item of
    this period that the user did not interact with in this period; deterministic per (seed, offset)."""
    from sml_b200 import ops
    from sml_b200.data.dataset import offlineDataset_withsample
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    data = np.stack([rng.integers(0, 50, 4000), rng.integers(0, 80, 4000)], 1)
    ds = offlineDataset_withsample(data)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
    u = T(ds.user)
    neg = ops.philox_negatives(u, T(ds.item_all), T(ds._keys), ds._span, seed=7, offset=1)
    n = neg.cpu().numpy()
    assert np.isin(n, ds.item_all).all()
    assert not ds.interacted(ds.user, n).any()
    assert torch.equal(neg, ops.philox_negatives(u, T(ds.item_all), T(ds._keys), ds._span, seed=7, offset=1))
    assert not torch.equal(neg, ops.philox_negatives(u, T(ds.item_all), T(ds._keys), ds._span, seed=7, offset=2))
    # every item of the period is reachable (the exclusion is per user, so the histogram is not flat)
    counts = np.bincount(n, minlength=80)[ds.item_all]
    assert counts.min() > 0


def _make_meta(g, tmp, **kw):
    from sml_b200.data.dataset2 import transfer_data
    from sml_b200.model.transfer import meta_train
    NP, U, I = write_fixture_stream(g, tmp)
    args = make_args(g, tmp)
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
    ds = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)],
                       test_list=[str(j) for j in range(5, NP)], validation_list=None, online_train_time=2, online_test_time=5)
    return args, meta_train(args, ds, U, I, 64, **kw)


def test_checkpoint_resume_is_exact(golden, tmp_path):
    """Run 2 stages, checkpoint, run 2 more; a fresh object restored from the checkpoint must reproduce the last 2
    stages bit for bit (tables, theta, Adam state, step counters, RNG streams)."""
    g = golden("period_run")
    args, a = _make_meta(g, str(tmp_path / "a"))
    a.train_one_stage3(args, 0); a.train_one_stage3(args, 1)
    sd = a.state_dict()
    a.train_one_stage3(args, 2); a.train_one_stage3(args, 3)
    args_b, b = _make_meta(g, str(tmp_path / "b"))
    b.load_state_dict(sd)
    b.train_one_stage3(args_b, 2); b.train_one_stage3(args_b, 3)
    # row-gradient scatter and theta-gradient reductions use fp32 atomics: equal to rounding, not bitwise
    assert (a.MFbase.user_laten.weight.data - b.MFbase.user_laten.weight.data).abs().max().item() < 1e-4
    assert (a.MFbase.item_laten.weight.data - b.MFbase.item_laten.weight.data).abs().max().item() < 1e-4
    assert (a.transfer.theta - b.transfer.theta).abs().max().item() < 1e-4
    assert a.MF_optimizer.step_count == b.MF_optimizer.step_count and a.recall == b.recall


def test_device_sampler_run(golden, tmp_path):
    """Throughput mode: batches sampled on the GPU.  Not bit-comparable with the reference (different RNG), but the
    loop must train: finite tables and a validation recall in the range of the reference's run."""
    g = golden("period_run")
    args, m = _make_meta(g, str(tmp_path), device_sampler=True, emulate_reference_rng=False)
    m.run(args)
    assert torch.isfinite(m.MFbase.user_laten.weight.data).all() and torch.isfinite(m.transfer.theta).all()
    assert len(m.recall) == 3 and abs(np.mean(m.recall) - np.mean(g["recall"])) < 0.15


def test_row_lazy_adam_is_bit_identical_to_the_dense_sweep():
    """ops.adam_rows / adam_flush replay the zero-gradient steps a row missed: after a flush the table, exp_avg and
    exp_avg_sq must equal the dense sweep (model/transfer.py:392, dense torch.optim.Adam over nn.Embedding) bit for bit,
    and a row read right after its catch-up must equal the dense row before that step's update."""
    from sml_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(5)
    N, T, K = 3000, 45, 257
    p0 = torch.randn(N, 64, generator=gen).to(dev)
    pd, md, vd, gd = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0), torch.zeros_like(p0)
    pl, ml, vl, gl = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0), torch.zeros_like(p0)
    sd, sl = ops.new_adam_state(dev), ops.new_adam_state(dev, history=True)
    stamp = ops.new_row_stamps(N, sl)
    for t in range(T):
        # Zipf-ish ids with duplicates; a third of the table never appears
        ids = (torch.rand(K, generator=gen) ** 3 * (2 * N // 3)).long().to(dev)
        grads = (torch.randn(K, 64, generator=gen) * 0.1).to(dev)
        ops.adam_tick(sd, 0.01); ops.adam_tick(sl, 0.01)
        ops.adam_rows(pl, ml, vl, None, stamp, ids, sl, apply=False)
        assert torch.equal(pl[ids], pd[ids]), "step %d: caught-up rows differ from the dense table" % t
        u, inv = torch.unique(ids, return_inverse=True)          # deterministic duplicate sum, same for both arms
        acc = torch.zeros(u.numel(), 64, device=dev).index_add_(0, inv, grads)
        gd[u] = acc; gl[u] = acc
        ops.adam_dense(pd, md, vd, gd, sd)
        ops.adam_rows(pl, ml, vl, gl, stamp, ids, sl, apply=True)
        assert float(gl.abs().max()) == 0.0 and float(gd.abs().max()) == 0.0
        if t == 20:
            ops.adam_flush(pl, ml, vl, stamp, sl)
            assert torch.equal(pl, pd)
    ops.adam_flush(pl, ml, vl, stamp, sl)
    assert torch.equal(pl, pd) and torch.equal(ml, md) and torch.equal(vl, vd)
    # every row is at step T, except the rows no gradient ever reached: they keep SML_STAMP_IDLE (exp_avg = exp_avg_sq = 0, skipped
    # by flushes on the stamp alone), and those are exactly the rows whose moments are all zero
    idle = stamp == ops.STAMP_IDLE
    assert bool(((stamp == T) | idle).all()) and 0 < int(idle.sum()) < N
    assert torch.equal(idle, (ml == 0).all(1) & (vl == 0).all(1))
    # stamps that start at the current step (moments not known to be zero) end the same way: the flush marks the idle rows itself
    st2 = ops.new_row_stamps(N, sl, idle=False)
    ops.adam_tick(sl, 0.01)
    ops.adam_flush(pl, ml, vl, st2, sl)
    assert torch.equal(st2 == ops.STAMP_IDLE, idle) and bool(((st2 == T + 1) | (st2 == ops.STAMP_IDLE)).all())


def test_mf_epoch_row_lazy_adam_matches_dense():
    """sml_mf_epoch with row stamps (row-lazy exact Adam + flush) against the same epoch with the dense sweeps."""
    from sml_b200 import ops
    from sml_b200.model.conv_transfer import ConvTransfer_com
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    U, I, B, n = 5000, 9000, 256, 256 * 9 + 100
    tr = ConvTransfer_com(64, 64).to(dev)
    lu, li = torch.randn(U, 64, device=dev) * 0.3, torch.randn(I, 64, device=dev) * 0.3
    hu0, hi0 = lu + 0.05 * torch.randn_like(lu), li + 0.05 * torch.randn_like(li)
    user, item, neg = torch.randint(0, U, (n,), device=dev), torch.randint(0, I, (n,), device=dev), torch.randint(0, I, (n,), device=dev)
    out = []
    for lazy in (False, True):
        hu, hi = hu0.clone(), hi0.clone()
        z = {k: torch.zeros_like(hu if "user" in k else hi) for k in ("g_user", "m_user", "v_user", "g_item", "m_item", "v_item")}
        st = ops.new_adam_state(dev, history=lazy)
        if lazy:
            z.update(stamp_user=ops.new_row_stamps(U, st), stamp_item=ops.new_row_stamps(I, st))
        loss = torch.zeros(2, device=dev)
        a = ops.make_step_args(user=user, item=item, neg=neg, batch=B, last_user=lu, last_item=li, hat_user=hu, hat_item=hi,
                               theta=tr.theta, adam_state=st, lr=0.01, l2=1e-6, loss_out=loss, **z)
        for _ in range(2):
            ops.mf_epoch(a, n)
        out.append((hu, hi, z["m_user"], z["v_item"], loss.clone(), int(st[0])))
    d, l = out
    assert d[5] == l[5] == 2 * 10
    # the gradient scatter and the split-K GEMM slices use fp32 atomics, so the two runs agree to rounding, not bitwise; Adam
    # turns an absolute gradient noise d (~1e-12 here) on elements with |g| <~ eps into up to lr * d / eps ~ 1e-6 per step
    for x, y in zip(d[:4], l[:4]):
        assert (x - y).abs().max().item() < 2e-5
    assert abs(float(d[4][1] - l[4][1])) < 1e-4 * abs(float(d[4][1]))


def test_eval_prefilter_counts_equal_the_exact_kernel():
    """sml_eval_candidates_prefilter (bf16 estimate + error bound + exact fp32 fallback) must return the same gt / eq as
    sml_eval_candidates for every input: random scores, exact ties, near-ties far inside the bf16 error, NaN / Inf / zero rows,
    huge and tiny magnitudes, ragged candidate counts."""
    from sml_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(9)
    U, I = 700, 5000
    ut = torch.randn(U, 64, generator=gen)
    it = torch.randn(I, 64, generator=gen)
    # adversarial item rows
    it[10] = it[11]                                   # exact duplicate -> exact ties when both are candidates
    it[12] = it[11] * (1 + 1e-7)                      # differs in the last fp32 bits only
    it[13] = it[11] + 1e-6 * torch.randn(64, generator=gen)
    it[14] = float("nan"); it[15, 3] = float("inf"); it[16, 5] = -float("inf"); it[17] = 0.0; it[18] = -0.0
    it[19] = 3e37; it[20] = 1e-38; it[21] = 1e-42     # overflow to inf in bf16 / fp32 denormals
    it[22:40] = it[11] + 3e-4 * torch.randn(18, 64, generator=gen)   # inside the bf16 error band of row 11
    ut[5] = 0.0; ut[6] = float("nan"); ut[7, 2] = float("inf"); ut[8] = 1e-30; ut[9] = 1e18
    ut, it = ut.to(dev), it.to(dev)
    for C in (1000, 37, 1):
        n = 900
        rows = torch.empty(n, 1 + C, dtype=torch.int64)
        rows[:, 0] = torch.randint(0, U, (n,), generator=gen)
        rows[:, 1:] = torch.randint(0, I, (n, C), generator=gen)
        rows[:200, 0] = torch.randint(0, 12, (200,), generator=gen)            # the special user rows
        rows[:, 1][::3] = 11                                                   # positive = the row with look-alikes
        if C > 30:
            rows[:, 2:32] = torch.randint(10, 40, (n, 30), generator=gen)     # candidates from the adversarial block
        rows = rows.to(dev)
        g0, e0 = ops.eval_candidates(ut, it, rows, prefilter=False)
        g1, e1 = ops.eval_candidates(ut, it, rows, prefilter=True)
        assert torch.equal(g0, g1) and torch.equal(e0, e1), C
        assert int(e0.sum()) > 0 or C == 1
    # an empty file
    z = ops.eval_candidates(ut, it, torch.empty(0, 1001, dtype=torch.int64, device=dev), prefilter=True)
    assert z[0].numel() == 0
