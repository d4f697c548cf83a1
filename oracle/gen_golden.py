"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, CPU torch) -- build-container only, TEST INFRASTRUCTURE.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

The reference has no tests or known-answer vectors, so these
outputs of the reference itself are what the oracle (oracle/sml_oracle.py) and the
CUDA path are pinned to.  Fixtures are small (a few hundred KB) and committed.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from sml_b200.data import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
D = 64


def npz(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, len(arrs), "arrays", os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB")


def sample(a):
    """Large (fc-weight sized) arrays are stored as a strided sample to keep fixtures small."""
    a = np.asarray(a)
    return a[::5, ::7].copy() if a.ndim == 2 and a.size > 20000 else a.copy()


def flat_theta(prefix, net):
    return {prefix + k: sample(v.detach().numpy()) for k, v in net.state_dict().items()}


def flat_grads(prefix, net):
    return {prefix + k: sample(p.grad.detach().numpy()) for k, p in net.named_parameters()}


def seed_theta(module, seed, rows):
    """Overwrite a ConvTransfer(_com) module's parameters with oracle.init_theta(default_rng(seed))
    (user net) / default_rng(seed+1) (item net): the tests regenerate theta from the same seeds
    instead of storing 4 x 790 KB per fixture.  Returns a checksum dict."""
    from oracle import sml_oracle as O
    chk = {}
    for off, name in ((0, "user_transfer"), (1, "item_transfer")):
        th = O.init_theta(np.random.default_rng(seed + off), d=D, rows=rows)
        getattr(module, name).load_state_dict({k: torch.from_numpy(v) for k, v in th.items()})
        chk[name] = float(sum(np.abs(v.astype(np.float64)).sum() for v in th.values()))
    return np.array([seed, rows, chk["user_transfer"], chk["item_transfer"]], dtype=np.float64)


def gen_transfer_fwd(ns):
    torch.manual_seed(11)
    com = ns.conv_transfer.ConvTransfer_com(D, D)
    conv = ns.conv_transfer.ConvTransfer(D, D)
    chk_com, chk_conv = seed_theta(com, 100, 3), seed_theta(conv, 200, 2)
    g = torch.Generator().manual_seed(5)
    x_t = torch.randn(37, D, generator=g)
    x_hat = torch.randn(37, D, generator=g) * 0.5
    with torch.no_grad():
        out = dict(x_t=x_t.numpy(), x_hat=x_hat.numpy(),
                   com_user=com(x_t, x_hat, "user").numpy(), com_item=com(x_t, x_hat, "item").numpy(),
                   conv_user=conv(x_t, x_hat, "user").numpy(), conv_item=conv(x_t, x_hat, "item").numpy())
    out["theta_com"] = chk_com; out["theta_conv"] = chk_conv
    npz("transfer_fwd.npz", **out)


def gen_run_mf(ns):
    torch.manual_seed(12)
    com = ns.conv_transfer.ConvTransfer_com(D, D)
    conv = ns.conv_transfer.ConvTransfer(D, D)
    chk_com, chk_conv = seed_theta(com, 110, 3), seed_theta(conv, 210, 2)
    g = torch.Generator().manual_seed(6)
    B = 24
    rows = {k: torch.randn(B, D, generator=g) * (0.3 if "hat" in k else 1.0)
            for k in ("u_last", "u_hat", "i_last", "i_hat", "j_last", "j_hat")}
    out = {k: v.numpy().copy() for k, v in rows.items()}
    for tag, net, kw in (("com_bce", com, dict(BCE=True)), ("com_bpr", com, dict(BCE=False)), ("conv_bpr", conv, {})):
        leaf = {k: v.clone().requires_grad_("hat" in k) for k, v in rows.items()}
        net.zero_grad()
        loss = net.run_MF(leaf["u_last"], leaf["u_hat"], leaf["i_last"], leaf["i_hat"], leaf["j_last"], leaf["j_hat"], **kw)
        loss.backward()
        out[tag + ".loss"] = np.float32(loss.item())
        for k in ("u_hat", "i_hat", "j_hat"):
            out[tag + ".d_" + k] = leaf[k].grad.numpy().copy()
        out.update(flat_grads(tag + ".g_user.", net.user_transfer))
        out.update(flat_grads(tag + ".g_item.", net.item_transfer))
    out["theta_com"] = chk_com; out["theta_conv"] = chk_conv
    npz("run_mf.npz", **out)


def gen_mf_steps(ns):
    """HOT LOOP A body (model/transfer.py:463-511) with supplied batches, through the
    reference modules and torch.optim.Adam; 3 steps; duplicate ids inside a batch."""
    torch.manual_seed(13)
    U, I, B, steps = 50, 70, 16, 3
    mf = ns.MF.MFbasemode(U, I, D)
    com = ns.conv_transfer.ConvTransfer_com(D, D)
    chk_com = seed_theta(com, 120, 3)
    last_u = torch.randn(U, D); last_i = torch.randn(I, D)
    opt = torch.optim.Adam(mf.parameters(), lr=0.01, weight_decay=0)
    rng = np.random.default_rng(3)
    ids = np.stack([np.stack([rng.integers(0, 12, B), rng.integers(0, 20, B), rng.integers(0, 20, B)]) for _ in range(steps)])
    out = dict(ids=ids.astype(np.int64), last_user=last_u.numpy().copy(), last_item=last_i.numpy().copy(),
               user0=mf.user_laten.weight.detach().numpy().copy(), item0=mf.item_laten.weight.detach().numpy().copy(),
               lr=np.float64(0.01), l2=np.float64(1e-6), theta_com=chk_com)
    l2 = 1e-6
    for s in range(steps):
        user, item, neg = (torch.from_numpy(ids[s, k]).long() for k in range(3))
        mf.zero_grad(); com.zero_grad()
        wu, wi, wj = mf.user_laten(user), mf.item_laten(item), mf.item_laten(neg)
        loss = com.run_MF(last_u[user], wu, last_i[item], wi, last_i[neg], wj, norm=False)
        l2loss = 0.5 * torch.sum(wu ** 2 + wi ** 2 + wj ** 2)
        loss = loss + l2 * l2loss
        loss.backward()
        out["loss%d" % s] = np.float32(loss.item())
        out["gu%d" % s] = mf.user_laten.weight.grad.numpy().copy()
        out["gi%d" % s] = mf.item_laten.weight.grad.numpy().copy()
        opt.step()
        out["user%d" % (s + 1)] = mf.user_laten.weight.detach().numpy().copy()
        out["item%d" % (s + 1)] = mf.item_laten.weight.detach().numpy().copy()
    st = opt.state[mf.user_laten.weight]
    out["m_user"] = st["exp_avg"].numpy().copy(); out["v_user"] = st["exp_avg_sq"].numpy().copy()
    st = opt.state[mf.item_laten.weight]
    out["m_item"] = st["exp_avg"].numpy().copy(); out["v_item"] = st["exp_avg_sq"].numpy().copy()
    assert mf.user_bais.weight.grad is None      # bias tables never get a gradient
    npz("mf_steps.npz", **out)


def gen_tr_steps(ns):
    """HOT LOOP B body (model/transfer.py:701-728): theta-only grads, Adam with coupled L2."""
    torch.manual_seed(14)
    U, I, B, steps = 40, 60, 8, 3
    com = ns.conv_transfer.ConvTransfer_com(D, D)
    chk_com = seed_theta(com, 130, 3)
    tabs = dict(last_user=torch.randn(U, D), user_hat=torch.randn(U, D) * 0.7,
                last_item=torch.randn(I, D), item_hat=torch.randn(I, D) * 0.7)
    opt = torch.optim.Adam(com.parameters(), lr=0.001, weight_decay=1e-4)
    rng = np.random.default_rng(4)
    ids = np.stack([np.stack([rng.integers(0, U, B), rng.integers(0, I, B), rng.integers(0, I, B)]) for _ in range(steps)])
    out = dict(ids=ids.astype(np.int64), lr=np.float64(0.001), wd=np.float64(1e-4), theta_com=chk_com)
    out.update({k: v.numpy().copy() for k, v in tabs.items()})
    for s in range(steps):
        user, item, neg = (torch.from_numpy(ids[s, k]).long() for k in range(3))
        com.zero_grad()
        loss = com.run_MF(tabs["last_user"][user], tabs["user_hat"][user], tabs["last_item"][item],
                          tabs["item_hat"][item], tabs["last_item"][neg], tabs["item_hat"][neg], norm=False)
        loss.backward()
        out["loss%d" % s] = np.float32(loss.item())
        if s == 0:
            out.update(flat_grads("g0.user.", com.user_transfer)); out.update(flat_grads("g0.item.", com.item_transfer))
        opt.step()
    out.update(flat_theta("t3.user.", com.user_transfer)); out.update(flat_theta("t3.item.", com.item_transfer))
    npz("tr_steps.npz", **out)


def gen_eval(ns):
    torch.manual_seed(15)
    U, I, N, C = 30, 200, 45, 60
    mf = ns.MF.MFbasemode(U, I, D)
    rng = np.random.default_rng(7)
    rows = np.concatenate([rng.integers(0, U, (N, 1)), np.stack([rng.permutation(I)[:C] for _ in range(N)])], axis=1)
    # make some positives strong so that hits exist at small K
    with torch.no_grad():
        for r in range(0, N, 3):
            mf.item_laten.weight[rows[r, 1]] = mf.user_laten.weight[rows[r, 0]] * (0.2 + 0.1 * (r % 5))
    out = dict(rows=rows.astype(np.int64), user=mf.user_laten.weight.detach().numpy().copy(),
               item=mf.item_laten.weight.detach().numpy().copy())
    data = torch.from_numpy(rows).long()
    with torch.no_grad():
        ue = mf.user_laten(data[:, 0]).unsqueeze(1)
        sc = torch.mul(ue, mf.item_laten(data[:, 1:])).sum(-1)
        out["scores"] = sc.numpy().copy()
        for K in (20, 10, 5):
            h, nd, idx = mf.test(data, topK=K)
            out["hits@%d" % K] = np.float64(h)
            out["ndcg@%d" % K] = np.float32(float(nd))
            out["idx@%d" % K] = idx.numpy().copy()
            loader = torch.utils.data.DataLoader(ns.dataset2.testDataset(rows), batch_size=16)
            r, n = ns.raw_test_model(mf, loader, topK=K)
            out["recall@%d" % K] = np.float64(r); out["tm_ndcg@%d" % K] = np.float64(n)
        # tie behaviour of the reference's topk on this torch build (CPU)
        tie = torch.tensor([[1.0, 1.0, 1.0, 0.5], [2.0, 3.0, 2.0, 2.0], [0.0, 0.0, 0.0, 0.0]])
        out["tie_scores"] = tie.numpy()
        out["tie_top2"] = torch.topk(tie, 2)[1].numpy()
    # legacy per-user evaluation (evalution/evaluation.py)
    users = [1, 2]; pos = [[3, 4], [5]]; negs = [list(range(10, 40)), list(range(50, 80))]
    with torch.no_grad():       # guarantee at least one hit per user (an all-miss list makes the reference's
        mf.item_laten.weight[3] = mf.user_laten.weight[1] * 0.5      # torch.tensor(NDCGs).mean() fail on ints)
        mf.item_laten.weight[5] = mf.user_laten.weight[2] * 0.5
    out["legacy_item"] = mf.item_laten.weight.detach().numpy().copy()
    res = ns.evaluation.test_model(mf, (users, pos, negs), topK=5)
    out["legacy"] = np.array([float(x) for x in res], dtype=np.float64)
    npz("eval.npz", **out)


def gen_period_run(ns, stop=False, news=False, opts=False):
    """End-to-end meta_train.run on a tiny stream, recording every batch the reference's
    DataLoaders produced so that the CUDA path can replay the same supplied triples."""
    U, I, NP, N, NNEG = 120, 150, 8, 96, 40
    periods = synth.make_stream(U, I, N, NP, n_neg=NNEG, seed=21 + news, churn=0.5 if news else 0.0)
    tmp = tempfile.mkdtemp(prefix="sml_golden_")
    synth.write_stream(tmp + "/", "mini", periods, U, I)
    parser = ns.main_yelp.get_parse()
    args = parser.parse_args([])
    args.data_name = "yelp"; args.data_path = tmp + "/"; args.numworkers = 0
    args.MF_batch_size = 32; args.TR_batch_size = 16; args.multi_num = 2
    args.MF_epochs = 1; args.TR_epochs = 1; args.pre_model = os.path.join(tmp, "pre.pkl")
    args.TR_stop_ = bool(stop)
    if opts:        # the options that are off by default: --need_adaptive, --clip_grad (tight max norm so it bites), --norm
        args.need_adaptive = True; args.clip_grad = True; args.maxnorm_grad = 0.05; args.norm = True
    if news:        # configs[2]: main_news.py settings (MF_epochs=2, TR_epochs=2) on a high-churn stream; data_name != 'yelp'
        args.data_name = "news"; args.MF_epochs = 2; args.TR_epochs = 2        # takes the other constructor branch (:314-325)
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
    pre = ns.MF.MFbasemode(U, I, D)
    torch.save(pre, args.pre_model)
    out = dict(pre_user=pre.user_laten.weight.detach().numpy().copy(), pre_item=pre.item_laten.weight.detach().numpy().copy(),
               n_periods=np.int64(NP), U=np.int64(U), I=np.int64(I))
    for p, (tr, te) in enumerate(periods):
        out["train%d" % p] = tr.astype(np.int32); out["test%d" % p] = te.astype(np.int32)

    log = []            # (kind, [triples...]) per dataset instance, in creation order

    def recording(cls, kind):
        class Rec(cls):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                self._rec = []
                log.append((kind, self._rec))

            def __getitem__(self, idx):
                t = super().__getitem__(idx)
                self._rec.append((int(idx), int(t[0]), int(t[1]), int(t[2])))
                return t
        return Rec
    ns.transfer.PreSampleDatast = recording(ns.dataset2.trainDataset_withPreSample, "MF")
    ns.transfer.SampleDaset = recording(ns.dataset.offlineDataset_withsample, "TR")

    # main_yelp.py:137-140 seeds, then dataset + model construction as :159-168
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
    file_list = [str(i) for i in range(NP)]
    test_list = [str(j) for j in range(5, NP)]
    ds = ns.dataset2.transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=file_list,
                                   test_list=test_list, validation_list=None, online_train_time=2, online_test_time=5)
    meta = ns.transfer.meta_train(args, ds, int(ds.user_number), int(ds.item_number), args.laten)
    out["theta_com"] = seed_theta(meta.transfer, 140, 3)
    stage_sums = []
    orig = meta.train_one_stage3

    def wrapped(a, stage_id):
        flag = orig(a, stage_id)
        if flag:
            stage_sums.append([float(meta.MFbase.user_laten.weight.double().sum()), float(meta.MFbase.user_laten.weight.double().abs().sum()),
                               float(meta.MFbase.item_laten.weight.double().sum()), float(meta.MFbase.item_laten.weight.double().abs().sum()),
                               float(sum(p.double().abs().sum() for p in meta.transfer.parameters()))])
        return flag
    meta.train_one_stage3 = wrapped
    # every test_model call of the run, in order: (topK, rows, recall, ndcg) -- model/transfer.py:444-446,517-519,
    # 684-686,738-741 (validation passes) and :810-823,855-868 (the real test at K = 20, 10, 5)
    eval_log = []
    tm = ns.transfer.test_model

    def logged_test_model(model, test_set, *a, **k):
        r, n = tm(model, test_set, *a, **k)
        eval_log.append([float(k.get("topK", 10)), float(len(test_set.dataset)), float(r), float(n)])
        return r, n
    ns.transfer.test_model = logged_test_model
    meta.run(args)
    ns.transfer.test_model = tm
    out["eval_log"] = np.array(eval_log, dtype=np.float64).reshape(-1, 4)
    out["stage_sums"] = np.array(stage_sums)
    out["final_user"] = meta.MFbase.user_laten.weight.detach().numpy().copy()
    out["final_item"] = meta.MFbase.item_laten.weight.detach().numpy().copy()
    out["final_user_hat"] = meta.user_weight_hat.numpy().copy()
    out["final_last_user"] = meta.last_user_weight.numpy().copy()
    out.update(flat_theta("tF.user.", meta.transfer.user_transfer)); out.update(flat_theta("tF.item.", meta.transfer.item_transfer))
    for k in ("recall", "ndcg", "recall_10", "ndcg_10", "recall_5", "ndcg_5", "test_num"):
        out[k] = np.array([float(x) for x in getattr(meta, k)], dtype=np.float64)
    out["log_kinds"] = np.array([k for k, _ in log])
    for n, (_, rec) in enumerate(log):
        out["log%d" % n] = np.array(rec, dtype=np.int32).reshape(-1, 4)
    out["args"] = np.array([args.MF_batch_size, args.TR_batch_size, args.multi_num, (2 if news else 1), args.TR_epochs,
                            args.seed], dtype=np.int64)
    out["hyper"] = np.array([args.MF_lr, args.l2, args.TR_lr, args.TR_l2], dtype=np.float64)
    npz("period_run_opts.npz" if opts else ("period_run_news.npz" if news else ("period_run_stop.npz" if stop else "period_run.npz")), **out)
    # restore
    ns.transfer.PreSampleDatast = ns.dataset2.trainDataset_withPreSample
    ns.transfer.SampleDaset = ns.dataset.offlineDataset_withsample


def gen_select_neg(ns):
    """Reference test-file builder (data/dataset2.py:356-414) on a small stream, numpy seed 77, neg_num 25."""
    rng = np.random.default_rng(3)
    files = [np.stack([rng.integers(0, 40, 90), rng.integers(0, 1500, 90)], 1).astype(np.int64) for _ in range(5)]
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "mini", "test"))
        np.save(os.path.join(tmp, "mini", "information.npy"), np.array([450, 40, 1500]))
        for k, f in enumerate(files):
            np.save(os.path.join(tmp, "mini", "%d.npy" % k), f)
        np.random.seed(77)
        ns.dataset2.select_neg_forinteraction(path=tmp + "/", datasetname="mini", file_path_list=[str(k) for k in range(5)],
                                              leave_for_init_train=0.6, neg_num=25)
        tests = {"test%d" % k: np.load(os.path.join(tmp, "mini", "test", "%d.npy" % k)) for k in (3, 4)}
    npz("select_neg.npz", seed=np.array(77), neg_num=np.array(25), leave=np.array(0.6),
        **{"file%d" % k: f for k, f in enumerate(files)}, **tests)


def gen_baseline(ns):
    """The MF baselines of model/baseline.py on a tiny stream: Reservious, fine-tune and full-retrain runs
    (SPMF.run -> run_one_stage2), base_train, compute_R_W_P and sample_batch -- with every batch the reference's
    DataLoaders produced.  Extra shims for this module only: ``np.long`` (removed from numpy 1.24; :73,117,566),
    ``DataLoader(num_workers=4)`` -> 0 (deterministic global-RNG consumption, like --numworkers 0 on the SML path) and a
    no-op ``torch.save`` for base_train's hard-coded checkpoint path (:213,219).  The SPMF method itself cannot run in the
    reference (run_one_stage unpacks four values into two, :250)."""
    import contextlib
    import importlib
    import io
    import re
    import types
    if not hasattr(np, "long"):
        np.long = np.int64
    bl = importlib.import_module("model.baseline")
    U, I, NP, N, NNEG = 120, 150, 6, 96, 40
    periods = synth.make_stream(U, I, N, NP, n_neg=NNEG, seed=31)
    tmp = tempfile.mkdtemp(prefix="sml_golden_bl_")
    root = synth.write_stream(tmp + "/", "mini", periods, U, I) + "/"
    rng = np.random.default_rng(5)
    new_user = np.sort(rng.choice(U, 25, replace=False)).astype(np.int64); new_item = np.sort(rng.choice(I, 30, replace=False)).astype(np.int64)
    np.save(root + "test_new_user.npy", new_user); np.save(root + "test_new_item.npy", new_item)
    out = dict(U=np.int64(U), I=np.int64(I), n_periods=np.int64(NP), new_user=new_user, new_item=new_item)
    for p, (tr, te) in enumerate(periods):
        out["train%d" % p] = tr.astype(np.int32); out["test%d" % p] = te.astype(np.int32)

    # ---- Reservious ----
    np.random.seed(11)
    r = bl.Reservious(50)
    feed = [np.stack([rng.integers(0, U, n), rng.integers(0, I, n)], 1).astype(np.int64) for n in (20, 45, 70, 30)]
    for k, f in enumerate(feed):
        r.updata(f)
        out["res_feed%d" % k] = f; out["res_pool%d" % k] = r.pool.copy(); out["res_state%d" % k] = np.array([r.t, r.pool_have])
    r2 = bl.Reservious(30)
    r2.init_pool(feed[2])
    out["res_init_pool"] = r2.pool.copy(); out["res_init_state"] = np.array([r2.t, r2.pool_have])
    out["res_after_draw"] = np.array(np.random.rand())          # the generator position after all of the above

    DL = torch.utils.data.DataLoader

    def dl0(*a, **k):
        k["num_workers"] = 0
        return DL(*a, **k)
    log = []

    # (the class calls super(offlineDataset_withsample, self) by its module-global name: patch its methods, do not rebind it)
    cls = bl.offlineDataset_withsample
    cls_init, cls_get = cls.__init__, cls.__getitem__

    def rec_init(self, *a, **k):
        with contextlib.redirect_stdout(io.StringIO()):
            cls_init(self, *a, **k)
        self._rec = []
        log.append(self._rec)

    def rec_get(self, idx):
        t = cls_get(self, idx)
        self._rec.append((int(idx), int(t[0]), int(t[1]), int(t[2])))
        return t

    def run(method, tag, base=False):
        del log[:]
        args = bl.get_parse().parse_args([])
        args.lr = 0.01; args.l2_u = args.l2_i = 1e-3; args.epochs = 2; args.batch_size = 32; args.pool_size = 0; args.pool_init_type = 0
        torch.manual_seed(2000); np.random.seed(2002)
        ds = bl.StreamingData(root)
        model = bl.SPMF(args, ds, int(ds.user_num), int(ds.item_num), args.laten_dim)
        out[tag + "_init_user"] = model.MFbase.user_laten.weight.detach().numpy().copy()
        out[tag + "_init_item"] = model.MFbase.item_laten.weight.detach().numpy().copy()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            if base:
                model.base_train(3, 3, 1e-3, 2e-3)
                model.recall = [model.test(ds.get_next(3)[1])[0]]; model.ndcg = [model.test(ds.get_next(3)[1])[1]]
            else:
                model.run(2, method=method)
        losses = [float(x) for x in re.findall(r"loss:(-?[0-9.]+)", buf.getvalue())]
        out[tag + "_losses4"] = np.array(losses)                 # printed with 4 decimals (:203,363)
        out[tag + "_final_user"] = model.MFbase.user_laten.weight.detach().numpy().copy()
        out[tag + "_final_item"] = model.MFbase.item_laten.weight.detach().numpy().copy()
        out[tag + "_recall"] = np.array(model.recall, dtype=np.float64); out[tag + "_ndcg"] = np.array(model.ndcg, dtype=np.float64)
        out[tag + "_hit_new_user"] = np.array(model.hit_new_user, dtype=np.float64)
        out[tag + "_hit_new_item"] = np.array(model.hit_new_item, dtype=np.float64)
        out[tag + "_test_num"] = np.array(model.test_num, dtype=np.int64)
        out[tag + "_n_logs"] = np.int64(len(log))
        for n, rec in enumerate(log):
            out["%s_log%d" % (tag, n)] = np.array(rec, dtype=np.int32).reshape(-1, 4)
        return model, ds

    cls.__init__, cls.__getitem__ = rec_init, rec_get
    torch.utils.data.DataLoader = dl0
    save = torch.save
    torch.save = lambda *a, **k: None
    try:
        run("fine", "fine")
        run("full", "full")
        model, ds = run(None, "base", base=True)
        # ---- rank-weighted sampling on the trained model (SPMF pieces that do run) ----
        data = np.concatenate([periods[0][0], periods[1][0]]).astype(np.int64)
        p = model.compute_R_W_P(data)
        out["rwp_data"] = data; out["rwp_p"] = p
        model.all_item = np.unique(data[:, 1]); model.user_hit = None
        model.user_hit_num_in_W_R(data)
        np.random.seed(77)
        bu, bi, bn = model.sample_batch(data, 48, p, 1)
        out["sb_user"] = bu; out["sb_item"] = bi; out["sb_neg"] = bn
    finally:
        torch.utils.data.DataLoader = DL
        torch.save = save
        cls.__init__, cls.__getitem__ = cls_init, cls_get
    npz("baseline.npz", **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(4)
    ns = ref_harness.load()
    gens = dict(transfer_fwd=gen_transfer_fwd, run_mf=gen_run_mf, mf_steps=gen_mf_steps, tr_steps=gen_tr_steps,
                eval=gen_eval, period_run=gen_period_run, period_run_stop=lambda n: gen_period_run(n, stop=True),
                period_run_news=lambda n: gen_period_run(n, news=True), period_run_opts=lambda n: gen_period_run(n, opts=True),
                select_neg=gen_select_neg, baseline=gen_baseline)
    for name, fn in gens.items():
        if a.only and name not in a.only.split(","):
            continue
        fn(ns)


if __name__ == "__main__":
    main()
