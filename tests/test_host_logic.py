"""Host-side logic on CPU: the period state machine and the batch streams of sml_b200.meta_train
must consume the global RNGs exactly like the reference (--numworkers 0) and therefore produce
bit-identical (user, item, neg) triples -- checked against the batches recorded from the
UNMODIFIED reference by oracle/gen_golden.py.  Device work is replaced by a recording test double
(this is a test of the host logic only; the product class has no CPU path)."""
import argparse
import os

import numpy as np
import pytest
import torch

from sml_b200.data import synth
from sml_b200.data.batching import ReferenceStream
from sml_b200.data.dataset import offlineDataset_withsample
from sml_b200.data.dataset2 import transfer_data, trainDataset_withPreSample
from sml_b200.model.transfer import meta_train


def make_args(g, tmp, stop=False, news=False, opts=False):
    mfb, trb, multi, mfe, tre, seed = (int(x) for x in g["args"])
    lr, l2, trlr, trl2 = (float(x) for x in g["hyper"])
    a = _make_args(g, tmp, stop, news, mfb, trb, multi, mfe, tre, seed, lr, l2, trlr, trl2)
    if opts:        # oracle/gen_golden.py: gen_period_run(opts=True)
        a.need_adaptive = True; a.clip_grad = True; a.maxnorm_grad = 0.05; a.norm = True
    return a


def _make_args(g, tmp, stop, news, mfb, trb, multi, mfe, tre, seed, lr, l2, trlr, trl2):
    return argparse.Namespace(
        data_name="news" if news else "yelp", data_path=tmp + "/", multi_num=multi, MF_lr=lr, MF_epochs=mfe, l2=l2, MF_batch_size=mfb, laten=64,
        pre_model=os.path.join(tmp, "pre.pt"), MF_sample="all", Load_W_hat=False, clip_grad=False, need_adaptive=False,
        maxnorm_grad=3.0, TR_lr=trlr, TR_l2=trl2, TR_epochs=tre, TR_batch_size=trb, TR_sample_type="alone",
        TR_with_MF_bias=False, TR_stop_=stop, transfer_type="conv_com", seed=seed, numworkers=0, cuda=0, topK=20, pass_num=1,
        norm=False, Lambda_lr=0.01, min_l2=0.0001, set_t_as_tt=False, tqdm=False, need_writer=False, test_in_TR_Train=False)


def write_fixture_stream(g, tmp):
    NP, U, I = int(g["n_periods"]), int(g["U"]), int(g["I"])
    periods = [(g["train%d" % p].astype(np.int64), g["test%d" % p].astype(np.int64)) for p in range(NP)]
    synth.write_stream(tmp + "/", "mini", periods, U, I)
    sd = {"user_laten.weight": torch.from_numpy(g["pre_user"]), "item_laten.weight": torch.from_numpy(g["pre_item"]),
          "user_bais.weight": torch.zeros(U, 1), "item_bais.weight": torch.zeros(I, 1)}
    torch.save(sd, os.path.join(tmp, "pre.pt"))
    return NP, U, I


class HostProbe(meta_train):
    """meta_train with every device operation replaced by a recorder."""

    def _require_device(self, device):
        self.device = torch.device("cpu")
        self.rec = []

    def _test_set(self, arr):
        class _Arr(object):
            shape = arr.shape
        return _Arr()

    def _eval(self, test_set, topK):
        ReferenceStream.loader_iter()
        self.rec.append(("eval", test_set.shape[0], topK))
        return self._const_value((0.0, torch.tensor(0.0)))

    def _mf_epoch(self, args, triples):
        self.rec.append(("MF", np.stack(triples, 1)))
        return self._const_value(0.0)

    def _tr_epoch(self, args, triples):
        self.rec.append(("TR", np.stack(triples, 1)))
        return self._const_value(0.0)

    def updata(self):
        self.rec.append(("updata",))


@pytest.mark.parametrize("name,stop", [("period_run", False), ("period_run_stop", True), ("period_run_news", False)])
def test_batch_stream_matches_reference(golden, tmp_path, name, stop):
    g = golden(name)
    tmp = str(tmp_path)
    NP, U, I = write_fixture_stream(g, tmp)
    args = make_args(g, tmp, stop, news=name.endswith("news"))
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)       # main_yelp.py:137-140
    ds = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)],
                       test_list=[str(j) for j in range(5, NP)], validation_list=None, online_train_time=2, online_test_time=5)
    probe = HostProbe(args, ds, U, I, 64)
    probe.run(args)
    kinds = [str(k) for k in g["log_kinds"]]
    got = [e for e in probe.rec if e[0] in ("MF", "TR")]
    # one reference dataset instance may span several epochs (MF_epochs = 2 in the stop branch)
    ref = []
    for n, kind in enumerate(kinds):
        rec = g["log%d" % n]
        N = 96
        for s in range(0, len(rec), N):
            ref.append((kind, rec[s:s + N, 1:4].astype(np.int64)))
    assert [k for k, _ in got] == [k for k, _ in ref]
    for (k, a), (_, b) in zip(got, ref):
        assert np.array_equal(a, b), k                               # sampled indices: bit-exact
    assert len(probe.test_num) == 3 and probe.test_num == [96, 96, 96]
    n_updata = sum(1 for e in probe.rec if e[0] == "updata")
    assert n_updata > 0


def test_presample_column_semantics():
    """The shuffled negative-column list includes the positive column (column 1) and the column
    advances once per full pass (data/dataset2.py:181-200)."""
    np.random.seed(5)
    data = np.arange(7 * 6).reshape(7, 6)
    ds = trainDataset_withPreSample(data)
    assert sorted(ds.neg_flag.tolist()) == [1, 2, 3, 4, 5] and ds.neg_all == 4
    c0 = ds.current_column()
    for i in range(7):
        u, it, ng = ds[i]
        assert ng == data[i, c0]
    assert ds.used_neg_count == 1 and ds.current_column() == int(ds.neg_flag[1])


def test_alone_sampler_vectorised_equals_sequential():
    rng = np.random.default_rng(0)
    data = np.stack([rng.integers(0, 20, 300), rng.integers(0, 15, 300)], 1)
    order = rng.permutation(300)
    np.random.seed(9)
    ds = offlineDataset_withsample(data)
    assert np.array_equal(ds.item_all, np.unique(data[:, 1])) and ds.item_all.dtype == data.dtype
    seq = np.array([ds[i][2] for i in order])
    state_after_seq = np.random.get_state()[1].copy()
    np.random.seed(9)
    vec = ReferenceStream.alone_negatives(ds, order)
    assert np.array_equal(seq, vec)
    assert np.array_equal(state_after_seq, np.random.get_state()[1])   # same number of draws consumed
    assert not ds.interacted(ds.user[order], vec).any()


def test_next_train_branches(golden, tmp_path):
    g = golden("period_run")
    tmp = str(tmp_path)
    NP, U, I = write_fixture_stream(g, tmp)
    args = make_args(g, tmp)
    ds = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)],
                       test_list=[str(j) for j in range(5, NP)], online_train_time=2, online_test_time=5)
    assert int(ds.user_number) == U and int(ds.item_number) == I
    a = ds.next_train(0)
    assert a[2] is None and a[1].shape[1] == 2 and a[0].shape[1] == 42      # train-only branch: set_tt = train file
    b = ds.next_train(2)
    assert b[2] is not None and np.array_equal(b[2], g["test5"])            # test+train branch
    assert ds.next_train(5) == (None, None, None, None)                      # end of stream
    args.TR_stop_ = True
    ds2 = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)],
                        test_list=[str(j) for j in range(5, NP)], online_train_time=2, online_test_time=5)
    c = ds2.next_train(2)
    assert c[1] is None and c[3] is c[2]                                     # stop-transfer branch


def test_synth_stream_layout(tmp_path):
    periods = synth.make_stream(50, 80, 40, 3, n_neg=20, seed=1)
    root = synth.write_stream(str(tmp_path) + "/", "s", periods, 50, 80)
    info = np.load(os.path.join(root, "information.npy"))
    assert info.tolist() == [120, 50, 80]
    te = np.load(os.path.join(root, "test", "1.npy"))
    assert te.shape == (40, 22) and te.dtype == np.int64
    for row in te:
        assert len(set(row[2:].tolist())) == 20 and row[1] not in row[2:]
    churn = synth.make_stream(60, 400, 50, 4, n_neg=20, seed=2, churn=0.5)
    assert all(t.shape == (50, 22) for _, t in churn)


def test_cli_flags_match_reference():
    """Flag names and defaults of main_yelp.py:10-120 / main_news.py:8-115 (recorded from the reference's parsers)."""
    import main_yelp
    import main_news
    ref_yelp = {'Lambda_lr': 0.01, 'Load_W_hat': False, 'MF_batch_size': 1024, 'MF_epochs': 1, 'MF_lr': 0.01, 'MF_sample': 'all',
                'TR_batch_size': 256, 'TR_epochs': 1, 'TR_l2': 0.0001, 'TR_lr': 0.001, 'TR_sample_type': 'alone', 'TR_stop_': False,
                'TR_with_MF_bias': False, 'clip_grad': False, 'cuda': 0, 'data_name': 'yelp', 'data_path': '/home/sml/dataset/',
                'l2': 1e-06, 'laten': 64, 'maxnorm_grad': 3.0, 'min_l2': 0.0001, 'multi_num': 10, 'need_adaptive': False,
                'need_writer': False, 'norm': False, 'numworkers': 4, 'pass_num': 1,
                'pre_model': '/home/sml/save_model/sml/yelp/BCE_init.pkl', 'seed': 2000, 'set_t_as_tt': False,
                'test_in_TR_Train': False, 'topK': 20, 'tqdm': False, 'transfer_type': 'conv_com'}
    ref_news = dict(ref_yelp, MF_epochs=2, TR_epochs=2, multi_num=7, data_name='news', pre_model='/home/sml/save_model/sml/news/BCE_init.pkl')
    for mod, ref in ((main_yelp, ref_yelp), (main_news, ref_news)):
        got = vars(mod.get_parse().parse_args([]))
        for k, v in ref.items():
            assert got[k] == v, (mod.__name__, k, got[k], v)
    a = main_yelp.get_parse().parse_args(["--MF_epochs=3", "--TR_stop_", "yes"])
    assert a.MF_epochs == 3 and a.TR_stop_ is True


def test_load_pre_model_accepts_reference_pickle(tmp_path):
    """The reference pickles a whole ``model.MF.MFbasemode`` (model/transfer.py:322-325).  A file pickled under that module
    path must load into this package's class even though ``model.MF`` does not exist here."""
    import sys
    import types
    from sml_b200.model import MF as ours
    from sml_b200.model.transfer import load_pre_model
    # pickle under the reference's module path
    pkg = types.ModuleType("model"); sub = types.ModuleType("model.MF")
    Fake = type("MFbasemode", (ours.MFbasemode,), {"__module__": "model.MF"})
    sub.MFbasemode = Fake; pkg.MF = sub
    sys.modules["model"] = pkg; sys.modules["model.MF"] = sub
    try:
        torch.manual_seed(1)
        m = Fake(7, 9, 64)
        path = str(tmp_path / "pre.pkl")
        torch.save(m, path)
    finally:
        del sys.modules["model"], sys.modules["model.MF"]
    got = load_pre_model(path, 7, 9, 64, "cpu")
    assert type(got) is ours.MFbasemode
    assert torch.equal(got.user_laten.weight, m.user_laten.weight) and got.item_num == 9
    # a plain state_dict works too and does not disturb the global generator
    torch.save(got.state_dict(), str(tmp_path / "sd.pt"))
    torch.manual_seed(5); a = torch.rand(1)
    torch.manual_seed(5); got2 = load_pre_model(str(tmp_path / "sd.pt"), 7, 9, 64, "cpu"); b = torch.rand(1)
    assert torch.equal(a, b) and torch.equal(got2.item_laten.weight, m.item_laten.weight)


def test_memory_stream_branches():
    from sml_b200.data.memory_stream import MemoryStream
    periods = [(np.zeros((3, 2), np.int64) + p, np.zeros((3, 5), np.int64) + p) for p in range(6)]
    ds = MemoryStream(periods, 10, 20, online_train_time=1, online_test_time=4)
    set_t, set_tt, now_test, val = ds.next_train(0)            # now = 1: train-only branch
    assert now_test is None and set_t[0, 0] == 1 and set_tt[0, 0] == 2 and val[0, 0] == 2 and set_tt.shape[1] == 2
    set_t, set_tt, now_test, val = ds.next_train(2)            # now = 3, next = 4 = first test period
    assert now_test[0, 0] == 4 and val[0, 0] == 4 and set_t[0, 0] == 3
    assert ds.next_train(4) == (None, None, None, None)
    ds.reinit(); assert ds.test_count == 0
    stop = MemoryStream(periods, 10, 20, online_train_time=1, online_test_time=4, tr_stop=True)
    assert stop.next_train(2)[1] is None


def test_select_neg_forinteraction_matches_reference(golden, tmp_path):
    """Test-file builder (data/dataset2.py:356-414): same files as the reference for the same np.random.seed."""
    from sml_b200.data.dataset2 import select_neg_forinteraction
    g = golden("select_neg")
    base = tmp_path / "mini"
    base.mkdir()
    for k in range(5):
        np.save(str(base / ("%d.npy" % k)), g["file%d" % k])
    np.random.seed(int(g["seed"]))
    out = select_neg_forinteraction(path=str(tmp_path) + "/", datasetname="mini", file_path_list=[str(k) for k in range(5)],
                                    leave_for_init_train=float(g["leave"]), neg_num=int(g["neg_num"]))
    assert len(out) == 2
    for k, t in zip((3, 4), out):
        assert np.array_equal(t, g["test%d" % k])
        assert np.array_equal(np.load(str(base / "test" / ("%d.npy" % k))), g["test%d" % k])
        # the defining properties: negatives distinct, never the row's own history
        assert all(len(set(r[2:])) == int(g["neg_num"]) and r[1] not in r[2:] for r in t)


def test_eval_cache_never_serves_stale_ranks(monkeypatch):
    """meta_train._eval (CPU: the scoring pass is stubbed).  A kept rank pass may only be reused when the same device
    file is evaluated again with NO table write in between: the reference re-scores every time
    (model/transfer.py:444-446,517-519,684-686,738-741).  Round-1 bug: a cache hit left ``frozen`` set, so the
    after-epoch evaluation of every outer phase >= 1 returned the pre-epoch ranks (9 of 31 passes per period)."""
    import types
    import torch
    from sml_b200.model import transfer as T
    from sml_b200.evalution.evaluation2 import DeviceTestSet
    from sml_b200.profiling import EventTimers

    m = T.meta_train.__new__(T.meta_train)
    w = lambda: types.SimpleNamespace(weight=torch.nn.Parameter(torch.zeros(2, 2)))
    m.MFbase = types.SimpleNamespace(user_laten=w(), item_laten=w(), eval=lambda: None)
    m._tab_version, m._eval_cache = 0, None
    m.defer, m._pending, m._resolved, m._handle, m._later_q, m._stage_depth = True, {}, {}, 0, [], 0
    m.eval_passes = dict(scored=0, reused=0)
    m.events = EventTimers(False)
    m.timers = dict(eval=0.0)
    seen = []

    def fake_score_sums(ts, topK):
        # what DeviceTestSet.ranks does, with the table version standing in for the scores
        if not (ts.frozen and ts._rank_cache is not None):
            ts._rank_cache = ("ranks@", m._tab_version)
        seen.append(ts._rank_cache[1])
        return torch.zeros(2)
    m._score_sums = fake_score_sums

    def file_set(rows):
        ts = DeviceTestSet.__new__(DeviceTestSet)
        ts.rows, ts.frozen, ts._rank_cache, ts.emulate_reference_rng, ts.batch = rows, False, None, False, 1024
        return ts
    val_rows, test_rows = torch.zeros(3, 4, dtype=torch.int64), torch.zeros(3, 4, dtype=torch.int64)
    multi_num = 10
    for phase in range(multi_num):                      # one period of train_one_stage3, MF_epochs = TR_epochs = 1
        val = file_set(val_rows)                        # MF_train_onestage builds one DeviceTestSet for before + after
        m._eval(val, 20); assert seen[-1] == m._tab_version
        m._tab_version += 1                             # _mf_epoch
        m._eval(val, 20); assert seen[-1] == m._tab_version, "after-epoch evaluation served pre-epoch ranks"
        m._tab_version += 1                             # updata
        if phase == 0:                                  # the real test: three K on unchanged tables
            for K in (20, 10, 5):
                m._eval(file_set(test_rows), K); assert seen[-1] == m._tab_version
        now = file_set(val_rows)                        # transfer_train_onestage
        m._eval(now, 20); assert seen[-1] == m._tab_version
        m._tab_version += 1                             # updata after the transfer epoch
        m._eval(now, 20); assert seen[-1] == m._tab_version
    # 40 validation passes of which only the 9 "before MF" repeats of phases >= 1 may be reused, + 1 of the 3 real tests
    assert m.eval_passes == dict(scored=31 + 1, reused=9 + 2), m.eval_passes
    # a write torch can see (in-place op on the Parameter) and an explicit invalidate() both force a new pass
    ts = file_set(val_rows)
    m._eval(ts, 20); assert m.eval_passes["reused"] == 12
    with torch.no_grad():
        m.MFbase.user_laten.weight.add_(1.0)
    m._eval(ts, 20); assert m.eval_passes["scored"] == 33
    m.invalidate()
    m._eval(ts, 20); assert m.eval_passes["scored"] == 34


def test_deferred_reads_keep_the_reference_print_order(golden, tmp_path, capsys):
    """meta_train reads losses / metrics back once per period (flush_deferred): the period's prints, list appends and loss
    attributes must come out exactly as with a blocking read at every print (SML_DEFER=0, the reference's behaviour), in the
    same order, and nothing may be left queued when train_one_stage3 returns."""
    g = golden("period_run")
    tmp = str(tmp_path)
    NP, U, I = write_fixture_stream(g, tmp)
    outs = []
    for defer in (True, False):
        args = make_args(g, tmp, False)
        torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
        ds = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)],
                           test_list=[str(j) for j in range(5, NP)], validation_list=None, online_train_time=2, online_test_time=5)
        probe = HostProbe(args, ds, U, I, 64)
        probe.defer = defer
        scalars = []

        class Writer(object):                                  # stands in for the TensorBoard writer of --need_writer
            def add_scalar(self, name, value, itr):
                scalars.append((name, round(float(value), 6), itr))
        probe.need_writer, probe.writer = True, Writer()
        n = [0]
        ev, mf, tr = probe._eval, probe._mf_epoch, probe._tr_epoch

        def eval_(test_set, topK, ev=ev, n=n):                # distinct values so that a swapped pair would show
            ev(test_set, topK)
            n[0] += 1
            return probe._const_value((n[0] / 1000.0, torch.tensor(n[0] / 500.0)))

        def mf_(a, t, mf=mf, n=n):
            mf(a, t)
            n[0] += 1
            return probe._const_value(n[0] * 1.5)

        def tr_(a, t, tr=tr, n=n):
            tr(a, t)
            n[0] += 1
            return probe._const_value(n[0] * 2.5)
        probe._eval, probe._mf_epoch, probe._tr_epoch = eval_, mf_, tr_
        capsys.readouterr()
        stage, seen = 0, []
        while probe.train_one_stage3(args, stage):
            assert not probe._pending and not probe._later_q          # flushed at the end of every period
            seen.append((len(probe.recall), getattr(probe, "last_MF_loss", None), getattr(probe, "last_TR_loss", None)))
            stage += 1
        text = capsys.readouterr().out
        # dataset constructors print while they run ("user max: ..."): not part of the deferred stream
        lines = [ln for ln in text.splitlines() if not ln.startswith("user max") and "time cost" not in ln]
        outs.append((lines, seen, [float(x) for x in probe.recall], [float(x) for x in probe.ndcg_5], scalars))
    assert outs[0][0] == outs[1][0] and len(outs[0][0]) > 50
    assert outs[0][1:] == outs[1][1:] and len(outs[0][2]) == 3
    # the writer sees the reference's scalars (model/transfer.py:447-449,520-527,686-690,716,741-744) with the right iteration counters
    sc = outs[0][4]
    names = {n for n, _, _ in sc}
    assert names == {"Acc/MF-recall20", "Acc/MF-ndcg20", "norm/user-norm", "Loss/MF-loss", "Acc/tr-TR-recall@20", "Acc/tr-TR-ndcg@20", "Loss/TR-loss"}
    mf_itrs = [i for n, _, i in sc if n == "Loss/MF-loss"]
    assert mf_itrs == sorted(mf_itrs) and len(set(mf_itrs)) == len(mf_itrs)
    tr_loss = [v for n, v, _ in sc if n == "Loss/TR-loss"]
    assert all(abs(v * 16 / 2.5 - round(v * 16 / 2.5)) < 1e-3 for v in tr_loss)      # = (2.5 x call number) / TR_batch_size (16)
