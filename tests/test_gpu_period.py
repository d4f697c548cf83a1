"""GPU: the whole period loop (meta_train.run) on a tiny stream against the UNMODIFIED reference's
run recorded in tests/golden/period_run*.npz -- same triples (replayed from the recording, and
re-derived through the RNG emulation), same state machine, tables / theta / metrics within the
stated tolerance after ~100 optimizer steps."""
import numpy as np
import pytest
import torch

from tests.helpers import theta_from_chk, sample
from tests.test_host_logic import make_args, write_fixture_stream

pytestmark = pytest.mark.gpu


def run_ours(g, tmp, stop, replay, news=False, opts=False, defer=True):
    from sml_b200.data.dataset2 import transfer_data
    from sml_b200.model.transfer import meta_train
    NP, U, I = write_fixture_stream(g, tmp)
    args = make_args(g, tmp, stop, news=news, opts=opts)
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
    ds = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)],
                       test_list=[str(j) for j in range(5, NP)], validation_list=None, online_train_time=2, online_test_time=5)
    src = None
    if replay:
        kinds = [str(k) for k in g["log_kinds"]]
        epochs = []
        for n, kind in enumerate(kinds):
            rec = g["log%d" % n]
            for s in range(0, len(rec), 96):
                epochs.append((kind, rec[s:s + 96, 1:4].astype(np.int64)))
        it = iter(epochs)

        def src(kind, stage_id, epoch, n_rows):
            k, tri = next(it)
            assert k == kind and len(tri) == n_rows
            return tri[:, 0].copy(), tri[:, 1].copy(), tri[:, 2].copy()
    meta = meta_train(args, ds, U, I, 64, batch_source=src)
    meta.defer = defer
    tu, ti = theta_from_chk(g["theta_com"])
    sd = {("user_transfer." + k): torch.from_numpy(v) for k, v in tu.items()}
    sd.update({("item_transfer." + k): torch.from_numpy(v) for k, v in ti.items()})
    meta.transfer.load_state_dict(sd)
    meta.eval_log = []
    inner = meta._eval

    def logged(test_set, topK):
        h = inner(test_set, topK)                  # handle of the deferred (recall, ndcg); read back at the end of the period
        meta._later(lambda v, h=h, k=float(topK), n=float(len(test_set)): meta.eval_log.append([k, n, float(v[h][0]), float(v[h][1])]))
        return h
    meta._eval = logged
    meta.run(args)
    return meta


@pytest.mark.parametrize("name,stop", [("period_run", False), ("period_run_stop", True), ("period_run_news", False)])
@pytest.mark.parametrize("replay", [True, False])
def test_period_run_matches_reference(golden, tmp_path, name, stop, replay):
    g = golden(name)
    meta = run_ours(g, str(tmp_path), stop, replay, news=name.endswith("news"), opts=name.endswith("opts"))
    fu = meta.MFbase.user_laten.weight.data.cpu().numpy()
    fi = meta.MFbase.item_laten.weight.data.cpu().numpy()
    # Differences against the reference grow ~10x per period on these tiny streams (chaotic training dynamics: the fp32
    # SIMT path shows the same growth from a 10x smaller start), so the tight check is the
    # north_star one -- 1e-4 after ONE period's updates, test_first_period_within_tolerance -- and the end-of-stream check
    # is loose.
    loose = 3e-3 if not name.endswith("news") else 5e-2
    scale = np.abs(g["final_user"]).max()
    assert np.abs(fu - g["final_user"]).max() < loose * scale, np.abs(fu - g["final_user"]).max()
    # end-of-stream item table, element-wise relative to its scale: measured 0.5e-4 .. 2.5e-4 (Yelp-like, 3 periods of drift),
    # 1.1e-5 .. 1.4e-5 (transfer frozen) and 1.1e-2 .. 1.3e-2 (news-like: 2 + 2 epochs on a churning item set amplify ~10x per
    # period) over repeated runs; the bounds sit ~4x above
    item_tol = {"period_run": 1e-3, "period_run_stop": 1e-4, "period_run_news": 5e-2, "period_run_opts": 1e-3}[name]
    assert np.abs(fi - g["final_item"]).max() < item_tol * np.abs(g["final_item"]).max()
    assert np.abs(meta.user_weight_hat.cpu().numpy() - g["final_user_hat"]).max() < loose * max(1.0, np.abs(g["final_user_hat"]).max())
    for net in ("user", "item"):
        mod = getattr(meta.transfer, net + "_transfer")
        for k in ("conv1.weight", "conv2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"):
            a, b = k.split(".")
            got = getattr(getattr(mod, a), b).detach().cpu().numpy()
            ref = g["tF.%s.%s" % (net, k)]
            assert np.abs(sample(got).reshape(ref.shape) - ref).max() < (2e-4 if not name.endswith("news") else 2e-3), (net, k)
    assert meta.test_num == [int(x) for x in g["test_num"]]
    for key in ("recall", "recall_10", "recall_5"):
        got = np.array([float(x) for x in getattr(meta, key)])
        # 96 test rows per period: allow one borderline row (score gaps of a few ulp) to flip
        assert np.abs(got - g[key]).max() <= (1.0 if not name.endswith("news") else 3.0) / 96 + 1e-9, (key, got, g[key])
    for key in ("ndcg", "ndcg_10", "ndcg_5"):
        got = np.array([float(x) for x in getattr(meta, key)])
        assert np.abs(got - g[key]).max() < (1.2e-2 if not name.endswith("news") else 3e-2), (key, got, g[key])
    # EVERY test_model call of the reference's run (validation passes before / after each MF epoch and each transfer epoch,
    # model/transfer.py:444-446,517-519,684-686,738-741, and the real tests): same sequence of (K, rows), metrics within one
    # borderline row.  A stale (re-used but outdated) scoring pass shows up here as a repeated value.
    ours, ref = np.array(meta.eval_log), g["eval_log"]
    assert ours.shape == ref.shape, (ours.shape, ref.shape)
    assert np.array_equal(ours[:, :2], ref[:, :2])
    flips = 1.0 if not name.endswith("news") else 3.0
    assert np.abs(ours[:, 2] - ref[:, 2]).max() <= flips / 96 + 1e-9, np.abs(ours[:, 2] - ref[:, 2]).max()
    assert np.abs(ours[:, 3] - ref[:, 3]).max() < (1.2e-2 if not name.endswith("news") else 3e-2)
    first = slice(0, 8)         # the first period's validation passes: before any drift, exact recall
    assert np.array_equal(np.round(ours[first, 2] * 96), np.round(ref[first, 2] * 96)), (ours[first, 2], ref[first, 2])
    assert np.abs(ours[first, 3] - ref[first, 3]).max() < 1e-4
    # scoring passes really launched: everything except the repeats on unchanged tables (the "before MF" pass of outer
    # phases >= 1, and K = 10 / 5 / the following "before transfer" pass after a real test)
    assert meta.eval_passes["scored"] + meta.eval_passes["reused"] == len(ref)
    # (a change of the reference's metrics between two consecutive calls at the same K proves the tables or the file changed)
    changed = 1 + int(np.sum(np.any(ref[1:, 2:] != ref[:-1, 2:], axis=1) & (ref[1:, 0] == ref[:-1, 0])))
    assert meta.eval_passes["scored"] >= changed, (meta.eval_passes, changed)
    # Adam step counters: one per optimizer step, surviving across periods (model/transfer.py:764)
    n_mf = sum(len(g["log%d" % n]) for n, k in enumerate(g["log_kinds"]) if str(k) == "MF") // 96 * 3
    assert meta.MF_optimizer.step_count == n_mf


@pytest.mark.parametrize("name", ["period_run", "period_run_news", "period_run_opts"])
def test_first_period_within_tolerance(golden, tmp_path, name):
    """north_star: fp32 embeddings and theta within 1e-4 (relative) after one period's updates.  The fixture stores
    sum / abs-sum checksums of both tables and of theta after every period of the reference's run.
    ``period_run_opts`` is the reference's run with --need_adaptive --clip_grad (maxnorm 0.05) --norm: the options that are
    off by default (model/transfer.py:490-499,507-510,723-727).  With them the mini stream collapses every user row onto
    the same vector within three periods and the end state is hypersensitive (two runs of this package differ by 0.3 there),
    so the options are pinned on the first period only."""
    import contextlib, io
    from sml_b200.data.dataset2 import transfer_data
    from sml_b200.model.transfer import meta_train
    g = golden(name)
    tmp = str(tmp_path)
    NP, U, I = write_fixture_stream(g, tmp)
    args = make_args(g, tmp, False, news=name.endswith("news"), opts=name.endswith("opts"))
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
    ds = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)],
                       test_list=[str(j) for j in range(5, NP)], validation_list=None, online_train_time=2, online_test_time=5)
    meta = meta_train(args, ds, U, I, 64)
    tu, ti = theta_from_chk(g["theta_com"])
    sd = {("user_transfer." + k): torch.from_numpy(v) for k, v in tu.items()}
    sd.update({("item_transfer." + k): torch.from_numpy(v) for k, v in ti.items()})
    meta.transfer.load_state_dict(sd)
    assert meta.train_one_stage3(args, 0)
    uw = meta.MFbase.user_laten.weight.data.double(); iw = meta.MFbase.item_laten.weight.data.double()
    th = float(sum(p.double().abs().sum() for p in meta.transfer.parameters()))
    ref = g["stage_sums"][0]
    assert abs(float(uw.abs().sum()) - ref[1]) < 1e-4 * ref[1]
    assert abs(float(iw.abs().sum()) - ref[3]) < 1e-4 * ref[3]
    assert abs(th - ref[4]) < 1e-4 * ref[4]
    assert abs(float(uw.sum()) - ref[0]) < 1e-4 * ref[1] and abs(float(iw.sum()) - ref[2]) < 1e-4 * ref[3]


def test_blocking_reads_give_the_same_run(golden, tmp_path):
    """SML_DEFER=0 (a blocking read at every print, like the reference) and the default (one read per period) enqueue the
    same kernels in the same order: the same run up to the run-to-run noise of the fp32 atomics (split-K sums, scatter-adds),
    which Adam amplifies over the ~100 steps of the stream like it does between any two runs."""
    g = golden("period_run")
    a = run_ours(g, str(tmp_path / "a"), False, True, defer=True)
    b = run_ours(g, str(tmp_path / "b"), False, True, defer=False)
    for x, y in ((a.MFbase.user_laten.weight.data, b.MFbase.user_laten.weight.data), (a.MFbase.item_laten.weight.data, b.MFbase.item_laten.weight.data),
                 (a.transfer.theta, b.transfer.theta)):
        assert float((x - y).abs().max()) <= 2e-3 * float(y.abs().max())
    la, lb = np.array(a.eval_log), np.array(b.eval_log)
    assert la.shape == lb.shape == g["eval_log"].shape and np.array_equal(la[:, :2], lb[:, :2])       # same calls in the same order
    assert np.abs(la[:, 2] - lb[:, 2]).max() <= 2.0 / 96 + 1e-9 and np.abs(la[:, 3] - lb[:, 3]).max() <= 2e-2
    assert len(a.recall) == len(b.recall) == 3 and np.abs(np.array(a.recall) - np.array(b.recall)).max() <= 2.0 / 96 + 1e-9
    assert abs(a.last_MF_loss - b.last_MF_loss) <= 1e-3 * abs(b.last_MF_loss) and abs(a.last_TR_loss - b.last_TR_loss) <= 1e-3 * abs(b.last_TR_loss)
