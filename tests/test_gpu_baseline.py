"""MF baselines (sml_b200/model/baseline.py) against the reference's own runs of model/baseline.py
(tests/golden/baseline.npz, written by oracle/gen_golden.py from the unmodified reference): Reservious, StreamingData
(host logic, CPU) and, on the GPU, fine-tune / full-retrain / base_train with the recorded batches replayed and with the
reference's RNG streams re-derived, the rank-weighted sampling pieces, and a smoke run of the SPMF method (which the
reference cannot run: model/baseline.py:250)."""
import argparse
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "baseline.npz"))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def _write_stream(g, tmp_path):
    root = str(tmp_path) + "/mini/"
    os.makedirs(root + "train"); os.makedirs(root + "test")
    n = int(g["n_periods"])
    tot = 0
    for p in range(n):
        np.save(root + "train/%d.npy" % p, g["train%d" % p].astype(np.int64)); np.save(root + "test/%d.npy" % p, g["test%d" % p].astype(np.int64))
        tot += len(g["train%d" % p])
    np.save(root + "information.npy", np.array([tot, int(g["U"]), int(g["I"])], dtype=np.int64))
    np.save(root + "test_new_user.npy", g["new_user"]); np.save(root + "test_new_item.npy", g["new_item"])
    return root


def _args(**over):
    a = argparse.Namespace(lr=0.01, l2_u=1e-3, l2_i=1e-3, epochs=2, batch_size=32, laten_dim=64, neg_num=1, pool_size=0, laten=64,
                           cuda=0, method="full", pool_init_type=0, save_dir=None)
    for k, v in over.items():
        setattr(a, k, v)
    return a


# ------------------------------------------------------------------ host logic (CPU)
def test_reservious_matches_reference(g):
    from sml_b200.model.baseline import Reservious
    np.random.seed(11)
    r = Reservious(50)
    for k in range(4):
        r.updata(g["res_feed%d" % k])
        assert np.array_equal(r.pool, g["res_pool%d" % k]), k
        assert [r.t, r.pool_have] == g["res_state%d" % k].tolist(), k
    r2 = Reservious(30)
    r2.init_pool(g["res_feed2"])
    assert np.array_equal(r2.pool, g["res_init_pool"]) and [r2.t, r2.pool_have] == g["res_init_state"].tolist()
    assert np.random.rand() == float(g["res_after_draw"])          # same consumption of the global generator


def test_streaming_data_state_machine(g, tmp_path):
    from sml_b200.model.baseline import StreamingData
    root = _write_stream(g, tmp_path)
    ds = StreamingData(root)
    assert (int(ds.user_num), int(ds.item_num)) == (int(g["U"]), int(g["I"]))
    tr, te = ds.get_next(3, types="only_new")
    assert np.array_equal(tr, g["train2"]) and np.array_equal(te, g["test3"]) and tr.dtype == np.int64
    tr, te = ds.get_next(3, types="not_only_new")
    assert np.array_equal(tr, np.concatenate([g["train0"], g["train1"], g["train2"]]))
    n = int(g["n_periods"])
    assert ds.get_next(n, types="only_new") == (None, None)        # no test file for the period after the last one
    assert ds.get_next(n + 1, types="only_new") == (None, None)


# ------------------------------------------------------------------ device path
def _model(g, tmp_path, dev, tag, **kw):
    from sml_b200.model.baseline import SPMF, StreamingData
    ds = StreamingData(_write_stream(g, tmp_path))
    torch.manual_seed(2000); np.random.seed(2002)
    args = kw.pop("args", _args())
    m = SPMF(args, ds, int(ds.user_num), int(ds.item_num), 64, device=dev, **kw)
    # the constructor draws the tables from the global torch generator exactly like the reference (model/baseline.py:108)
    assert np.array_equal(m.MFbase.user_laten.weight.detach().cpu().numpy(), g[tag + "_init_user"])
    assert np.array_equal(m.MFbase.item_laten.weight.detach().cpu().numpy(), g[tag + "_init_item"])
    return m, ds


def _replay(g, tag, first_stage):
    def src(stage_id, epoch, n_rows):
        rec = g["%s_log%d" % (tag, stage_id - first_stage)]
        r = rec[epoch * n_rows:(epoch + 1) * n_rows]
        assert len(r) == n_rows
        return r[:, 1].astype(np.int64), r[:, 2].astype(np.int64), r[:, 3].astype(np.int64)
    return src


def _check_run(g, m, tag):
    m._flush()
    fu, fi = m.MFbase.user_laten.weight.detach().cpu().numpy(), m.MFbase.item_laten.weight.detach().cpu().numpy()
    # ~25-70 optimizer steps; fp32 Adam amplifies last-bit differences of the dots: 1e-4 after updates (north_star)
    assert np.abs(fu - g[tag + "_final_user"]).max() < 2e-4 and np.abs(fi - g[tag + "_final_item"]).max() < 2e-4
    assert m.test_num == g[tag + "_test_num"].tolist()
    assert np.allclose(np.array(m.recall), g[tag + "_recall"], atol=1.5 / min(m.test_num))      # at most one row near a tie
    assert np.allclose(np.array(m.ndcg, dtype=np.float64), g[tag + "_ndcg"], atol=1.0 / min(m.test_num))
    assert np.allclose(np.array(m.hit_new_user), g[tag + "_hit_new_user"], atol=1.5 / min(m.test_num))
    assert np.allclose(np.array(m.hit_new_item), g[tag + "_hit_new_item"], atol=1.5 / min(m.test_num))


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["fine", "full"])
@pytest.mark.parametrize("mode", ["replay", "rng"])
def test_fine_tune_and_full_retrain_match_reference(g, tmp_path, dev, method, mode):
    """SPMF.run(2, method) = run_one_stage2 per period (model/baseline.py:306-386, 505-556): recorded batches replayed, and
    batches re-derived from the global RNG streams (bit-identical triples => same trajectory)."""
    kw = dict(batch_source=_replay(g, method, 2)) if mode == "replay" else dict(emulate_reference_rng=True)
    m, _ = _model(g, tmp_path, dev, method, **kw)
    losses = []
    orig = m.run_one_stage2

    def wrapped(*a, **k):
        f = orig(*a, **k)
        if f:
            losses.extend(m.stage_losses)
        return f
    m.run_one_stage2 = wrapped
    m.run(2, method=method)
    _check_run(g, m, method)
    assert len(losses) == len(g[method + "_losses4"])
    assert np.abs(np.array(losses) - g[method + "_losses4"]).max() < 1.5e-4        # the reference prints 4 decimals
    assert set(m.summary) >= {"val_recall", "test_recall", "recall", "ndcg"}


@pytest.mark.gpu
def test_base_train_matches_reference(g, tmp_path, dev):
    """The pre-training loop (model/baseline.py:161-225) with different l2_u / l2_i, three epochs, recorded batches."""
    m, ds = _model(g, tmp_path, dev, "base", batch_source=_replay(g, "base", 3))
    m.base_train(3, 3, 1e-3, 2e-3)
    assert np.abs(np.array(m.base_losses) - g["base_losses4"]).max() < 1.5e-4
    te = ds.get_next(3)[1]
    m.recall = [m.test(te)[0]]; m.ndcg = [m.test(te)[1]]; m.test_num = []
    fu = m.MFbase.user_laten.weight.detach().cpu().numpy()
    assert np.abs(fu - g["base_final_user"]).max() < 2e-4
    assert np.allclose(m.recall[0], g["base_recall"][0], atol=1.5 / len(te)) and np.allclose(m.ndcg[0], g["base_ndcg"][0], atol=1.0 / len(te))


@pytest.mark.gpu
def test_rank_weighted_sampling_matches_reference(g, tmp_path, dev):
    """compute_R_W_P (:448-476) on the reference's trained tables and sample_batch (:489-503) with the same numpy seed."""
    m, _ = _model(g, tmp_path, dev, "base")
    m.MFbase.user_laten.weight.data.copy_(torch.from_numpy(g["base_final_user"]).to(dev))
    m.MFbase.item_laten.weight.data.copy_(torch.from_numpy(g["base_final_item"]).to(dev))
    data = g["rwp_data"]
    p = m.compute_R_W_P(data)
    assert p.shape == g["rwp_p"].shape and abs(p.sum() - 1.0) < 1e-5
    # rank-based weights: the same multiset, and the same entry for every interaction that occurs once (an interaction that is
    # repeated in the data has tied scores, and argsort may order ties either way -- on the reference's CPU / CUDA paths too)
    assert np.abs(np.sort(p) - np.sort(g["rwp_p"])).max() < 1e-8
    key = data[:, 0] * 100000 + data[:, 1]
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    once = cnt[inv] == 1
    assert once.mean() > 0.5 and np.mean(np.abs(p - g["rwp_p"])[once] < 1e-8) > 0.98
    m.all_item = np.unique(data[:, 1]); m.user_hit = None
    m.user_hit_num_in_W_R(data)
    np.random.seed(77)
    bu, bi, bn = m.sample_batch(data, 48, g["rwp_p"], 1)
    assert np.array_equal(bu, g["sb_user"]) and np.array_equal(bi, g["sb_item"]) and np.array_equal(bn, g["sb_neg"])


@pytest.mark.gpu
def test_spmf_method_runs_and_learns(g, tmp_path, dev):
    """The SPMF method proper (reservoir + rank-weighted sampling, model/baseline.py:227-304).  The reference raises in its
    first period (:250), so there is no run to compare with: the reservoir bookkeeping follows Reservious (checked above), the
    training loss must fall and the reservoir must fill."""
    m, ds = _model(g, tmp_path, dev, "fine", args=_args(pool_size=60, epochs=3, method="spmf"))
    m.base_train_not_train(1)
    assert m.Reservious.pool_have > 0
    before = m.test(ds.get_next(2)[1])[0][-1]
    m.run(2, method="spmf")
    assert len(m.recall) == int(g["n_periods"]) - 2 and m.Reservious.t > 60
    train = np.concatenate([g["train%d" % p] for p in range(int(g["n_periods"]))]).astype(np.int64)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    from sml_b200 import ops
    m._flush()
    s = ops.pair_scores(m.MFbase.user_laten.weight.data, m.MFbase.item_laten.weight.data, T(train[:, 0]), T(train[:, 1]))
    assert float(torch.sigmoid(s).mean()) > 0.5 and np.isfinite(before)      # the positives it trained on score above chance
