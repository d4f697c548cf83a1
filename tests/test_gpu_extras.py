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
    assert (a.transfer.theta - b.transfer.theta).abs().max().item() < 1e-5
    assert a.MF_optimizer.step_count == b.MF_optimizer.step_count and a.recall == b.recall


def test_device_sampler_run(golden, tmp_path):
    """Throughput mode: batches sampled on the GPU.  Not bit-comparable with the reference (different RNG), but the
    loop must train: finite tables and a validation recall in the range of the reference's run."""
    g = golden("period_run")
    args, m = _make_meta(g, str(tmp_path), device_sampler=True, emulate_reference_rng=False)
    m.run(args)
    assert torch.isfinite(m.MFbase.user_laten.weight.data).all() and torch.isfinite(m.transfer.theta).all()
    assert len(m.recall) == 3 and abs(np.mean(m.recall) - np.mean(g["recall"])) < 0.15


def test_baseline_trainers_learn():
    """Fine-tune / full-retrain baselines on the plain-MF kernels: the loss falls and the positives move up."""
    from sml_b200.data import synth
    from sml_b200.model.baseline import FineTune, FullRetrain
    periods = synth.make_stream(300, 400, 3000, 3, n_neg=50, seed=4)
    for cls in (FineTune, FullRetrain):
        torch.manual_seed(0)
        t = cls(300, 400, lr=0.01, l2_u=1e-5, l2_i=1e-5, batch_size=256)
        r0, _ = t.test(periods[1][1], topK=10)
        first = t.run_period(periods[0][0], epochs=1)[0]
        last = t.run_period(periods[1][0], epochs=6)[-1]
        r1, _ = t.test(periods[1][1], topK=10)
        assert last < first and r1 > r0 + 0.05, (cls.__name__, first, last, r0, r1)
