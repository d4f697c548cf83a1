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


def run_ours(g, tmp, stop, replay, news=False):
    from sml_b200.data.dataset2 import transfer_data
    from sml_b200.model.transfer import meta_train
    NP, U, I = write_fixture_stream(g, tmp)
    args = make_args(g, tmp, stop, news=news)
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
    tu, ti = theta_from_chk(g["theta_com"])
    sd = {("user_transfer." + k): torch.from_numpy(v) for k, v in tu.items()}
    sd.update({("item_transfer." + k): torch.from_numpy(v) for k, v in ti.items()})
    meta.transfer.load_state_dict(sd)
    meta.run(args)
    return meta


@pytest.mark.parametrize("name,stop", [("period_run", False), ("period_run_stop", True), ("period_run_news", False)])
@pytest.mark.parametrize("replay", [True, False])
def test_period_run_matches_reference(golden, tmp_path, name, stop, replay):
    g = golden(name)
    meta = run_ours(g, str(tmp_path), stop, replay, news=name.endswith("news"))
    fu = meta.MFbase.user_laten.weight.data.cpu().numpy()
    fi = meta.MFbase.item_laten.weight.data.cpu().numpy()
    scale = np.abs(g["final_user"]).max()
    assert np.abs(fu - g["final_user"]).max() < 1e-3 * scale, np.abs(fu - g["final_user"]).max()
    assert np.abs(fi - g["final_item"]).max() < 1e-3 * np.abs(g["final_item"]).max()
    assert np.abs(meta.user_weight_hat.cpu().numpy() - g["final_user_hat"]).max() < 1e-3 * np.abs(g["final_user_hat"]).max()
    assert np.abs(meta.last_user_weight.cpu().numpy() - g["final_last_user"]).max() < 1e-3 * scale
    for net in ("user", "item"):
        mod = getattr(meta.transfer, net + "_transfer")
        for k in ("conv1.weight", "conv2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"):
            a, b = k.split(".")
            got = getattr(getattr(mod, a), b).detach().cpu().numpy()
            ref = g["tF.%s.%s" % (net, k)]
            assert np.abs(sample(got).reshape(ref.shape) - ref).max() < 2e-4, (net, k)
    assert meta.test_num == [int(x) for x in g["test_num"]]
    for key in ("recall", "recall_10", "recall_5"):
        got = np.array([float(x) for x in getattr(meta, key)])
        # 96 test rows per period: allow one borderline row (score gaps of a few ulp) to flip
        assert np.abs(got - g[key]).max() <= 1.0 / 96 + 1e-9, (key, got, g[key])
    for key in ("ndcg", "ndcg_10", "ndcg_5"):
        got = np.array([float(x) for x in getattr(meta, key)])
        assert np.abs(got - g[key]).max() < 1.2e-2, (key, got, g[key])
    # Adam step counters: one per optimizer step, surviving across periods (model/transfer.py:764)
    n_mf = sum(len(g["log%d" % n]) for n, k in enumerate(g["log_kinds"]) if str(k) == "MF") // 96 * 3
    assert meta.MF_optimizer.step_count == n_mf
