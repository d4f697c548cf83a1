import sys, os, numpy as np, torch, tempfile, io, contextlib
sys.path.insert(0, "/root/repo")
from tests.test_host_logic import make_args, write_fixture_stream
from tests.helpers import theta_from_chk
from sml_b200.data.dataset2 import transfer_data
from sml_b200.model.transfer import meta_train
name = sys.argv[1]
g = np.load("/root/repo/tests/golden/%s.npz" % name)
tmp = tempfile.mkdtemp()
with contextlib.redirect_stdout(io.StringIO()):
    NP, U, I = write_fixture_stream(g, tmp)
    args = make_args(g, tmp, False, news=name.endswith("news"))
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
    ds = transfer_data(args, path=args.data_path, datasetname="mini", file_path_list=[str(i) for i in range(NP)], test_list=[str(j) for j in range(5, NP)], validation_list=None, online_train_time=2, online_test_time=5)
    meta = meta_train(args, ds, U, I, 64)
    tu, ti = theta_from_chk(g["theta_com"])
    sd = {("user_transfer." + k): torch.from_numpy(v) for k, v in tu.items()}; sd.update({("item_transfer." + k): torch.from_numpy(v) for k, v in ti.items()})
    meta.transfer.load_state_dict(sd)
ref = g["stage_sums"]
for st in range(len(ref)):
    with contextlib.redirect_stdout(io.StringIO()):
        meta.train_one_stage3(args, st)
    uw = meta.MFbase.user_laten.weight.data.double(); iw = meta.MFbase.item_laten.weight.data.double()
    th = sum(p.double().abs().sum() for p in meta.transfer.parameters())
    got = [float(uw.sum()), float(uw.abs().sum()), float(iw.sum()), float(iw.abs().sum()), float(th)]
    print(name, os.environ.get("SML_GEMM", "tc"), "stage", st, " ".join("%.3e" % (abs(a - b) / max(abs(b), 1e-9)) for a, b in zip(got, ref[st])))
