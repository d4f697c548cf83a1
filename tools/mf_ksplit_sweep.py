"""Split-K factors of fc2 and d1 in the MF step (B = 1024, 3 072 rows): sml_debug_set_ksplit, CUDA-graph replay of 10 steps."""
import contextlib, io, itertools, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200 import ops  # noqa: E402
from sml_b200.model.conv_transfer import ConvTransfer_com  # noqa: E402

dev = torch.device("cuda:0")
U, I, B = 59082, 122816, int(os.environ.get("SML_BM", 1024))
g = torch.Generator().manual_seed(0)
R = lambda *s: torch.randn(*s, generator=g).to(dev)
with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
    tr = ConvTransfer_com(64, 64).to(dev)
lu, li, hu, hi = R(U, 64), R(I, 64), R(U, 64), R(I, 64)
u, i, j = (torch.randint(0, n, (B,), generator=g).to(dev) for n in (U, I, I))
loss = torch.zeros(2, device=dev)
ws = torch.zeros(int(ops.lib().sml_step_workspace_bytes(B)), dtype=torch.uint8, device=dev)
z = torch.zeros_like
st = ops.new_adam_state(dev, history=True)
a = ops.make_step_args(user=u, item=i, neg=j, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta, adam_state=st,
                       lr=1e-6, l2=1e-6, loss_out=loss, workspace=ws, g_user=z(hu), g_item=z(hi), m_user=z(hu), v_user=z(hu),
                       m_item=z(hi), v_item=z(hi), stamp_user=ops.new_row_stamps(U, st), stamp_item=ops.new_row_stamps(I, st))


def timed():
    ops.mf_step(a, flush=False); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10):
            ops.mf_step(a, flush=False)
    graph.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10 * 1e3)
    return best


base = timed()
print("built-in: %.1f us" % base)
for fc2, d1 in itertools.product((1, 2, 4, 8), (1, 2, 4)):
    ops.lib().sml_debug_set_ksplit(fc2, d1, 0, 0)
    print("fc2 %d  d1 %d : %.1f us" % (fc2, d1, timed()))
ops.lib().sml_debug_set_ksplit(0, 0, 0, 0)
