"""Sweep of the split-K factors of the transfer step's small GEMMs (fc2, d1 on the main stream; dW2, dW1 on the side stream):
more slices spread one kernel over more SMs but leave no room for the concurrent branch.  Tuning aid (sml_debug_set_ksplit)."""
import contextlib, io, itertools, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200 import ops  # noqa: E402
from sml_b200.model.conv_transfer import ConvTransfer_com  # noqa: E402

dev = torch.device("cuda:0")
U, I, B = 59082, 122816, int(os.environ.get("SML_B", 256))
g = torch.Generator().manual_seed(0)
R = lambda *s: torch.randn(*s, generator=g).to(dev)
with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
    tr = ConvTransfer_com(64, 64).to(dev)
lu, li, hu, hi = R(U, 64), R(I, 64), R(U, 64), R(I, 64)
u, i, j = (torch.randint(0, n, (B,), generator=g).to(dev) for n in (U, I, I))
loss = torch.zeros(2, device=dev)
ws = torch.zeros(int(ops.lib().sml_step_workspace_bytes(B)), dtype=torch.uint8, device=dev)
z = torch.zeros_like
a = ops.make_step_args(user=u, item=i, neg=j, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta,
                       adam_state=ops.new_adam_state(dev), lr=1e-6, l2=1e-4, loss_out=loss, workspace=ws,
                       g_theta=tr.theta_grad, m_theta=z(tr.theta), v_theta=z(tr.theta))


def timed():
    ops.tr_step(a); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10):
            ops.tr_step(a)
    graph.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10 * 1e3)
    return best


res = []
for fc2, d1, w2, w1 in itertools.product((0,), (1, 2, 4, 8), (1, 2, 4, 8, 16), (1, 2, 4, 8)):      # (fc2 is fused into fc1 at this batch size)
    ops.lib().sml_debug_set_ksplit(fc2, d1, w2, w1)
    res.append((timed(), fc2, d1, w2, w1))
ops.lib().sml_debug_set_ksplit(0, 0, 0, 0)
base = timed()
res.sort()
print("built-in choice (d1 4, dW2 8, dW1 4): %.1f us" % base)
for t, fc2, d1, w2, w1 in res[:16] + res[-3:]:
    print("fc2 %d  d1 %d  dW2 %d  dW1 %d : %.1f us" % (fc2, d1, w2, w1, t))
