"""Critical path of one transfer step (B = 256) by difference: sml_debug_set_mask drops stages of sml_tr_step, each variant is
captured in a CUDA graph (10 steps) and timed with CUDA events.  Profiling aid; the numbers go to profiles/."""
import contextlib, io, os, sys
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
def timed(fn):
    fn(); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10):
            fn()
    graph.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10 * 1e3)
    return best


print("transfer step, B = %d" % B)
for mask, name in [(0, "full step"), (512, "full step, wide (128 x 128) fc1 / d2 tiles"), (1, "no weight gradients"), (2, "no dA / conv backward"), (3, "neither (through dZ1, then Adam)"),
                   (64, "no Adam"), (4, "through the loss"), (8, "through fc2"), (128, "through fc1"),
                   (256, "theta packing || conv prologue only")]:
    ops.lib().sml_debug_set_mask(mask)
    print("mask %3d  %-45s %7.1f us" % (mask, name, timed(lambda: ops.tr_step(a))))

Bm = int(os.environ.get("SML_BM", 1024))
u, i, j = (torch.randint(0, n, (Bm,), generator=g).to(dev) for n in (U, I, I))
wsm = torch.zeros(int(ops.lib().sml_step_workspace_bytes(Bm)), dtype=torch.uint8, device=dev)
for lazy in (False, True):
    st = ops.new_adam_state(dev, history=lazy)
    stamps = dict(stamp_user=ops.new_row_stamps(U, st), stamp_item=ops.new_row_stamps(I, st)) if lazy else {}
    am = ops.make_step_args(user=u, item=i, neg=j, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta, adam_state=st,
                            lr=1e-6, l2=1e-6, loss_out=loss, workspace=wsm, g_user=z(hu), g_item=z(hi), m_user=z(hu), v_user=z(hu),
                            m_item=z(hi), v_item=z(hi), **stamps)
    print("MF step, B = %d, %s Adam (theta packed every step here; once per epoch in sml_mf_epoch)" % (Bm, "row-lazy" if lazy else "dense"))
    for mask, name in [(0, "full step"), (64, "no Adam update"), (2, "no dA / conv backward (then Adam)"), (4, "through the loss"),
                       (8, "through fc2"), (128, "through fc1"), (256, "tick, (catch-up,) theta packing, conv prologue")]:
        ops.lib().sml_debug_set_mask(mask)
        print("mask %3d  %-45s %7.1f us" % (mask, name, timed(lambda: ops.mf_step(am, flush=False))))
ops.lib().sml_debug_set_mask(0)
