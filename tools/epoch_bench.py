"""Host launch time vs device time of one transfer epoch and one MF epoch enqueued by the C-side epoch loops."""
import contextlib, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200 import ops  # noqa: E402
from sml_b200.model.conv_transfer import ConvTransfer_com  # noqa: E402

dev = torch.device("cuda:0")
U, I = 59082, 122816
g = torch.Generator().manual_seed(0)
R = lambda *s: torch.randn(*s, generator=g).to(dev)
with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
    tr = ConvTransfer_com(64, 64).to(dev)
lu, li, hu, hi = R(U, 64), R(I, 64), R(U, 64), R(I, 64)
z = torch.zeros_like
for kind, B, steps in (("tr", 256, 293), ("mf", 1024, 74), ("mf-lazy", 1024, 74)):
    n = B * steps
    u = torch.randint(0, U, (n,), generator=g).to(dev); i = torch.randint(0, I, (n,), generator=g).to(dev); j = torch.randint(0, I, (n,), generator=g).to(dev)
    loss = torch.zeros(2, device=dev)
    ws = torch.zeros(int(ops.lib().sml_step_workspace_bytes(B)), dtype=torch.uint8, device=dev)
    if kind == "tr":
        a = ops.make_step_args(user=u, item=i, neg=j, batch=B, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta,
                               adam_state=ops.new_adam_state(dev), lr=1e-5, l2=1e-4, loss_out=loss, workspace=ws,
                               g_theta=tr.theta_grad, m_theta=z(tr.theta), v_theta=z(tr.theta))
        fn = lambda: ops.tr_epoch(a, n)
    else:
        st = ops.new_adam_state(dev, history=kind == "mf-lazy")
        stamps = dict(stamp_user=ops.new_row_stamps(U, st), stamp_item=ops.new_row_stamps(I, st)) if kind == "mf-lazy" else {}
        a = ops.make_step_args(user=u, item=i, neg=j, batch=B, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta,
                               adam_state=st, lr=1e-4, l2=1e-6, loss_out=loss, workspace=ws,
                               g_user=z(hu), g_item=z(hi), m_user=z(hu), v_user=z(hu), m_item=z(hi), v_item=z(hi), **stamps)
        fn = lambda: ops.mf_epoch(a, n)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print("%s epoch %d steps: host enqueue %.1f ms (%.1f us/step), device %.1f ms (%.1f us/step), wall %.1f ms"
          % (kind, steps, t_host * 1e3, t_host / steps * 1e6, e0.elapsed_time(e1), e0.elapsed_time(e1) / steps * 1e3, t_all * 1e3))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    print("   graph replay: device %.2f ms per epoch (%.1f us/step)" % (e0.elapsed_time(e1) / 3, e0.elapsed_time(e1) / 3 / steps * 1e3))
