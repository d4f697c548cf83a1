"""Times one SML MF step, one transfer step and the full-table transfer on Yelp-shaped tables with CUDA
events; the steps are captured in a CUDA graph so host launch overhead does not hide the device time.
SML_GEMM=simt selects the SIMT GEMM path for an A/B comparison."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200 import ops  # noqa: E402
from sml_b200.model.conv_transfer import ConvTransfer_com  # noqa: E402


PROFILE = os.environ.get("SML_PROFILE", "")     # under ncu: plain launches, no graph


def timed_graph(fn, reps):
    fn(); torch.cuda.synchronize()
    if PROFILE:
        fn(); torch.cuda.synchronize()
        return float("nan")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    U, I = 59082, 122816
    g = torch.Generator().manual_seed(0)
    T = lambda *s: torch.randn(*s, generator=g).to(dev)
    with torch.random.fork_rng(devices=[]):
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            tr = ConvTransfer_com(64, 64).to(dev)
    lu, li, hu, hi = T(U, 64), T(I, 64), T(U, 64), T(I, 64)
    z = lambda t: torch.zeros_like(t)
    for B, kind in ((1024, "mf"), (1024, "mfl"), (256, "tr"), (248, "mf"), (8192, "mf"), (8192, "mfl"), (8192, "tr")):
        if PROFILE and "%s%d" % (kind, B) not in PROFILE.split(","):
            continue
        u = torch.randint(0, U, (B,), generator=g).to(dev); i = torch.randint(0, I, (B,), generator=g).to(dev)
        j = torch.randint(0, I, (B,), generator=g).to(dev)
        loss = torch.zeros(2, device=dev)
        ws = torch.zeros(int(ops.lib().sml_step_workspace_bytes(B)), dtype=torch.uint8, device=dev)
        if kind in ("mf", "mfl"):      # mf: dense Adam sweep; mfl: row-lazy exact Adam (meta_train's default), same ids every step
            st = ops.new_adam_state(dev, history=kind == "mfl")
            stamps = dict(stamp_user=ops.new_row_stamps(U, st), stamp_item=ops.new_row_stamps(I, st)) if kind == "mfl" else {}
            a = ops.make_step_args(user=u, item=i, neg=j, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta,
                                   adam_state=st, lr=1e-4, l2=1e-6, loss_out=loss, workspace=ws,
                                   g_user=z(hu), g_item=z(hi), m_user=z(hu), v_user=z(hu), m_item=z(hi), v_item=z(hi), **stamps)
            us = timed_graph(lambda: ops.mf_step(a, flush=False), 10)
        else:
            a = ops.make_step_args(user=u, item=i, neg=j, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta,
                                   adam_state=ops.new_adam_state(dev), lr=1e-5, l2=1e-4, loss_out=loss, workspace=ws,
                                   g_theta=tr.theta_grad, m_theta=z(tr.theta), v_theta=z(tr.theta))
            us = timed_graph(lambda: ops.tr_step(a), 10)
        print("%-3s step  B=%5d : %8.1f us   (%.2f M triples/s)" % (kind, B, us, B / us))
    out = torch.empty_like(hu)
    us = timed_graph(lambda: ops.transfer_forward(lu, hu, tr.theta[:ops.NET_STRIDE], out=out), 3)
    print("transfer_forward %d rows: %8.1f us  (%.1f TFLOP/s fp32-equiv, %.0f M rows/s)" % (U, us, U * 403456 / us / 1e6, U / us))
    out = torch.empty_like(hi)
    us = timed_graph(lambda: ops.transfer_forward(li, hi, tr.theta[ops.NET_STRIDE:], out=out), 3)
    print("transfer_forward %d rows: %8.1f us  (%.1f TFLOP/s fp32-equiv, %.0f M rows/s)" % (I, us, I * 403456 / us / 1e6, I / us))


if __name__ == "__main__":
    main()
