"""Timeline of the kernels of one transfer step inside a CUDA-graph replay (torch.profiler / CUPTI timestamps): start offset and
duration of every kernel of a middle step, so that gaps (dependent-launch latency) and overlaps (side branch) are visible.

    python tools/tr_timeline.py [--kind tr|mf] [--steps 12]
"""
import argparse
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200 import ops  # noqa: E402
from sml_b200.model.conv_transfer import ConvTransfer_com  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="tr")
    ap.add_argument("--steps", type=int, default=12)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    U, I = 59082, 122816
    g = torch.Generator().manual_seed(0)
    R = lambda *s: torch.randn(*s, generator=g).to(dev)
    with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
        tr = ConvTransfer_com(64, 64).to(dev)
    lu, li, hu, hi = R(U, 64), R(I, 64), R(U, 64), R(I, 64)
    z = torch.zeros_like
    B = 256 if a.kind == "tr" else 1024
    n = B * a.steps
    u = torch.randint(0, U, (n,), generator=g).to(dev); i = torch.randint(0, I, (n,), generator=g).to(dev); j = torch.randint(0, I, (n,), generator=g).to(dev)
    loss = torch.zeros(2, device=dev)
    ws = torch.zeros(int(ops.lib().sml_step_workspace_bytes(B)), dtype=torch.uint8, device=dev)
    if a.kind == "tr":
        args = ops.make_step_args(user=u, item=i, neg=j, batch=B, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta,
                                  adam_state=ops.new_adam_state(dev), lr=1e-5, l2=1e-4, loss_out=loss, workspace=ws,
                                  g_theta=tr.theta_grad, m_theta=z(tr.theta), v_theta=z(tr.theta))
        fn = lambda: ops.tr_epoch(args, n)
    else:
        st = ops.new_adam_state(dev, history=True)
        args = ops.make_step_args(user=u, item=i, neg=j, batch=B, last_user=lu, last_item=li, hat_user=hu, hat_item=hi, theta=tr.theta,
                                  adam_state=st, lr=1e-4, l2=1e-6, loss_out=loss, workspace=ws, g_user=z(hu), g_item=z(hi), m_user=z(hu),
                                  v_user=z(hu), m_item=z(hi), v_item=z(hi), stamp_user=ops.new_row_stamps(U, st), stamp_item=ops.new_row_stamps(I, st))
        fn = lambda: ops.mf_epoch(args, n)
    fn(); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay(); torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        graph.replay()
        torch.cuda.synchronize()
    ev = sorted([(e.time_range.start, e.time_range.end - e.time_range.start, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
                key=lambda x: x[0])
    ev = [e for e in ev if "k_" in e[2]]
    per = len(ev) // a.steps
    print("%d kernels, %d per step; total %.1f us per step" % (len(ev), per, (ev[-1][0] + ev[-1][1] - ev[0][0]) / a.steps))
    s0 = (a.steps // 2) * per
    # a step starts at its theta packing / first kernel: align on the first k_pack_theta / k_adam_tick at or after s0
    first = next(k for k in range(s0, len(ev)) if ("k_pack_theta" in ev[k][2] or "k_adam_tick" in ev[k][2]))
    base = ev[first][0]
    end_prev = 0.0
    for k in range(first, min(first + per, len(ev))):
        t, d, name = ev[k]
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = name[:name.index("(")] if "(" in name else name
        print("  +%7.1f us  %6.1f us  (gap after the previous end %+6.1f)  %s" % (t - base, d, (t - base) - end_prev, name[:70]))
        end_prev = max(end_prev, t - base + d)
    nxt = ev[min(first + per, len(ev) - 1)][0] - base
    print("  next step starts at +%.1f us" % nxt)


if __name__ == "__main__":
    main()
