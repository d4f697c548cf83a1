"""One launch of each hot kernel on the bench shapes (Yelp-shaped tables), for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:'k_eval_candidates|k_umma_packed|k_conv_fwd|k_fullcat_rank|k_adam_dense' \
        -o gpurun_out/r01_kernels python tools/profile_kernels.py
Also prints CUDA-event timings (warm, not under ncu) when SML_TIME=1."""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200 import ops  # noqa: E402
from sml_b200.model.conv_transfer import ConvTransfer_com  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    U, I, N = 59082, 122816, 75000
    g = torch.Generator().manual_seed(0)
    R = lambda *s: torch.randn(*s, generator=g).to(dev)
    with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
        tr = ConvTransfer_com(64, 64).to(dev)
    lu, li, hu, hi = R(U, 64), R(I, 64), R(U, 64), R(I, 64)
    rows = torch.cat([torch.randint(0, U, (N, 1), generator=g), torch.randint(0, I, (N, 1000), generator=g)], 1).to(dev)
    out = torch.empty_like(hu)
    z = torch.zeros_like
    B = 1024
    u = torch.randint(0, U, (B,), generator=g).to(dev); i = torch.randint(0, I, (B,), generator=g).to(dev); j = torch.randint(0, I, (B,), generator=g).to(dev)
    loss = torch.zeros(2, device=dev)
    a = ops.make_step_args(user=u, item=i, neg=j, last_user=lu, last_item=li, hat_user=hu.clone(), hat_item=hi.clone(), theta=tr.theta,
                           adam_state=ops.new_adam_state(dev), lr=1e-4, l2=1e-6, loss_out=loss,
                           g_user=z(hu), g_item=z(hi), m_user=z(hu), v_user=z(hu), m_item=z(hi), v_item=z(hi))
    eu = torch.randint(0, U, (16384,), generator=g).to(dev); ep = torch.randint(0, I, (16384,), generator=g).to(dev)
    ipk = ops.pack_rows(hi)
    work = [("eval_candidates 75000x1001", lambda: ops.eval_candidates(hu, hi, rows), N * 264264 / 1e9, "GB"),
            ("transfer_forward 59082 rows", lambda: ops.transfer_forward(lu, hu, tr.theta[:ops.NET_STRIDE], out=out), U * 403456 / 1e12, "TFLOP"),
            ("mf_step B=1024", lambda: ops.mf_step(a), 1024, "triples"),
            ("fullcat_rank 16384 users x 122816 items", lambda: ops.fullcat_ranks(hu, hi, eu, ep, items_packed=ipk, n_items=I), 16384 * I * 128 / 1e12, "TFLOP")]
    for name, fn, units, unit in work:
        fn(); torch.cuda.synchronize()
        if os.environ.get("SML_TIME"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("%-45s %9.3f ms   %10.2f %s/s" % (name, ms, units / ms * 1e3, unit))


if __name__ == "__main__":
    main()
