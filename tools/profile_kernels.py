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
    Bt = 256
    a_tr = ops.make_step_args(user=u[:Bt].contiguous(), item=i[:Bt].contiguous(), neg=j[:Bt].contiguous(), last_user=lu, last_item=li, hat_user=hu.clone(),
                              hat_item=hi.clone(), theta=tr.theta, adam_state=ops.new_adam_state(dev), lr=1e-5, l2=1e-4, loss_out=loss,
                              g_theta=tr.theta_grad, m_theta=z(tr.theta), v_theta=z(tr.theta))
    eu = torch.randint(0, U, (16384,), generator=g).to(dev); ep = torch.randint(0, I, (16384,), generator=g).to(dev)
    ipk = ops.pack_rows(hi)
    work = [("eval_candidates 75000x1001", lambda: ops.eval_candidates(hu, hi, rows), N * 264264 / 1e9, "GB"),
            ("transfer_forward 59082 rows", lambda: ops.transfer_forward(lu, hu, tr.theta[:ops.NET_STRIDE], out=out), U * 403456 / 1e12, "TFLOP"),
            ("mf_step B=1024", lambda: ops.mf_step(a), 1024, "triples"),
            ("tr_step B=256", lambda: ops.tr_step(a_tr), 256, "triples"),
            ("fullcat_rank 16384 users x 122816 items", lambda: ops.fullcat_ranks(hu, hi, eu, ep, items_packed=ipk, n_items=I), 16384 * I * 128 / 1e12, "TFLOP")]
    # round 2 additions: fused transfer forward, fused plain-MF step, row-lazy Adam, owner-side exchange kernels, top-k
    big = 1_000_000
    a1, b1, o1 = R(big, 64), R(big, 64), torch.empty(big, 64, device=dev)
    work.append(("transfer_fused 1M rows", lambda: ops.transfer_forward(a1, b1, tr.theta[:ops.NET_STRIDE], out=o1), big * 403456 / 1e12, "TFLOP"))
    st = ops.new_adam_state(dev, history=True)
    pu, pi = R(2_000_000, 64), R(2_000_000, 64)
    mu, vu, mi, vi = z(pu), z(pu), z(pi), z(pi)
    su, si = ops.new_row_stamps(2_000_000, st), ops.new_row_stamps(2_000_000, st)
    hd_u, hd_i = ops.new_list_heads(2_000_000, dev), ops.new_list_heads(2_000_000, dev)
    Bp = 65536
    pu_ids = [torch.randint(0, 2_000_000, (Bp,), generator=g).to(dev) for _ in range(3)]
    work.append(("plain_mf_step 65536 triples (exact dense Adam, row-lazy)",
                 lambda: ops.plain_mf_step(pu, pi, mu, vu, mi, vi, hd_u, hd_i, pu_ids[0], pu_ids[1], pu_ids[2], st, 0.01, loss, loss=ops.LOSS_BPR,
                                           optimizer=ops.OPT_ADAM_DENSE_EXACT, stamp_user=su, stamp_item=si), Bp * 4632 / 1e9, "GB"))
    work.append(("adam_flush 2M rows", lambda: ops.adam_flush(pu, mu, vu, su, st), 2_000_000 * 6 * 256 / 1e9, "GB"))
    loc = torch.randint(0, U, (24576,), generator=g).to(dev)
    drow = R(24576, 64)
    gtab = z(hu)
    work.append(("gather_pairs 24576 rows", lambda: ops.gather_pairs(lu, hu, loc), 24576 * 1024 / 1e9, "GB"))
    work.append(("scatter_grads 24576 rows", lambda: ops.scatter_grads(gtab, hu, loc, drow, 1.0, 1e-6), 24576 * 768 / 1e9, "GB"))
    work.append(("fullcat_topk 16384 users x 122816 items, K=20", lambda: ops.fullcat_topk(hu, hi, eu, 20, items_packed=ipk, n_items=I), 16384 * I * 128 / 1e12, "TFLOP"))
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
