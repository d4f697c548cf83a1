"""Throughput of the fused plain-MF step (sml_plain_mf_step, north_star kernel 1) on tables that do not fit L2.

    python tools/plain_mf_bench.py [--rows 20000000] [--batch 65536] [--steps 20] [--mode sparse|exact] [--loss bpr|bce]

Prints one JSON line: triples/s, algorithmic GB/s (4 632 B per triple) and the fraction of the measured
HBM copy bandwidth (MEASURED_PEAKS.json).  Ids are uniform, so practically every row of a batch is distinct."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sml_b200 import ops  # noqa: E402

BYTES_PER_TRIPLE = 6 * 3 * 256 + 24


def run(rows, batch, steps, mode, loss, warmup=5, dev=None, seed=0):
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev); g.manual_seed(seed)
    U = I = rows
    pu = torch.randn(U, 64, device=dev, generator=g) * 0.1
    pi = torch.randn(I, 64, device=dev, generator=g) * 0.1
    z = torch.zeros_like
    mu, vu, mi, vi = z(pu), z(pu), z(pi), z(pi)
    st = ops.new_adam_state(dev, history=True)
    su, si = ops.new_row_stamps(U, st), ops.new_row_stamps(I, st)
    hu, hi = ops.new_list_heads(U, dev), ops.new_list_heads(I, dev)
    lossb = torch.zeros(2, device=dev)
    opt = ops.OPT_ADAM_SPARSE if mode == "sparse" else ops.OPT_ADAM_DENSE_EXACT
    lk = ops.LOSS_BPR if loss == "bpr" else ops.LOSS_BCE
    ids = [(torch.randint(0, U, (batch,), device=dev, generator=g), torch.randint(0, I, (batch,), device=dev, generator=g),
            torch.randint(0, I, (batch,), device=dev, generator=g)) for _ in range(warmup + steps)]

    def step(k):
        u, i, j = ids[k]
        ops.plain_mf_step(pu, pi, mu, vu, mi, vi, hu, hi, u, i, j, st, 0.01, lossb, loss=lk, l2_u=1e-4, l2_i=1e-4, optimizer=opt,
                          stamp_user=su, stamp_item=si)
    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(warmup, warmup + steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gbs = batch * BYTES_PER_TRIPLE / ms / 1e6
    return dict(kernel="k_plain_mf_step", rows_per_table=rows, batch=batch, steps=steps, mode=mode, loss=loss, ms_per_step=ms,
                triples_per_s=batch / ms * 1e3, algorithmic_bytes_per_triple=BYTES_PER_TRIPLE, achieved_gbs=gbs, peak_gbs=peak,
                frac=gbs / peak, peak_source="measured" if "hbm_gbs" in peaks else "fallback", loss_value=float(lossb[0]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mode", default="sparse", choices=["sparse", "exact"])
    ap.add_argument("--loss", default="bpr", choices=["bpr", "bce"])
    a = ap.parse_args()
    print(json.dumps(run(a.rows, a.batch, a.steps, a.mode, a.loss)))


if __name__ == "__main__":
    main()
