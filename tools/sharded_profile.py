"""Kernel-level breakdown of one sharded MF epoch and one sharded transfer epoch at world = 1 on the bench's per-GPU shape
(torch.profiler, CUDA activities): where the time of a sharded step goes apart from the collectives.

    python tools/sharded_profile.py [--users 25000000] [--items 2500000] [--batch 8192] [--steps 16]
"""
import argparse
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=25_000_000)
    ap.add_argument("--items", type=int, default=2_500_000)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=16)
    a = ap.parse_args()
    from sml_b200.model.conv_transfer import ConvTransfer_com
    from sml_b200.shard import ShardedSML
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    with torch.random.fork_rng(devices=[]), contextlib.redirect_stdout(io.StringIO()):
        tr = ConvTransfer_com(64, 64).to(dev)
    ut = torch.empty(a.users, 64, device=dev).normal_(0, 0.1, generator=g)
    it = torch.empty(a.items, 64, device=dev).normal_(0, 0.1, generator=g)
    sh = ShardedSML(ut, it, tr)
    rnd = lambda n, hi: torch.randint(0, hi, (n,), device=dev, generator=g)
    n = a.batch * a.steps
    tri = lambda: (rnd(n, a.users), rnd(n, a.items), rnd(n, a.items))
    sh.save_last()
    sh.mf_epoch(*tri(), a.batch); sh.flush(); sh.save_hat(); sh.tr_epoch(*tri(), a.batch)
    torch.cuda.synchronize()
    for name, fn in (("mf_epoch", lambda t: (sh.mf_epoch(*t, a.batch), sh.flush())), ("tr_epoch", lambda t: sh.tr_epoch(*t, a.batch))):
        t = tri()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(t); e1.record(); torch.cuda.synchronize()
        print("%s: %.3f ms per step (CUDA events, %d steps)" % (name, e0.elapsed_time(e1) / a.steps, a.steps))
        t = tri()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
            fn(t)
            torch.cuda.synchronize()
        rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0]
        rows.sort(key=lambda r: -r[1])
        tot = sum(r[1] for r in rows)
        print("  device time by kernel (us per step), total %.1f us per step:" % (tot / a.steps))
        for k, us, c in rows[:22]:
            print("    %-90s %8.1f  x%-4d %5.1f%%" % (k[:90], us / a.steps, c, 100.0 * us / tot))


if __name__ == "__main__":
    main()
