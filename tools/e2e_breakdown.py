"""Where the wall time of an e2e period goes (bench.py's e2e arm: host period files, every loss / metric read back).

    python tools/e2e_breakdown.py [--parity] [--periods 3]

Wraps the methods of one meta_train instance with host wall-clock accumulators (no extra synchronisation: every
evaluation and every epoch already ends with a blocking read) and prints ms per period and calls per period.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--periods", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    from sml_b200.data.memory_stream import MemoryStream
    from sml_b200.model import MF
    from sml_b200.model.transfer import meta_train
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    shape = dict(bench.YELP)
    U, I = shape["n_users"], shape["n_items"]
    W, K = a.warmup, a.periods
    periods = bench.synth_periods(W + K + 1, shape, seed=100)
    args = bench.make_args()
    torch.manual_seed(args.seed); np.random.seed(args.seed + 2)
    pre = MF.MFbasemode(U, I, 64)
    args.pre_model = "/tmp/sml_e2e_pre.pt"
    torch.save(pre.state_dict(), args.pre_model)
    pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
    ds = MemoryStream([(pin(tr), pin(te)) for tr, te in periods], U, I)
    out = sys.stdout
    sys.stdout = open(os.devnull, "w")
    meta = meta_train(args, ds, U, I, 64, device=dev, device_sampler=not a.parity, emulate_reference_rng=a.parity)
    acc = collections.defaultdict(lambda: [0, 0.0])
    depth = [0]

    def wrap(obj, name, label=None):
        f = getattr(obj, name)
        label = label or name

        def g(*x, **k):
            t = time.perf_counter()
            try:
                return f(*x, **k)
            finally:
                acc[label][0] += 1; acc[label][1] += time.perf_counter() - t
        setattr(obj, name, g)
    for n in ("_triples", "_mf_epoch", "_tr_epoch", "_eval", "flush_deferred", "updata", "save_MF_weight", "_test_set", "_sample_dataset", "prefetch_files",
              "get_next_data", "_upload", "MF_TrainDataset", "_run_epoch"):
        wrap(meta, n)
    stage = 0
    for _ in range(W):
        meta.train_one_stage3(args, stage); stage += 1
    torch.cuda.synchronize()
    acc.clear()
    t0 = time.perf_counter()
    for _ in range(K):
        meta.train_one_stage3(args, stage); stage += 1
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sys.stdout = out
    res = {"ms_per_period": wall / K * 1e3, "mode": "parity" if a.parity else "device_sampler",
           "sections_ms_per_period": {k: dict(calls=v[0] / K, ms=v[1] / K * 1e3) for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
