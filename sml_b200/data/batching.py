"""Host-side batch construction for the two hot loops.

The reference feeds its loops from ``torch.utils.data.DataLoader(shuffle=True)`` over per-sample
Python ``__getitem__`` calls (model/transfer.py:438-443,692-696).  Here a whole epoch is
materialised at once as three int64 arrays (user, item, neg) in batch order, uploaded once, and
the step kernels walk it on the device.

``ReferenceStream`` reproduces, draw for draw, what the reference consumes from the *global* torch
and numpy generators with ``--numworkers 0``, so that a run seeded like main_yelp.py:137-140 sees
bit-identical triples ("sampled indices bit-exact", BASELINE.json north_star):
  * every DataLoader iterator draws one int64 ``_base_seed`` from the global torch generator,
    also the un-shuffled evaluation loaders (torch/utils/data/dataloader.py, _BaseDataLoaderIter);
  * a shuffled loader then draws the RandomSampler seed and permutes with a private generator;
  * trainDataset_withPreSample shuffles its column list with ``np.random.shuffle``;
  * offlineDataset_withsample calls ``np.random.choice(item_all, 1)`` once per sample (and once per
    rejection) in sampler order.
"""
from __future__ import annotations

import numpy as np
import torch

from .dataset import offlineDataset_withsample
from .dataset2 import trainDataset_withPreSample


class ReferenceStream(object):
    """Mirrors the reference's consumption of the global RNGs (see module docstring)."""

    @staticmethod
    def loader_iter():
        """One DataLoader iterator creation (train or eval)."""
        torch.empty((), dtype=torch.int64).random_()

    @staticmethod
    def shuffled_order(n):
        ReferenceStream.loader_iter()
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g).numpy()

    @staticmethod
    def alone_negatives(ds: offlineDataset_withsample, order):
        """Sequential rejection sampling in sampler order, vectorised between rejections."""
        from .._lib import lib
        n = len(order)
        users = np.ascontiguousarray(ds.user[order], dtype=np.int64)
        item_all = np.ascontiguousarray(ds.item_all, dtype=np.int64)
        keys = np.ascontiguousarray(ds._keys, dtype=np.int64)
        P = len(item_all)
        state = np.random.get_state()
        margin = max(64, n // 4)
        neg = np.empty(n, dtype=np.int64)
        walk = lib().sml_host_rejection_walk
        while True:
            draws = np.ascontiguousarray(np.random.randint(0, P, size=n + margin), dtype=np.int64)
            p = walk(draws.ctypes.data, len(draws), users.ctypes.data, n, item_all.ctypes.data, keys.ctypes.data, len(keys),
                     ds._span, neg.ctypes.data)
            if p >= 0:
                break
            np.random.set_state(state)
            margin *= 4
        # leave the global generator exactly where the reference would: p draws consumed
        np.random.set_state(state)
        if p:
            np.random.randint(0, P, size=p)
        return neg.astype(ds.item_all.dtype, copy=False)


def mf_epoch_triples(ds: trainDataset_withPreSample, order):
    """One MF epoch over a pre-sampled dataset in the given sample order (data/dataset2.py:191-201)."""
    col = ds.current_column()
    d = ds.all_data
    u, i, j = d[order, 0], d[order, 1], d[order, col]
    ds.advance_epoch()
    return np.ascontiguousarray(u, dtype=np.int64), np.ascontiguousarray(i, dtype=np.int64), np.ascontiguousarray(j, dtype=np.int64)


def tr_epoch_triples(ds: offlineDataset_withsample, order):
    neg = ReferenceStream.alone_negatives(ds, order)
    return (np.ascontiguousarray(ds.user[order], dtype=np.int64), np.ascontiguousarray(ds.item[order], dtype=np.int64),
            np.ascontiguousarray(neg, dtype=np.int64))
