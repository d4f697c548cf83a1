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
        """Sequential rejection sampling in sampler order: the draws come from numpy's global generator in exactly the
        amounts the reference consumes (as many as samples remain, again for the rejected ones, ...), the walk over them
        is a C helper with a hashed membership test."""
        from .._lib import lib
        import ctypes as C
        n = len(order)
        users = np.ascontiguousarray(ds.user[order], dtype=np.int64)
        item_all = np.ascontiguousarray(ds.item_all, dtype=np.int64)
        table = ds.key_table()
        P = len(item_all)
        neg = np.empty(n, dtype=np.int64)
        walk = lib().sml_host_rejection_walk_hashed
        s = C.c_int64(0)
        while s.value < n:
            k = n - s.value
            draws = np.ascontiguousarray(np.random.randint(0, P, size=k), dtype=np.int64)
            used = walk(draws.ctypes.data, k, users.ctypes.data, n, C.byref(s), item_all.ctypes.data, table.ctypes.data, len(table),
                        ds._span, neg.ctypes.data)
            assert used == k                   # every remaining sample needs at least one draw
        return neg.astype(ds.item_all.dtype, copy=False)


_UI_CACHE = []      # [(file array, contiguous user column, contiguous item column)], last two period files


def _user_item_columns(d):
    """Contiguous copies of columns 0 and 1 of a period file, made once per file: the per-epoch gathers then hit two
    600 KB arrays instead of striding through the 600 MB file."""
    for arr, u, i in _UI_CACHE:
        if arr is d:
            return u, i
    u, i = np.ascontiguousarray(d[:, 0], dtype=np.int64), np.ascontiguousarray(d[:, 1], dtype=np.int64)
    _UI_CACHE.append((d, u, i))
    del _UI_CACHE[:-2]
    return u, i


def mf_epoch_triples(ds: trainDataset_withPreSample, order):
    """One MF epoch over a pre-sampled dataset in the given sample order (data/dataset2.py:191-201)."""
    col = ds.current_column()
    d = ds.all_data
    u_all, i_all = _user_item_columns(d)
    u, i, j = u_all[order], i_all[order], d[order, col]
    ds.advance_epoch()
    return u, i, np.ascontiguousarray(j, dtype=np.int64)


def tr_epoch_triples(ds: offlineDataset_withsample, order):
    neg = ReferenceStream.alone_negatives(ds, order)
    return (np.ascontiguousarray(ds.user[order], dtype=np.int64), np.ascontiguousarray(ds.item[order], dtype=np.int64),
            np.ascontiguousarray(neg, dtype=np.int64))
