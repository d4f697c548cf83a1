"""On-the-fly negative sampler -- drop-in for ``offlineDataset_withsample`` of the reference's
data/dataset.py:41-71 (the only class of that file on the SML path; it feeds the transfer step
with ``TR_sample_type="alone"``, model/transfer.py:661-662)."""
from __future__ import annotations

import numpy as np


class offlineDataset_withsample(object):
    def __init__(self, dataset):
        self.user = dataset[:, 0]
        self.item = dataset[:, 1]
        print("user max:", self.user.max())
        print("user max:", self.item.max())          # sic (data/dataset.py:49)
        n_item_span = int(self.item.max()) + 1
        if self.item.min() >= 0 and n_item_span <= 8 * max(len(self.item), 1 << 16):
            self.item_all = np.flatnonzero(np.bincount(self.item, minlength=n_item_span)).astype(self.item.dtype)   # == np.unique
        else:
            self.item_all = np.unique(self.item)
        self._span = n_item_span
        self._keys_sorted = None
        self._table = None

    @property
    def _keys(self):
        """user -> items of this period as a sorted key array (vectorised membership tests, the GPU sampler)."""
        if self._keys_sorted is None:
            self._keys_sorted = np.unique(self.user.astype(np.int64) * self._span + self.item.astype(np.int64))
        return self._keys_sorted

    @_keys.setter
    def _keys(self, value):
        self._keys_sorted = value

    def key_table(self):
        """Open-addressing hash set of the same keys for the host rejection walk (built once per period file)."""
        if self._table is None:
            from .._lib import lib
            n = len(self.user)
            size = 1 << max(4, int(2 * n - 1).bit_length())
            table = np.empty(size, dtype=np.int64)
            u = np.ascontiguousarray(self.user, dtype=np.int64); it = np.ascontiguousarray(self.item, dtype=np.int64)
            rc = lib().sml_host_keyset_build(u.ctypes.data, it.ctypes.data, n, self._span, table.ctypes.data, size)
            if rc != 0:
                raise RuntimeError("sml_host_keyset_build failed (%d)" % rc)
            self._table = table
        return self._table

    def __len__(self):
        return self.user.shape[0]

    def interacted(self, users, items):
        """True where (user, item) is an interaction of this period."""
        k = np.asarray(users, dtype=np.int64) * self._span + np.asarray(items, dtype=np.int64)
        pos = np.minimum(np.searchsorted(self._keys, k), len(self._keys) - 1)
        return self._keys[pos] == k

    def __getitem__(self, idx):
        user = self.user[idx]
        item = self.item[idx]
        neg_item = np.random.choice(self.item_all, 1)[0]
        while self.interacted(user, neg_item):
            neg_item = np.random.choice(self.item_all, 1)[0]
        return (user, item, neg_item)
