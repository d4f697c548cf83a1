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
        self.item_all = np.unique(self.item)
        # user -> items of this period, as a sorted key array for vectorised membership tests
        n_item_span = int(self.item.max()) + 1
        self._span = n_item_span
        self._keys = np.unique(self.user.astype(np.int64) * n_item_span + self.item.astype(np.int64))

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
