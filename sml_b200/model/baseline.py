"""MF baselines on the fused plain-MF kernels: the comparison methods of the reference's
model/baseline.py -- fine-tuning on each new period and full retraining on all periods seen so far -- reuse
``MFbasemode`` and the BCE objective of ``base_train`` (model/baseline.py:161-225: mean log-sigmoid terms + 0.5 * l2 * ||.||^2,
dense Adam).  One step = sml_plain_mf_grads (gather - dot - loss - scatter) + sml_adam_dense x2; negatives come from the GPU
Philox sampler with the reference's exclusion rule (items of the training window the user has not interacted with).
SPMF's reservoir (model/baseline.py:68-100,448-476) is not rebuilt.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..data.dataset import offlineDataset_withsample
from ..evalution.evaluation2 import DeviceTestSet, test_model
from . import MF


class MFTrainer(object):
    def __init__(self, num_user, num_item, laten=64, lr=0.001, l2_u=1e-4, l2_i=1e-4, batch_size=1024, device=None, seed=2000):
        ops.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.MFbase = MF.MFbasemode(num_user, num_item, laten).to(self.device)
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        z = torch.zeros_like
        self._st = dict(m_u=z(uw), v_u=z(uw), g_u=z(uw), m_i=z(iw), v_i=z(iw), g_i=z(iw))
        self.adam_state = ops.new_adam_state(self.device)
        self.lr, self.l2_u, self.l2_i, self.batch_size, self.seed = lr, l2_u, l2_i, batch_size, seed
        self.loss = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._epoch = 0

    def train_epoch(self, interactions):
        """One pass over ``interactions`` ([N, 2] numpy, (user, item)) -> mean loss per batch."""
        ds = offlineDataset_withsample.__new__(offlineDataset_withsample)
        ds.user, ds.item = interactions[:, 0], interactions[:, 1]
        ds.item_all = np.unique(ds.item)
        ds._span = int(ds.item.max()) + 1
        ds._keys = np.unique(ds.user.astype(np.int64) * ds._span + ds.item.astype(np.int64))
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(self.device)
        order = torch.randperm(len(ds.user), device=self.device)
        u, i = T(ds.user)[order].contiguous(), T(ds.item)[order].contiguous()
        self._epoch += 1
        j = ops.philox_negatives(u, T(ds.item_all), T(ds._keys), ds._span, seed=self.seed, offset=self._epoch)
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        s = self._st
        self.loss.zero_()
        nb = 0
        for b in range(0, u.numel(), self.batch_size):
            e = min(b + self.batch_size, u.numel())
            ops.plain_mf_grads(uw, iw, u[b:e], i[b:e], j[b:e], s["g_u"], s["g_i"], self.loss, loss=ops.LOSS_BCE, l2_u=self.l2_u, l2_i=self.l2_i)
            ops.adam_tick(self.adam_state, self.lr)
            ops.adam_dense(uw, s["m_u"], s["v_u"], s["g_u"], self.adam_state)
            ops.adam_dense(iw, s["m_i"], s["v_i"], s["g_i"], self.adam_state)
            nb += 1
        return self.loss[1].item() / max(nb, 1)

    def test(self, test_rows, topK=20):
        rows = torch.from_numpy(np.ascontiguousarray(test_rows, dtype=np.int64)).to(self.device)
        r, n = test_model(self.MFbase, DeviceTestSet(rows), topK=topK)
        return r, float(n)


class FineTune(MFTrainer):
    """Fine-tune on every new period only."""

    def run_period(self, train_t, epochs=1):
        return [self.train_epoch(train_t) for _ in range(epochs)]


class FullRetrain(MFTrainer):
    """Retrain on all periods seen so far."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._seen = []

    def run_period(self, train_t, epochs=1):
        self._seen.append(train_t)
        return [self.train_epoch(np.concatenate(self._seen)) for _ in range(epochs)]
