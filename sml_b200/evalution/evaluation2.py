"""Candidate-list evaluation -- drop-in for the reference's evalution/evaluation2.py.

``test_model(model, test_set, ...)`` accepts what the reference accepts (any iterable of
``[b, 1+C]`` id batches, e.g. a DataLoader over ``testDataset``) and, additionally, a
``DeviceTestSet`` holding the whole ``[N, 1+C]`` array on the GPU, in which case the evaluation is
ONE fused gather-dot-rank launch plus a per-1024-row reduction (the reference's batch size,
model/transfer.py:431-435) instead of N/1024 topk round trips.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


class DeviceTestSet(object):
    """Whole test/validation file resident on the device (int64 [N, 1+C])."""

    def __init__(self, rows, emulate_reference_rng=False, batch=ops.EVAL_BATCH):
        if not (isinstance(rows, torch.Tensor) and rows.is_cuda and rows.dtype == torch.int64 and rows.dim() == 2):
            raise TypeError("DeviceTestSet needs a 2-D int64 CUDA tensor")
        self.rows = rows.contiguous()
        self.batch = batch
        self.emulate_reference_rng = emulate_reference_rng
        self.frozen = False              # set by callers that KNOW the tables do not change between calls
        self._rank_cache = None          # (gt, eq) of the last scoring pass, reused only while frozen

    def __len__(self):
        return self.rows.shape[0]

    def ranks(self, model):
        """One scoring pass.  The kernels update the tables through raw pointers, so tensor version
        counters cannot detect changes: the pass is only reused (K = 20 / 10 / 5 of the real test,
        model/transfer.py:855-868) while the caller holds ``frozen``."""
        if self.frozen and self._rank_cache is not None:
            return self._rank_cache
        uw, iw = model.user_laten.weight.data, model.item_laten.weight.data
        self._rank_cache = ops.eval_candidates(uw, iw, self.rows)
        return self._rank_cache


def test_model(model, test_set, old_user=None, old_item=None, topK=10, need_pbar=False):
    """reference: evalution/evaluation2.py:8-26 -> (recall, ndcg) = (sum hits / N, sum ndcg / N).
    recall is a python float, ndcg a 0-d CPU float32 tensor (the reference's callers do
    ``ndcg.cpu().numpy()``, model/transfer.py:813)."""
    model.eval()
    if isinstance(test_set, DeviceTestSet):
        if test_set.emulate_reference_rng:
            from ..data.batching import ReferenceStream
            ReferenceStream.loader_iter()              # the reference creates one DataLoader iterator per call
        n = len(test_set)
        if n == 0:
            return 0.0, torch.tensor(0.0)
        gt, eq = test_set.ranks(model)
        hits, ndcg = ops.eval_reduce(gt, eq, topK, batch=test_set.batch)
        res = torch.stack([hits.sum().float(), ndcg.sum()]).cpu()       # one D2H read
        return float(res[0]) / n, (res[1] / n)
    num_test = 0
    recall_all = 0.0
    ndcg_all = 0.0
    for batch_idx, datas in enumerate(test_set):
        datas = torch.as_tensor(datas).long().to(model.user_laten.weight.device)
        batch_hit, batch_ndcg, _ = model.test(datas, topK=topK)
        recall_all += batch_hit
        ndcg_all += float(batch_ndcg)
        num_test += datas.shape[0]
    return recall_all / num_test, torch.tensor(ndcg_all / num_test, dtype=torch.float32)


def test_model_pre(model, test_set, old_user=None, new_user=None, old_item=None, new_item=None, topK=10, need_pbar=False):
    """reference: evalution/evaluation2.py:28-70 (old/new user-item breakdown of the hits; analysis
    helper with no caller on the SML path)."""
    model.eval()
    num_test = 0
    recall_all, ndcg_all = 0.0, 0.0
    counts = dict(oo=0, on=0, no=0, nn=0)
    for batch_idx, datas in enumerate(test_set):
        datas = torch.as_tensor(datas).long().to(model.user_laten.weight.device)
        idx, rank, batch_hit, batch_ndcg = model.test2(datas)
        hit_inter = datas[idx][:, 0:2].cpu().numpy()
        for u, i in hit_inter:
            k = ("o" if u in old_user else "n") + ("o" if i in old_item else "n")
            counts[k] += 1
        recall_all += batch_hit
        ndcg_all += float(batch_ndcg)
        num_test += datas.shape[0]
    all_hit = max(1, sum(counts.values()))
    print("old user old item:", counts["oo"] * 1.0 / all_hit, counts["oo"] * 1.0 / num_test)
    print("old user new item", counts["on"] * 1.0 / all_hit, counts["on"] * 1.0 / num_test)
    print("new user old item", counts["no"] * 1.0 / all_hit, counts["no"] * 1.0 / num_test)
    print("new user new item", counts["nn"] * 1.0 / all_hit, counts["nn"] * 1.0 / num_test)
    print("num test:", num_test)
    return recall_all / num_test, np.float32(ndcg_all / num_test)
