"""Rank-list metrics -- drop-in for the reference's evalution/evalution_function.py.
``ranklist``: indices of the top-K candidates; ``target_items``: arange(n_pos) (the positives are
the first n_pos candidates).  Tiny per-user host-side tensor expressions, as in the reference."""
from __future__ import annotations

import torch


def _rank_of_target(ranklist, target_items):
    return torch.nonzero(ranklist < (target_items[-1] + 1))[:, 0]


def hit(ranklist, target_items):
    return _rank_of_target(ranklist, target_items).shape[0]


def IDCG(n):
    """reference: evalution_function.py:86-95."""
    arr = torch.arange(n).float() + 2
    return (1.0 / torch.log2(arr)).sum()


def get_Rec_NDCG(ranklist, target_items):
    """reference: evalution_function.py:26-39."""
    idcg = IDCG(target_items.shape[0])
    rank_of_target = _rank_of_target(ranklist, target_items)
    hits = rank_of_target.shape[0]
    if hits > 0:
        dcg = (1.0 / torch.log2(rank_of_target.float() + 2)).sum() / idcg.to(ranklist.device)
    else:
        dcg = 0
    return hits / target_items.shape[0], dcg


def get_Precision(ranklist, target_items, topK):
    return hit(ranklist, target_items) / topK


def get_Recall(ranklist, target_items, topK=10):
    return hit(ranklist, target_items) / target_items.shape[0]


def get_NDCG(ranklist, target_items):
    return get_Rec_NDCG(ranklist, target_items)[1]


def get_MRR(ranklist, target_items):
    r = _rank_of_target(ranklist, target_items)
    return 1.0 / (r[0] + 1).float() if r.shape[0] > 0 else 0


def get_MAP(ranklist, target_items):
    r = _rank_of_target(ranklist, target_items).float()
    if r.shape[0] > 0:
        r = r + 1
        hits = torch.arange(r.shape[0]).float().to(ranklist.device) + 1
        return torch.sum(hits / r) / (min(ranklist.shape[0], target_items.shape[0]) * 1.0)
    return 0
