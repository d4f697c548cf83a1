"""Legacy per-user evaluation -- drop-in for the reference's evalution/evaluation.py (no importer
in the reference; API kept).  Scores come from the fused pair-score kernel via ``model(u, i)``."""
from __future__ import annotations

import torch

from .evalution_function import get_Rec_NDCG


def test_model(model, test_set, topK=10, need_pbar=False):
    """reference: evalution/evaluation.py:7-30 -> (Pre, Rec, MAP, NDCG, MRR) means; Pre/MAP/MRR are
    hard-wired to 0 in the reference (:52-55)."""
    model.eval()
    user, item, neg_item = test_set[0], test_set[1], test_set[2]
    Pres, Recs, MAPs, NDCGs, MRRs = [], [], [], [], []
    for i in range(len(user)):
        (Pre, Rec, MAP, NDCG, MRR) = evalution_for_user(model, user[i], item[i], neg_item[i], topK)
        Pres.append(float(Pre)); Recs.append(float(Rec)); MAPs.append(float(MAP)); NDCGs.append(float(NDCG)); MRRs.append(float(MRR))
    t = lambda x: torch.tensor(x, dtype=torch.float32).mean()
    return (t(Pres), t(Recs), t(MAPs), t(NDCGs), t(MRRs))


def evalution_for_user(model, u, item_list, neg_list, topK):
    """reference: evalution/evaluation.py:34-60: positives first, then negatives; top-K by score."""
    item_list = list(item_list)
    item_len = len(item_list)
    target_list = torch.arange(0, item_len)
    all_item = item_list + list(neg_list)
    dev = model.user_laten.weight.device
    users = torch.full((len(all_item),), int(u), dtype=torch.int64, device=dev)
    items = torch.tensor(all_item, dtype=torch.int64, device=dev)
    _, _, pre = model(users, items)
    (_, ranklist) = torch.topk(pre, topK)
    target_list = target_list.to(ranklist.device)
    Pre = torch.tensor([0.0])
    Rec, dcg = get_Rec_NDCG(ranklist, target_list)
    Ap = torch.tensor([0.0])
    rr = torch.tensor([0.0])
    return (Pre, Rec, Ap, dcg, rr)
