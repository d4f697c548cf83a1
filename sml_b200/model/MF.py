"""MF base model -- drop-in for the reference's model/MF.py (class names, constructor signature,
attribute names and state_dict keys are the reference's; spelling included).

The four tables stay ``nn.Embedding`` modules because the reference's callers read and write
``.weight`` directly (model/transfer.py:348-361,926-933,953-959); every computation on them runs
in the sm_100a kernels of libsml_b200.so -- there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class MFbasemode(nn.Module):
    """reference: model/MF.py:18-114."""

    def __init__(self, num_user=0, num_item=0, laten_factor=10):
        super(MFbasemode, self).__init__()
        # construction order = the reference's (model/MF.py:21-24) so that the same torch seed
        # yields the same initial tables
        self.user_bais = nn.Embedding(num_user, 1)
        self.item_bais = nn.Embedding(num_item, 1)
        self.user_laten = nn.Embedding(num_user, laten_factor)
        self.item_laten = nn.Embedding(num_item, laten_factor)
        self.user_num = num_user
        self.item_num = num_item
        self.hidden_dim = laten_factor

    def reset_parameters(self):
        self.user_bais.reset_parameters()
        self.user_laten.reset_parameters()
        self.item_bais.reset_parameters()
        self.item_laten.reset_parameters()

    def forward(self, user, item, norm=False):
        """model/MF.py:34-43 -> (user rows, item rows, <u,i> [/ ||u||]).  The scores come from the
        fused gather-dot kernel; the two row tensors are plain gathers.  Not differentiable: the
        training paths of this package are the fused step kernels (ops.mf_step / ops.plain_mf_grads)."""
        user = user.long().contiguous()
        item = item.long().contiguous()
        uw, iw = self.user_laten.weight.data, self.item_laten.weight.data
        result = ops.pair_scores(uw, iw, user.reshape(-1), item.reshape(-1), norm=norm).reshape(user.shape)
        return uw[user], iw[item], result

    def _ranks(self, inputs_data):
        rows = inputs_data.long().contiguous()
        return ops.eval_candidates(self.user_laten.weight.data, self.item_laten.weight.data, rows)

    def test(self, inputs_data, topK=20):
        """model/MF.py:45-80: column 0 = user, columns 1.. = candidates (candidate 0 is the positive).
        Returns (n_hits float, sum of 1/log2(rank+2) over hits (0-d tensor, or int 0 when there is
        no hit, as the reference), int64 indices of the hit rows)."""
        gt, eq = self._ranks(inputs_data)
        n = gt.numel()
        hits, ndcg = ops.eval_reduce(gt, eq, topK, batch=max(n, 1))
        rank = gt + eq
        hit_rows = (rank < topK).nonzero()[:, 0]
        have_hit_num = int(hits[0].item()) if n else 0
        batch_NDCG = ndcg[0] if have_hit_num > 0 else 0
        return have_hit_num * 1.0, batch_NDCG, hit_rows

    def test2(self, inputs_data, topK=20):
        """model/MF.py:82-108 -> (hit row idx, top-K candidate indices, n_hits, ndcg).  The rank list
        is only needed by an analysis helper (evaluation2.test_model_pre); scores come from the
        pair-score kernel and the ordering from a sort of the [n, C] score matrix."""
        rows = inputs_data.long().contiguous()
        n, w = rows.shape
        user = rows[:, :1].expand(n, w - 1).reshape(-1).contiguous()
        item = rows[:, 1:].reshape(-1).contiguous()
        scores = ops.pair_scores(self.user_laten.weight.data, self.item_laten.weight.data, user, item).reshape(n, w - 1)
        _, rank = torch.topk(scores, topK)
        hit, ndcg, idx = self.test(inputs_data, topK)
        return idx, rank, hit, ndcg

    def set_parameters(self, user_weight, item_weight):
        self.user_laten.weight.data.copy_(user_weight[:, 0:-1])
        self.user_bais.weight.data.copy_(user_weight[:, -1].unsqueeze(-1))
        self.item_laten.weight.data.copy_(item_weight[:, 0:-1])
        self.item_bais.weight.data.copy_(item_weight[:, -1].unsqueeze(-1))


class MF2(nn.Module):
    """reference: model/MF.py:118-156 (true BPR with biases; no caller in the reference).  The
    training branch returns the BPR loss and the reference's l2 value computed by the fused
    plain-MF kernel; gradients land in ``self.grads`` (dense buffers) rather than autograd."""

    def __init__(self, num_user=0, num_item=0, laten_factor=10):
        super(MF2, self).__init__()
        self.user_bais = nn.Embedding(num_user, 1)
        self.item_bais = nn.Embedding(num_item, 1)
        self.user_laten = nn.Embedding(num_user, laten_factor)
        self.item_laten = nn.Embedding(num_item, laten_factor)
        self.user_num = num_user
        self.item_num = num_item
        self.hidden_dim = laten_factor
        self.grads = None

    def forward(self, user, item, neg_item=None):
        uw, iw = self.user_laten.weight.data, self.item_laten.weight.data
        user = user.long().contiguous(); item = item.long().contiguous()
        if neg_item is not None:
            neg_item = neg_item.long().contiguous()
            if self.grads is None:
                self.grads = dict(user=torch.zeros_like(uw), item=torch.zeros_like(iw),
                                  item_bias=torch.zeros_like(self.item_bais.weight.data),
                                  loss=torch.zeros(2, device=uw.device))
            g = self.grads
            ops.plain_mf_grads(uw, iw, user, item, neg_item, g["user"], g["item"], g["loss"], loss=ops.LOSS_BPR,
                               item_bias=self.item_bais.weight.data.reshape(-1), g_item_bias=g["item_bias"].reshape(-1))
            l2loss = (torch.norm(uw[user], dim=-1).sum() + torch.norm(iw[item], dim=-1).sum() + torch.norm(iw[neg_item]).sum())
            return g["loss"][0].clone(), l2loss
        result = ops.pair_scores(uw, iw, user, item) + self.user_bais.weight.data[user, 0] + self.item_bais.weight.data[item, 0]
        return uw[user], iw[item], result
