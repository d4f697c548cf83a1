"""MF baselines -- drop-in for the reference's model/baseline.py: SPMF (reservoir +
rank-weighted sampling), full-retrain MF and fine-tune MF, plus the pre-training loop ``base_train``.

Same class / function names, constructor signatures, attributes, period state machine and printed summaries as the
reference (``SPMF`` model/baseline.py:102-556, ``Reservious`` :68-100, ``StreamingData`` :558-587, ``test_hit_new``
:18-30, ``get_parse`` :592-626), but every device operation is a kernel of libsml_b200.so:

  reference (stock PyTorch)                                        here
  ---------------------------------------------------------------  -----------------------------------------------
  2 x MFbase.forward + BCE + l2 + backward + Adam.step (:188-201,  ops.plain_mf_step: ONE fused kernel per batch
  :269-280, :348-361; dense Adam over both tables)                 (gather - dot - loss - row grads + l2 - exact dense
                                                                   Adam in its row-lazy form, csrc/plain_mf_step.cu)
  MFbase.test per 1024-row batch and per K (:388-443)              one ops.eval_candidates pass, reduced for every K
  MFbase.forward over the pool for the sampling weights (:448-476) ops.pair_scores

Batches: the reference draws them through ``DataLoader(shuffle=True, num_workers=4)`` over ``offlineDataset_withsample``
(per-sample ``np.random.choice`` rejection); here a whole epoch of (user, item, neg) is built at once with the same
consumption of the global torch / numpy generators as ``num_workers=0`` (sml_b200.data.batching.ReferenceStream), or
supplied by the caller (``batch_source``: "the same supplied negative-sample indices").

Reference defects that make parts of it unrunnable as shipped, and what this module does instead:
  * ``np.long`` (:73,117,566,...) was removed from numpy 1.24 -> int64;
  * ``run_one_stage`` unpacks the 4-tuple of ``self.test`` into two names (:250) and raises ValueError, and ``__main__``
    calls ``base_train_not_train(start_idx-1)`` with an undefined name (:667): the SPMF method cannot run in the reference.
    Here ``run_one_stage`` unpacks all four values; everything else follows the reference line by line;
  * ``base_train`` saves checkpoints to a hard-coded absolute path (:213,219): here to ``args.save_dir`` when given.
There is no CPU / eager fallback: without the CUDA library every step raises.
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch

from .. import ops
from ..data.batching import ReferenceStream, tr_epoch_triples
from ..data.dataset import offlineDataset_withsample as _SampleDataset
from . import MF
from .transfer import FusedAdam


def test_hit_new(data, have_idx, new_user, new_item):
    """reference: model/baseline.py:18-30 -- how many of the hit rows belong to a new user / a new item."""
    hit_itr = data[have_idx][:, 0:2]
    hit_new_user = int(torch.isin(hit_itr[:, 0], new_user).sum())
    hit_new_item = int(torch.isin(hit_itr[:, 1], new_item).sum())
    return hit_new_user, hit_new_item


class offlineDataset_withsample(_SampleDataset):
    """reference: model/baseline.py:32-66 (same sampler as data/dataset.py:41-71, plus the ``neg_num`` argument; like the
    reference, ``__getitem__`` returns ONE negative whatever neg_num is, :59-66)."""

    def __init__(self, dataset, neg_num=1):
        super().__init__(dataset)
        self.neg_num = neg_num


class Reservious(object):
    """reference: model/baseline.py:68-100 -- reservoir of past interactions, quirks included (``pool_have`` is advanced
    by ``max_id``, not by the number of rows copied, :84)."""

    def __init__(self, length):
        self.t = 0
        self.len = length
        self.pool = np.zeros((length, 2), dtype=np.int64)
        print("pool size:", self.pool.shape)
        self.pool_have = 0

    def updata(self, new_data):
        if self.t <= self.len:
            new_num = new_data.shape[0]
            max_id = min(self.len, self.pool_have + new_num)
            self.pool[self.pool_have:max_id] = new_data[:max_id - self.pool_have]
            if max_id != self.len:
                new_data = new_data[max_id - self.pool_have:]
            self.pool_have = self.pool_have + max_id
            self.t = max_id
        new_num = new_data.shape[0]
        p = self.len * 1.0 / (self.t + np.arange(new_num) + 1)
        m = np.random.rand(new_num)
        select_data = new_data[np.where(m < p)]
        for i in range(select_data.shape[0]):
            idx = np.random.randint(0, self.len, 1)
            self.pool[idx] = select_data[i]
        self.t += new_num

    def init_pool(self, new_data):
        num = new_data.shape[0]
        rand_idx = np.random.randint(0, num, self.len)          # drawn and unused, like the reference (:96)
        del rand_idx
        self.pool[:] = new_data[-self.len:]
        self.pool_have = self.len
        self.t = num


class SPMF(object):
    """MF-based baselines: SPMF, full-retrain MF, fine-tune MF (reference: model/baseline.py:102-556)."""

    def __init__(self, args, datasets, user_num, item_num, laten_dim, device=None, batch_source=None, emulate_reference_rng=True):
        """``batch_source``: optional callable (stage_id, epoch, n_rows) -> (user, item, neg) int64 arrays in batch order for
        the DataLoader-fed loops (base_train, run_one_stage2).  ``emulate_reference_rng``: draw batches with the reference's
        consumption of the global generators (num_workers = 0)."""
        ops.lib()                                                # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise RuntimeError("sml_b200 baselines need a CUDA device (no CPU path)")
        if int(laten_dim) != 64:
            raise ValueError("sml_b200 kernels are specialised for laten_dim=64 (the reference default)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.MFbase = MF.MFbasemode(num_user=int(user_num), num_item=int(item_num), laten_factor=laten_dim).to(self.device)
        print("args lr:", args.lr)
        self.lr = args.lr
        self.optimizer = FusedAdam(self.MFbase.parameters(), lr=args.lr, weight_decay=0)
        print("optimizer state:", self.optimizer.state_dict())
        self.pool_size = args.pool_size
        self.Reservious = Reservious(self.pool_size)
        print("self.pool_size:", self.pool_size)
        self.all_item = np.ones(0, dtype=np.int64)
        self.dataset = datasets
        self.new_user = torch.from_numpy(np.asarray(datasets.test_new_user)).long().to(self.device)
        self.new_item = torch.from_numpy(np.asarray(datasets.test_new_item)).long().to(self.device)
        self.neg_num = args.neg_num
        if int(self.neg_num) != 1:
            raise NotImplementedError("neg_num != 1 is not supported (reference default 1; its DataLoader path ignores it too)")
        self.batch_size = args.batch_size
        self.lambda_u = args.l2_u
        self.lambda_i = args.l2_i
        self.recall = []
        self.ndcg = []
        self.hit_new_user = []
        self.hit_new_item = []
        self.epochs = args.epochs
        self.run_stage = 0
        self.test_num = []
        self.user_hit = None
        self.pool_init_type = args.pool_init_type
        self.save_dir = getattr(args, "save_dir", None)
        self.batch_source = batch_source
        self.emulate_reference_rng = emulate_reference_rng
        # optimizer state of the fused step: dense Adam (model/baseline.py:111) in its bit-identical row-lazy form
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        z = torch.zeros_like
        self._st = dict(m_u=z(uw), v_u=z(uw), m_i=z(iw), v_i=z(iw))
        self.optimizer.adam_state = ops.new_adam_state(self.device, history=True)
        self._stamp_u = ops.new_row_stamps(uw.shape[0], self.optimizer.adam_state)
        self._stamp_i = ops.new_row_stamps(iw.shape[0], self.optimizer.adam_state)
        self._head_u = ops.new_list_heads(uw.shape[0], self.device)
        self._head_i = ops.new_list_heads(iw.shape[0], self.device)
        self.optimizer.state = {self.MFbase.user_laten.weight: dict(exp_avg=self._st["m_u"], exp_avg_sq=self._st["v_u"]),
                                self.MFbase.item_laten.weight: dict(exp_avg=self._st["m_i"], exp_avg_sq=self._st["v_i"])}
        self._loss = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._since_flush = 0

    # ------------------------------------------------------------------ device helpers
    def _flush(self):
        """Every row up to date with the dense optimizer before the tables are read as a whole."""
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        ops.adam_flush(uw, self._st["m_u"], self._st["v_u"], self._stamp_u, self.optimizer.adam_state)
        ops.adam_flush(iw, self._st["m_i"], self._st["v_i"], self._stamp_i, self.optimizer.adam_state)
        self._since_flush = 0

    def _step(self, u, i, j, l2_u, l2_i):
        """One batch: forward of (u, i) and (u, j), loss = -bce + l2, backward, Adam.step (model/baseline.py:186-201)."""
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        s = self._st
        ops.plain_mf_step(uw, iw, s["m_u"], s["v_u"], s["m_i"], s["v_i"], self._head_u, self._head_i, u, i, j, self.optimizer.adam_state,
                          self.optimizer.param_groups[0]["lr"], self._loss, loss=ops.LOSS_BCE, l2_u=l2_u, l2_i=l2_i,
                          optimizer=ops.OPT_ADAM_DENSE_EXACT, stamp_user=self._stamp_u, stamp_item=self._stamp_i)
        self._since_flush += 1
        if self._since_flush >= ops.ADAM_HISTORY // 2:           # before the ring of per-step scalars wraps
            self._flush()

    def _epoch_triples(self, ds, stage_id, epoch):
        if self.batch_source is not None:
            u, i, j = self.batch_source(stage_id, epoch, len(ds))
        else:
            order = ReferenceStream.shuffled_order(len(ds)) if self.emulate_reference_rng else np.random.permutation(len(ds))
            u, i, j = tr_epoch_triples(ds, order)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(self.device)
        return T(u), T(i), T(j)

    def _train_epoch(self, ds, stage_id, epoch, l2_u, l2_i):
        """One DataLoader pass (model/baseline.py:181-201,340-361) -> (sum of the batch losses, batches)."""
        u, i, j = self._epoch_triples(ds, stage_id, epoch)
        self._loss.zero_()
        B = int(self.batch_size)
        nb = 0
        for b in range(0, u.numel(), B):
            self._step(u[b:b + B], i[b:b + B], j[b:b + B], l2_u, l2_i)
            nb += 1
        return self._loss[1].item(), nb

    # ------------------------------------------------------------------ reference methods
    def get_next_data(self, stage_id, types="only_new"):
        set_t, now_test = self.dataset.get_next(stage_id, types=types)
        return set_t, now_test

    def base_train_not_train(self, stage_id):
        set_t, now_test = self.get_next_data(stage_id, types="not_only_new")
        if self.pool_init_type == 1:
            self.Reservious.init_pool(set_t)
        F_recall, F_ndcg, _, _ = self.test(now_test)
        print("before train test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg)
        if self.pool_init_type == 0:
            self.updata_reservious(set_t)

    def init_pool(self):
        pass

    def base_train(self, stage_id, epochs, l2_u, l2_i):
        """Pre-training of the MF model the baselines start from (reference: model/baseline.py:161-225)."""
        print("********base train: (l2_u,l2_i): ({},{})*****".format(l2_u, l2_i))
        set_t, now_test = self.get_next_data(stage_id, types="not_only_new")
        train = offlineDataset_withsample(set_t, neg_num=1)
        max_recall20 = 0
        max_REC = None
        max_Ndcg = None
        max_epoch = 0
        not_change_num = 0
        self.base_losses = []
        for epoch in range(epochs):
            self.MFbase.train()
            s_time = time.time()
            loss_sum, nb = self._train_epoch(train, stage_id, epoch, l2_u, l2_i)
            loss_all = loss_sum / (nb * self.batch_size)                       # (bat_num+1)*batch_size (:202)
            self.base_losses.append(loss_all)
            print("epoch:{}, time:{:.1f}, loss:{:.4f}".format(epoch, time.time() - s_time, loss_all))
            if (epoch % 2) == 0:
                F_recall, F_ndcg, _, _ = self.test(now_test)
                not_change_num += 1
                if F_recall[-1] > max_recall20:
                    max_recall20 = F_recall[-1]
                    max_REC = F_recall
                    max_Ndcg = F_ndcg
                    max_epoch = epoch
                    not_change_num = 0
                    self._save("best-mean-start29-spmf-" + "-" + str(l2_u) + "-" + str(self.lr) + "lr.pt")
                print("test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg, "max reccall", max_REC, "max epoch:", max_epoch)
                if not_change_num > 50:
                    print("max not change up to 20 epochs, stop ......")
                    break
            if epoch % 50 == 0:
                self._save("mean-start29-spmf-" + str(epoch) + "-" + str(l2_u) + "-" + str(self.lr) + "lr.pt")
        F_recall, F_ndcg, _, _ = self.test(now_test)
        print("FInal test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg)
        print("max: epoch", max_epoch, "max_recall:", max_REC, "max_Ndcg", max_Ndcg)

    def _save(self, name):
        if self.save_dir:
            self._flush()
            os.makedirs(self.save_dir, exist_ok=True)
            torch.save(self.MFbase.state_dict(), os.path.join(self.save_dir, name))

    def _begin_stage(self, stage_id, types):
        """What both per-period methods do first (reference: model/baseline.py:233-245,313-326): next (train, test) pair, the
        item universe, reservoir + new data as the training set, the user -> items index the negative samplers exclude."""
        set_t, now_test = self.get_next_data(stage_id, types=types)
        if set_t is None:
            return None
        self.test_num.append(now_test.shape[0])
        self.all_item = np.union1d(self.all_item, set_t[:, 1])
        have = self.Reservious.pool_have
        train_data = np.concatenate([self.Reservious.pool[0:have], set_t], axis=0) if have > 0 else set_t
        self.user_hit_num_in_W_R(train_data)
        return set_t, now_test, train_data

    class _Best(object):
        """Best recall@20 so far and the evaluations since it improved (the early-stopping bookkeeping of :252-255,289-299,
        :334-337,366-378; the reference only ever stops early on the news data, pool_init_type = 1)."""

        def __init__(self):
            self.recall20, self.recall, self.ndcg, self.stale = 0, None, None, 0

        def update(self, F_recall, F_ndcg):
            if self.recall20 < F_recall[-1]:
                self.recall20, self.recall, self.ndcg, self.stale = F_recall[-1], F_recall, F_ndcg, 0

    def run_one_stage(self, stage_id):
        """SPMF, one period (reference: model/baseline.py:227-304): train on reservoir + new data with rank-weighted
        sampling, then update the reservoir."""
        begun = self._begin_stage(stage_id, "only_new")
        if begun is None:
            return False
        set_t, now_test, train_data = begun
        itr = round(train_data.shape[0] / self.batch_size)
        p = self.compute_R_W_P(train_data)
        print("start train...")
        F_recall, F_ndcg, _, _ = self.test(now_test)         # (the reference unpacks two of the four values here and raises, :250)
        print("before train test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg)
        best = self._Best()
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64).reshape(-1)).to(self.device)
        for epoch in range(self.epochs):
            self.MFbase.train()
            s_time = time.time()
            self._loss.zero_()
            for bat_num in range(itr):
                bat_user, bat_item, bat_neg = self.sample_batch(train_data, self.batch_size, p, self.neg_num)
                self._step(T(bat_user), T(bat_item), T(bat_neg), self.lambda_u, self.lambda_i)
            loss_all = self._loss[1].item() / max(itr, 1)
            print("epoch: {} ,time:{:.1f}, loss:{:.4f}".format(epoch, time.time() - s_time, loss_all))
            best.stale += 1                                   # tested after every epoch (:287)
            F_recall, F_ndcg, _, _ = self.test(now_test)
            print("        epoch test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg)
            best.update(F_recall, F_ndcg)
            if best.stale >= 5 and self.pool_init_type == 1:
                break
        self.updata_reservious(set_t)
        F_recall, F_ndcg, hit_new_user, hit_new_item = self.test(now_test)
        print("FInal test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg, "hit new user:", hit_new_user, "hit new item:", hit_new_item)
        self.recall.append(F_recall)
        self.ndcg.append(F_ndcg)
        return True

    def run_one_stage2(self, stage_id, read_data_type="only_new"):
        """Full-retrain ("not_only_new") and fine-tune ("only_new"), one period (reference: model/baseline.py:306-386)."""
        begun = self._begin_stage(stage_id, read_data_type)
        if begun is None:
            return False
        _, now_test, train_data = begun
        if self.Reservious.pool_have > 0:
            print("pool having.....")
        train = offlineDataset_withsample(train_data)
        print("start train...")
        F_recall, F_ndcg, _, _ = self.test(now_test)
        print("before train test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg)
        best = self._Best()
        self.stage_losses = []
        for epoch in range(self.epochs):
            self.MFbase.train()
            s_time = time.time()
            loss_sum, nb = self._train_epoch(train, stage_id, epoch, self.lambda_u, self.lambda_i)
            loss_all = loss_sum / nb                                           # / (bat_num+1) (:362)
            self.stage_losses.append(loss_all)
            print("epoch: {} ,time:{:.1f}, loss:{:.4f}".format(epoch, time.time() - s_time, loss_all))
            best.stale += 1
            if epoch % 5 == 0:                                # tested every fifth epoch (:365)
                F_recall, F_ndcg, _, _ = self.test(now_test)
                print("        epoch test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg)
                best.update(F_recall, F_ndcg)
                if best.stale > 5 and self.pool_init_type == 1:
                    break
        F_recall, F_ndcg, hit_newu, hit_newi = self.test(now_test, stage_idx=stage_id)
        print("max result ", best.recall, best.ndcg)
        print("FInal test---", "recall(5,10,20):", F_recall, "ndcg (5,10,20):", F_ndcg, "hit user:", hit_newu, "hit item:", hit_newi)
        self.recall.append(F_recall)
        self.ndcg.append(F_ndcg)
        self.hit_new_user.append(hit_newu)
        self.hit_new_item.append(hit_newi)
        return True

    def test(self, test_data, topk=[5, 10, 20], stage_idx=None):
        """reference: model/baseline.py:388-443 -> (recall[len(topk)], ndcg[len(topk)], hits on new users / N, hits on new
        items / N); the last two are counted at the LAST K of ``topk``, like the reference (:420-421)."""
        self.MFbase.eval()
        self._flush()
        test_num = test_data.shape[0]
        rows = test_data if isinstance(test_data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(test_data, dtype=np.int64))
        rows = rows.to(self.device, dtype=torch.int64)
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        gt, eq = ops.eval_candidates(uw, iw, rows)
        recall, ndcg = [], []
        for k in topk:
            hits, nd = ops.eval_reduce(gt, eq, k, batch=1024)                  # per 1 024-row batch, like :400-404
            recall.append(float(hits.sum().item()))
            ndcg.append(float(nd.sum().item()))
        hit = (gt + eq) < topk[-1]
        hit_u, hit_i = test_hit_new(rows, hit, self.new_user, self.new_item)
        return (np.array(recall) / test_num, np.array(ndcg, dtype=np.float32) / test_num, np.float64(hit_u * 1.0) / test_num,
                np.float64(hit_i * 1.0) / test_num)

    def updata_reservious(self, train_data):
        self.Reservious.updata(train_data)

    def compute_R_W_P(self, R_TR_data):
        """Sampling probabilities from the rank of every interaction's score (reference: model/baseline.py:448-476)."""
        self.MFbase.eval()
        self._flush()
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(self.device)
        score = ops.pair_scores(uw, iw, T(R_TR_data[:, 0]), T(R_TR_data[:, 1]))
        rank_idx = torch.argsort(score, descending=True)
        num = rank_idx.shape[0]
        rank_ = (torch.arange(num, device=score.device) + 1)
        p = torch.zeros_like(score)
        p[rank_idx] = rank_.float()
        w = torch.exp(p * 1.0 / num)
        p = w / w.sum()
        return p.cpu().numpy()

    def user_hit_num_in_W_R(self, data):
        if self.user_hit is None:
            self.user_hit = {}
        for k in range(data.shape[0]):
            u = data[k, 0]
            i = data[k, 1]
            try:
                self.user_hit[u].add(i)
            except KeyError:
                self.user_hit[u] = set([i])

    def sample_batch(self, data, batch_size, p, neg_num):
        """reference: model/baseline.py:489-503 (same draws from the global numpy generator)."""
        idx = np.arange(data.shape[0])
        bat_idx = np.random.choice(idx, batch_size, p=p)
        bat_data = data[bat_idx]
        bat_user = bat_data[:, 0]
        bat_item = bat_data[:, 1]
        bat_neg = []
        for i in range(bat_data.shape[0]):
            m = np.random.choice(self.all_item, neg_num)
            u = bat_user[i]
            while m[0] in self.user_hit[u]:
                m = np.random.choice(self.all_item, neg_num)
            bat_neg.append(m)
        bat_neg = np.array(bat_neg)
        return bat_user.reshape(-1, 1), bat_item.reshape(-1, 1), bat_neg

    def run(self, start_stage, method="full"):
        """reference: model/baseline.py:505-556: periods until the stream ends, then the weighted summaries (first third of
        the test periods = validation, the rest = test, all periods)."""
        self.run_stage = 0
        stage_id = start_stage
        self.summary = {}
        step = {"spmf": self.run_one_stage,
                "full": lambda sid: self.run_one_stage2(sid, read_data_type="not_only_new")}.get(
                    method, lambda sid: self.run_one_stage2(sid, read_data_type="only_new"))
        while True:
            print("#################################runing stage:{}########################".format(stage_id))
            if not step(stage_id):
                break
            stage_id += 1
            self.run_stage += 1
        test_num = np.array(self.test_num).reshape(-1, 1)
        recall, ndcg = np.array(self.recall), np.array(self.ndcg)
        print("average recall:", recall.mean(axis=0))
        print("average recall:", ndcg.mean(axis=0))               # sic (:521)
        print(test_num)
        print(recall)
        print(ndcg)
        print("hit new user:", self.hit_new_user)
        print("hit new item:", self.hit_new_item)
        N3 = round(test_num.shape[0] * 1.0 / 3)

        def weighted(lo, hi):
            w = test_num[lo:hi] / test_num[lo:hi].sum()
            return (recall[lo:hi] * w).sum(axis=0), (ndcg[lo:hi] * w).sum(axis=0)
        val, tst, both = weighted(0, N3), weighted(N3, None), weighted(0, None)
        print("pre 3 (val) reslut,recall,ndcg:", val[0], val[1])
        print("last 7 (test) results,recall ,ndcg:", tst[0], tst[1])
        self.summary = dict(val_recall=val[0], val_ndcg=val[1], test_recall=tst[0], test_ndcg=tst[1], recall=both[0], ndcg=both[1])
        print("weight average recall@20:", both[0])
        print("weight average ndcg@20:", both[1])


class StreamingData(object):
    """reference: model/baseline.py:558-587 (same on-disk layout: information.npy, test_new_user.npy, test_new_item.npy,
    train/<p>.npy, test/<p>.npy)."""

    def __init__(self, file_pathe):
        information = np.load(file_pathe + "information.npy")
        self.user_num = information[1]
        self.item_num = information[2]
        self.itr_num = information[0]
        self.path = file_pathe
        self.test_new_user = np.load(file_pathe + "test_new_user.npy").astype(np.int64)
        self.test_new_item = np.load(file_pathe + "test_new_item.npy").astype(np.int64)

    def get_next(self, stage_id, types="not_only_new"):
        try:
            if types == "not_only_new":
                train_data = []
                for i in range(0, stage_id):
                    train_data.append(np.load(self.path + "train/" + str(i) + ".npy").astype(np.int64))
                train_data = np.concatenate(train_data, axis=0)
            else:
                train_data = np.load(self.path + "train/" + str(stage_id - 1) + ".npy").astype(np.int64)
        except (OSError, ValueError):
            print("read train data roung , may be there is no new data,finished")
            return None, None
        try:
            test_data = np.load(self.path + "test/" + str(stage_id) + ".npy").astype(np.int64)
        except (OSError, ValueError):
            print("read test data roung , may be there is no new data,finished")
            return None, None
        print("NOTICED: will train: {} , will test:{} ".format(stage_id - 1, stage_id))
        return train_data, test_data


def get_parse():
    """The reference's flags and defaults (model/baseline.py:592-626) + --save_dir."""
    parser = argparse.ArgumentParser(description="MF and TR parameters.")
    parser.add_argument("--lr", type=float, default=0.01, help="Learning rate.")
    parser.add_argument("--l2_u", type=float, default=1e-5, help="user l2. should be same to l2_i")
    parser.add_argument("--l2_i", type=float, default=1e-5, help="item l2.should be same to l2_u ")
    parser.add_argument("--epochs", type=int, default=20, help="Number of epochs to train of each stage.")
    parser.add_argument("--batch_size", type=int, default=256, help="batch size of train.")
    parser.add_argument("--laten_dim", type=int, default=64, help="batch size of train.")
    parser.add_argument("--neg_num", type=int, default=1, help="neg num.")
    parser.add_argument("--pool_size", type=int, default=0, help="batch size of train.")
    parser.add_argument("--laten", type=int, default=64, help="dim of embedding.")
    parser.add_argument("--cuda", type=int, default=1, help="which GPU be used?.default 1")
    parser.add_argument("--method", default="full", help="full, fine, spmf")
    parser.add_argument("--pool_init_type", type=int, default=0, help="Reservious of SPMF init methods, 0: update , 1: init, yelp=0, news (adressa) =1 ")
    parser.add_argument("--data_path", default="dataset/", help="data path")
    parser.add_argument("--data_name", default="yelp", help="dataset name")
    parser.add_argument("--pre_model", default="", help="pre-trained MFbasemode state_dict")
    parser.add_argument("--start_idx", type=int, default=30, help="retraining from which period: yelp 30, news(adressa) 48")
    parser.add_argument("--save_dir", default=None, help="where base_train saves checkpoints (the reference uses a hard-coded path)")
    return parser


def main(argv=None):
    """reference: model/baseline.py:628-670."""
    print("start")
    args = get_parse().parse_args(argv)
    print("parameters:", args)
    data_path = args.data_path + args.data_name + "/"
    args.pool_init_type = 1 if args.data_name == "news" else 0
    dataset = StreamingData(data_path)
    user_num, item_num, laten_dim = dataset.user_num, dataset.item_num, args.laten_dim
    if torch.cuda.is_available() and args.cuda < torch.cuda.device_count():
        torch.cuda.set_device(args.cuda)
    args.l2_i = args.l2_u
    print(args)
    torch.manual_seed(2000)
    torch.cuda.manual_seed(2001)
    np.random.seed(2002)
    model = SPMF(args, dataset, user_num, item_num, laten_dim)
    if args.pre_model:
        model.MFbase.load_state_dict(torch.load(args.pre_model, map_location=model.device))
    if args.method == "spmf":
        model.base_train_not_train(args.start_idx - 1)          # (NameError in the reference, :667)
    model.run(args.start_idx, method=args.method)
    return model


if __name__ == "__main__":
    main()
