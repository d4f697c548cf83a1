"""Period stream -- drop-in for the hot-path part of the reference's data/dataset2.py.

``transfer_data`` keeps the reference's constructor signature, attributes and the
``next_train(d_time) -> (set_t, set_tt, now_test, val)`` state machine
(data/dataset2.py:204-351); the Dataset classes keep their names and per-item semantics.
Differences that do not change results: the constructor does not load every train file just to
print statistics (data/dataset2.py:234-237), and ``np.load`` results are cached per file.
"""
from __future__ import annotations

import copy
import os

import numpy as np


class testDataset(object):
    """reference: data/dataset2.py:50-60 -- row passthrough."""

    def __init__(self, dataset):
        self.data = dataset

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, idx):
        return self.data[idx]


class trainDataset_withPreSample(object):
    """reference: data/dataset2.py:172-201.  Rows are [user, pos, neg_1 .. neg_k]; one pre-sampled
    column serves as the negative and the column advances after every full pass.  Note the
    reference's quirk: the shuffled column list starts at column 1, i.e. it includes the positive
    column, and ``neg_all = n_cols - 2`` (:181-184)."""

    def __init__(self, input_dataset):
        self.all_data = input_dataset
        self.have_read = 0
        self.neg_flag = np.arange(1, self.all_data.shape[1])
        np.random.shuffle(self.neg_flag)
        self.neg_all = input_dataset.shape[1] - 2
        self.used_neg_count = 0
        self.data_len = input_dataset.shape[0]

    def __len__(self):
        return self.all_data.shape[0]

    def current_column(self):
        return int(self.neg_flag[self.used_neg_count])

    def advance_epoch(self):
        """What data_len consecutive __getitem__ calls do to the column state (:195-200)."""
        self.have_read = 0
        self.used_neg_count += 1
        if self.used_neg_count >= self.neg_all:
            np.random.shuffle(self.neg_flag)
            self.used_neg_count = 0

    def __getitem__(self, idx):
        user = self.all_data[idx, 0]
        item = self.all_data[idx, 1]
        neg_item = self.all_data[idx, self.neg_flag[self.used_neg_count]]
        self.have_read += 1
        if self.have_read >= self.data_len:
            self.advance_epoch()
        return user, item, neg_item


class transfer_data(object):
    """reference: data/dataset2.py:203-351."""

    def __init__(self, args, path="dataset/", datasetname="News", online_train_time=21, file_path_list=None,
                 test_list=None, validation_list=None, online_test_time=48):
        self.TR_sample_type = args.TR_sample_type
        self.TR_stop_ = args.TR_stop_
        self.MF_sample = args.MF_sample
        self.current_as_set_tt = args.set_t_as_tt
        self.path = path
        self.dataname = datasetname
        self.file_list = file_path_list
        self.test_list = test_list
        self.val_list = validation_list
        self.len = len(file_path_list)
        self.online_trian_time = online_train_time
        self.online_test_time = online_test_time
        self.start_test_time = online_test_time
        self.test_count = 0
        information = np.load(self.path + self.dataname + "/" + "information.npy")
        self.user_number = information[1]
        self.inter_all = information[0]
        self.item_number = information[2]
        print(information)
        self._cache = {}
        self.cache_files = 6                       # np.load results kept (LRU): periods t and t+1, train + test

    def reinit(self):
        self.test_count = 0
        self.start_test_time = copy.deepcopy(self.online_test_time)

    def _load(self, kind, name):
        # only periods t and t+1 are ever re-read within a stage (set_t / set_tt / val / now_test): keep the last few
        # files (LRU) instead of the whole stream (a Yelp-shaped test file is ~600 MB of int64)
        key = (kind, name)
        hit = self._cache.pop(key, None)
        if hit is None:
            hit = np.load(self.path + self.dataname + "/" + kind + "/" + name + ".npy")
        self._cache[key] = hit
        while len(self._cache) > self.cache_files:
            self._cache.pop(next(iter(self._cache)))
        return hit

    def peek_files(self, d_time):
        """Files of stage ``d_time`` that are ALREADY in the host cache (no disk I/O here: reading a 600 MB file would stall the
        host while it should be enqueueing kernels); meta_train uploads them ahead of time."""
        now_time = self.online_trian_time + d_time
        if (now_time + 1) >= self.len:
            return []
        names = [("test", self.file_list[now_time]), ("test", self.file_list[now_time + 1])]
        return [self._cache[k] for k in names if k in self._cache]

    def _set_t(self, now_time):
        if self.MF_sample == "alone":
            return self._load("train", self.file_list[now_time])
        elif self.MF_sample == "all":
            return self._load("test", self.file_list[now_time])
        raise TypeError("now such type when read next train sets")

    def _set_tt(self, now_time):
        t = now_time if self.current_as_set_tt else now_time + 1
        if self.TR_sample_type == "alone":
            return self._load("train", self.file_list[t])
        elif self.TR_sample_type == "all":
            return self._load("test", self.file_list[t])
        raise TypeError("no such TR sample type")

    def next_train(self, d_time):
        """-> (set_t = D_t, set_tt = D_{t+1} or None, now_test or None, val); all None at the end of
        the stream.  Same three branches as the reference (:257-351)."""
        now_time = self.online_trian_time + d_time
        if (now_time + 1) >= self.len:
            return None, None, None, None
        print("now time:", now_time)
        print("will be test data:", now_time + 1)
        if (now_time + 1) < self.start_test_time:
            set_t = self._set_t(now_time)
            val = self._load("test", self.file_list[now_time + 1])
            set_tt = self._set_tt(now_time)
            return set_t, set_tt, None, val
        elif self.TR_stop_:
            set_t = self._set_t(now_time)
            now_test = self._load("test", self.test_list[self.test_count])
            val = now_test
            self.test_count += 1
            return set_t, None, now_test, val
        else:
            set_t = self._set_t(now_time)
            val = self._load("test", self.file_list[now_time + 1])
            set_tt = self._set_tt(now_time)
            now_test = self._load("test", self.test_list[self.test_count])
            print("real test:", self.test_list[self.test_count])
            self.test_count += 1
            return set_t, set_tt, now_test, val


def select_neg_forinteraction(path="dataset/", datasetname="News", file_path_list=None, leave_for_init_train=0.7, neg_num=999,
                              extra_draws=1000):
    """Test-file builder (reference: data/dataset2.py:356-414): for every interaction of the periods after the first
    ``round(len * leave_for_init_train)``, ``neg_num`` distinct items seen so far that the user has not interacted
    with (up to and including that row), in random order; writes ``<path>/<name>/test/<i>.npy`` = [user, item, negs...].

    Draws from the global numpy RNG exactly like the reference (one ``np.random.choice`` of neg_num + 1000 per attempt,
    then one ``np.random.shuffle``), and indexes the same ``list(set)`` ordering of the items seen so far, so the files
    are bit-identical for the same ``np.random.seed``; the reference rebuilds that item array (O(items)) for every row,
    here it is rebuilt only when a new item appears.  Returns the list of test arrays."""
    base = os.path.join(path, datasetname)
    inter_all = [np.load(os.path.join(base, f + ".npy")) for f in file_path_list]
    n_files = len(inter_all)
    start = round(n_files * leave_for_init_train)
    seen = np.concatenate(inter_all[0:start], axis=0)
    user_item = {}
    for u, it in zip(seen[:, 0], seen[:, 1]):
        user_item.setdefault(u, set()).add(it)
    items_so_far = set(np.unique(seen[:, 1]))
    pool = np.array(list(items_so_far))
    out = []
    os.makedirs(os.path.join(base, "test"), exist_ok=True)
    for p in range(start, n_files):
        inter = inter_all[p]
        negs = np.empty((inter.shape[0], neg_num), dtype=pool.dtype)
        for r, (u, it) in enumerate(zip(inter[:, 0], inter[:, 1])):
            if it not in items_so_far:
                items_so_far.add(it)
                pool = np.array(list(items_so_far))          # iteration order of the set = the reference's all_item
            hist = user_item.setdefault(u, set())
            hist.add(it)
            have = np.fromiter(hist, dtype=pool.dtype, count=len(hist))
            while True:
                cand = np.setdiff1d(np.random.choice(pool, neg_num + extra_draws), have)   # sorted, unique
                if cand.shape[0] >= neg_num:
                    break
            np.random.shuffle(cand)
            negs[r] = cand[:neg_num]
        test = np.concatenate([inter, negs], axis=1)
        np.save(os.path.join(base, "test", str(p) + ".npy"), test)
        out.append(test)
    return out
