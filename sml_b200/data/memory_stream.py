"""In-memory period stream with the ``transfer_data`` interface (``next_train`` / ``reinit`` /
``user_number`` / ``item_number``), for benchmarks and tests that synthesise their periods instead
of reading ``.npy`` files.  Same branch logic as data/dataset2.py:257-351 of the reference with
``MF_sample="all"``, ``TR_sample_type="alone"`` (the defaults, main_yelp.py:47,75)."""
from __future__ import annotations


class MemoryStream(object):
    def __init__(self, periods, n_users, n_items, online_train_time=0, online_test_time=None, tr_stop=False):
        """periods: list of (train [N,2], test [N,2+n_neg]) int64 arrays, one per period."""
        self.periods = periods
        self.user_number = n_users
        self.item_number = n_items
        self.len = len(periods)
        self.online_trian_time = online_train_time
        self.online_test_time = self.len if online_test_time is None else online_test_time
        self.start_test_time = self.online_test_time
        self.TR_stop_ = tr_stop
        self.test_count = 0

    def reinit(self):
        self.test_count = 0
        self.start_test_time = self.online_test_time

    def next_train(self, d_time):
        now = self.online_trian_time + d_time
        if now + 1 >= self.len:
            return None, None, None, None
        set_t = self.periods[now][1]              # MF_sample="all": D_t is the test-format file
        val = self.periods[now + 1][1]
        if now + 1 < self.start_test_time:
            return set_t, self.periods[now + 1][0], None, val
        now_test = self.periods[self.online_test_time + self.test_count][1]
        self.test_count += 1
        if self.TR_stop_:
            return set_t, None, now_test, now_test
        return set_t, self.periods[now + 1][0], now_test, val

    def peek_files(self, d_time):
        """The test-format files stage ``d_time`` will read (set_t, val), without side effects: what meta_train uploads ahead of time."""
        now = self.online_trian_time + d_time
        if now + 1 >= self.len:
            return []
        return [self.periods[now][1], self.periods[now + 1][1]]
