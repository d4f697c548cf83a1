"""Synthetic period streams in the reference's on-disk layout.

Layout (README.md:22-25, data/dataset2.py:229-232,411-414 of the reference):
  <path>/<name>/information.npy        int64 [3] = [n_interactions, n_users, n_items]
  <path>/<name>/train/<p>.npy          int64 [N_p, 2]        (user, item)
  <path>/<name>/test/<p>.npy           int64 [N_p, 2 + n_neg] (user, pos item, n_neg negatives)
The negatives of a row are distinct items that are not in the user's history up to
and including that period (mirrors select_neg_forinteraction, data/dataset2.py:356-414).

There is no network in the build or bench environment, so every dataset used by
tests and bench.py comes from here ("data": "synthetic").
"""
from __future__ import annotations

import os

import numpy as np

YELP_SHAPE = dict(n_users=59082, n_items=122816, rows_per_period=75000, n_periods=40)
ADRESSA_SHAPE = dict(n_users=478612, n_items=20875, rows_per_period=58000, n_periods=63)


def _zipf_ids(rng, n, size, a=1.0):
    """Bounded Zipf(a) over [0, n) through the inverse CDF, then a fixed random
    relabelling so that popular ids are spread over the table."""
    ranks = np.arange(1, n + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-a))
    cdf /= cdf[-1]
    r = np.searchsorted(cdf, rng.random(size), side="left")
    return r.astype(np.int64)


def _distinct_negatives(rng, n_rows, n_neg, pool):
    """n_neg distinct members of ``pool`` per row: (a + s*k) mod P with s coprime to P
    enumerates distinct residues.  Marginally uniform, rows independent."""
    P = len(pool)
    if n_neg > P:
        raise ValueError("pool smaller than the number of negatives")
    a = rng.integers(0, P, size=(n_rows, 1))
    s = rng.integers(1, P, size=(n_rows, 1))
    g = np.gcd(s, P)
    while (g != 1).any():
        bad = g != 1
        s[bad] = rng.integers(1, P, size=int(bad.sum()))
        g = np.gcd(s, P)
    k = np.arange(n_neg, dtype=np.int64)[None, :]
    return pool[(a + s * k) % P]


def make_period(rng, n_users, n_items, n_rows, n_neg, item_pool=None, history=None, zipf_a=1.0,
                user_perm=None, item_perm=None):
    """One period: returns (train [N,2], test [N,2+n_neg]).  ``history`` is a dict
    user -> set(items) that is updated in place."""
    pool = np.arange(n_items, dtype=np.int64) if item_pool is None else np.asarray(item_pool, dtype=np.int64)
    u = _zipf_ids(rng, n_users, n_rows, zipf_a)
    it = pool[_zipf_ids(rng, len(pool), n_rows, zipf_a)]
    if user_perm is not None:
        u = user_perm[u]
    if item_perm is not None and item_pool is None:
        it = item_perm[it]
    train = np.stack([u, it], axis=1)
    neg = _distinct_negatives(rng, n_rows, n_neg, pool)
    if history is not None:
        for a, b in train:
            history.setdefault(int(a), set()).add(int(b))
        # fix the (rare) negatives that collide with the user's history
        for r in range(n_rows):
            h = history[int(u[r])]
            row = neg[r]
            bad = [c for c in range(n_neg) if int(row[c]) in h]
            if bad:
                taken = set(row.tolist()) | h
                for c in bad:
                    x = int(pool[rng.integers(0, len(pool))])
                    while x in taken:
                        x = int(pool[rng.integers(0, len(pool))])
                    row[c] = x
                    taken.add(x)
    else:
        # only guard against the positive itself
        hit = neg == it[:, None]
        if hit.any():
            rr, cc = np.nonzero(hit)
            for r, c in zip(rr, cc):
                taken = set(neg[r].tolist()) | {int(it[r])}
                x = int(pool[rng.integers(0, len(pool))])
                while x in taken:
                    x = int(pool[rng.integers(0, len(pool))])
                neg[r, c] = x
    test = np.concatenate([train, neg], axis=1)
    return train, test


def make_stream(n_users, n_items, rows_per_period, n_periods, n_neg=999, seed=0, churn=0.0,
                track_history=True, zipf_a=1.0):
    """List of (train, test) per period.  ``churn`` > 0 gives an Adressa-like stream:
    each period draws that fraction of its interactions from items first seen in
    that period (short-lived items)."""
    rng = np.random.default_rng(seed)
    user_perm = rng.permutation(n_users).astype(np.int64)
    item_perm = rng.permutation(n_items).astype(np.int64)
    history = {} if track_history else None
    out = []
    if churn <= 0.0:
        for _ in range(n_periods):
            out.append(make_period(rng, n_users, n_items, rows_per_period, n_neg, None, history, zipf_a,
                                   user_perm, item_perm))
        return out
    fresh_per_period = max(n_neg + 2, n_items // (n_periods + 2))
    seen = item_perm[:2 * fresh_per_period]
    nxt = 2 * fresh_per_period
    for _ in range(n_periods):
        fresh = item_perm[nxt:nxt + fresh_per_period]
        nxt = min(n_items, nxt + fresh_per_period)
        n_new = int(rows_per_period * churn) if len(fresh) else 0
        alive = np.concatenate([seen[-2 * fresh_per_period:], fresh])
        tr_a, te_a = make_period(rng, n_users, n_items, rows_per_period - n_new, n_neg, alive, history, zipf_a,
                                 user_perm, None)
        if n_new:
            pool_new = np.concatenate([fresh, alive[:max(0, n_neg + 2 - len(fresh))]])
            tr_b, te_b = make_period(rng, n_users, n_items, n_new, n_neg, alive, history, zipf_a, user_perm, None)
            tr_b[:, 1] = fresh[rng.integers(0, len(fresh), size=n_new)]
            te_b[:, 1] = tr_b[:, 1]
            del pool_new
            train = np.concatenate([tr_a, tr_b]); test = np.concatenate([te_a, te_b])
            p = rng.permutation(len(train))
            train, test = train[p], test[p]
        else:
            train, test = tr_a, te_a
        seen = np.concatenate([seen, fresh])
        out.append((train, test))
    return out


def write_stream(path, name, periods, n_users, n_items, file_names=None):
    """Write ``periods`` (list of (train, test)) in the reference's layout."""
    root = os.path.join(path, name)
    os.makedirs(os.path.join(root, "train"), exist_ok=True)
    os.makedirs(os.path.join(root, "test"), exist_ok=True)
    total = 0
    for p, (train, test) in enumerate(periods):
        fn = str(p) if file_names is None else file_names[p]
        np.save(os.path.join(root, "train", fn + ".npy"), train.astype(np.int64))
        np.save(os.path.join(root, "test", fn + ".npy"), test.astype(np.int64))
        total += len(train)
    np.save(os.path.join(root, "information.npy"), np.array([total, n_users, n_items], dtype=np.int64))
    return root
