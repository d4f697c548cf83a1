"""Row-sharded SML tables across the GPUs of one box (BASELINE.json north_star item 4).

The reference is single-GPU (``main_yelp.py:125``); this is the scale-out of the same arithmetic for
tables that do not fit one device (config 4: 200 M users x 20 M items, d = 64):

  * every table copy (live MF rows = w_hat, w_{t-1}, Adam m / v, dense gradient) is sharded by row id,
    ``owner = id % world``, ``local row = id // world`` (round-robin spreads the Zipf-popular ids);
  * a step is data parallel over the batch: each rank takes B / world triples, asks the owners for the
    (w_{t-1}, w_hat) pairs of its ids (all-to-all of ids, all-to-all of 512 B row pairs), runs the fused
    forward/backward on the received rows (``sml_run_mf_grads`` with pitch-128 views, no de-interleave copy),
    and returns one 256 B row gradient per id (all-to-all); owners scatter-add (+ the per-occurrence l2 term)
    and run dense Adam on their shard.  Duplicated ids travel once per occurrence, exactly like the reference's
    ``embedding_dense_backward`` sums them;
  * the transfer step exchanges rows the same way and all-reduces the 1.58 MB theta gradient; every rank then
    applies the identical Adam update (theta is replicated);
  * ``updata`` (w_t = Transfer(w_{t-1}, w_hat) on every row) and Adam are row-local: no communication.

The exchange plan (``RowExchange``) only needs ``torch.distributed`` and two injected local operators, so the
host logic is testable on CPU with the gloo backend (tests/test_shard_gloo.py); on the GPU the operators are the
CUDA kernels ``sml_gather_pairs`` / ``sml_scatter_grads`` and the collectives run over NCCL / NVLink.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class CommTimers(object):
    """CUDA-event timers + byte counters around the collectives of the sharded path (bench.py's ``sharded`` object).
    ``with comm("rows", nbytes_to_peers): dist.all_to_all_single(...)`` records one event pair on the current stream
    (synchronous torch.distributed collectives make the current stream wait for NCCL's, so the pair brackets them)."""

    def __init__(self):
        self.events, self.bytes, self._tag = {}, {}, None

    def __call__(self, tag, nbytes=0):
        self._tag = tag
        self.bytes[tag] = self.bytes.get(tag, 0) + int(nbytes)
        return self

    def __enter__(self):
        self._e0 = torch.cuda.Event(enable_timing=True)
        self._e0.record()

    def __exit__(self, *exc):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.events.setdefault(self._tag, []).append((self._e0, e1))

    def summary(self):
        """-> {tag: dict(calls, ms, bytes_to_peers)} (synchronises)."""
        torch.cuda.synchronize()
        return {t: dict(calls=len(ev), ms=float(sum(a.elapsed_time(b) for a, b in ev)), bytes_to_peers=self.bytes.get(t, 0))
                for t, ev in self.events.items()}


class _NoTimer(object):
    def __call__(self, *a, **k):
        return self

    def __enter__(self):
        pass

    def __exit__(self, *exc):
        pass


class ExchangePlan(object):
    __slots__ = ("order", "inverse", "send_counts", "recv_counts", "recv_loc", "n")


class RowExchange(object):
    """ids -> rows -> gradients exchange for one sharded table."""

    def __init__(self, world, rank, group=None):
        self.world, self.rank, self.group = world, rank, group
        self.comm = _NoTimer()               # bench.py installs a CommTimers here

    def _a2a(self, send, send_counts, recv_counts, tag="ids"):
        out = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
        if self.world == 1:
            out.copy_(send)
            return out
        row_bytes = send.element_size() * (send[0].numel() if send.dim() > 1 and send.shape[0] else 1)
        to_peers = (int(sum(send_counts)) - int(send_counts[self.rank])) * row_bytes
        with self.comm(tag, to_peers):
            dist.all_to_all_single(out, send.contiguous(), output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                                   group=self.group)
        return out

    def plan(self, ids):
        """ids: int64 [n] global row ids needed by this rank (duplicates allowed)."""
        p = ExchangePlan()
        owner = ids % self.world
        p.order = torch.argsort(owner, stable=True)          # group the requests by owner
        p.inverse = torch.empty_like(p.order)
        p.inverse[p.order] = torch.arange(ids.numel(), device=ids.device)
        counts = torch.bincount(owner, minlength=self.world)
        if self.world > 1:
            rc = torch.empty_like(counts)
            dist.all_to_all_single(rc, counts, group=self.group)
        else:
            rc = counts
        p.send_counts = [int(x) for x in counts.tolist()]
        p.recv_counts = [int(x) for x in rc.tolist()]
        p.recv_loc = self._a2a((ids // self.world)[p.order], p.send_counts, p.recv_counts)   # local rows others ask of me
        p.n = ids.numel()
        return p

    def plan_epoch(self, ids, sizes):
        """Plans for a whole epoch at once.  ``ids``: int64 [sum(sizes)], the ids this rank needs, step after step
        (``sizes[s]`` of them at step s; the step count must be the same on every rank).  Returns one ExchangePlan per
        step, equal to what ``plan`` would return for that step's ids -- but with ONE host synchronisation and two
        all-to-alls per epoch instead of per step, so the steps themselves enqueue without ever waiting for the GPU."""
        W, S = self.world, len(sizes)
        dev = ids.device
        starts = [0]
        for n in sizes:
            starts.append(starts[-1] + int(n))
        if ids.numel() != starts[-1]:
            raise ValueError("plan_epoch: ids has %d entries, sizes sum to %d" % (ids.numel(), starts[-1]))
        step = torch.repeat_interleave(torch.arange(S, device=dev), torch.tensor(sizes, device=dev), output_size=starts[-1])
        owner = ids % W
        by_step = torch.argsort(step * W + owner, stable=True)            # per step: requests grouped by owner
        inv = torch.empty_like(by_step)
        inv[by_step] = torch.arange(ids.numel(), device=dev)
        key = owner * S + step
        by_owner = torch.argsort(key, stable=True)                        # per owner: steps in order (the send layout)
        counts = torch.bincount(key, minlength=W * S).view(W, S)          # [owner, step]
        if W > 1:
            rc = torch.empty_like(counts)
            dist.all_to_all_single(rc, counts.contiguous(), group=self.group)   # rc[r, s]: what rank r asks of me at step s
        else:
            rc = counts
        both = torch.stack([counts, rc]).tolist()                         # the one host sync of the epoch
        counts_h, rc_h = both
        recv_all = self._a2a((ids // W)[by_owner], [sum(r) for r in counts_h], [sum(r) for r in rc_h])   # grouped (source, step)
        n_recv = recv_all.numel()
        src_step = torch.repeat_interleave(torch.arange(W * S, device=dev), rc.reshape(-1), output_size=n_recv)
        regroup = torch.argsort((src_step % S) * W + src_step // S, stable=True)     # -> grouped (step, source)
        recv_by_step = recv_all[regroup]
        plans, r0 = [], 0
        for s in range(S):
            p = ExchangePlan()
            a, b = starts[s], starts[s + 1]
            p.order = by_step[a:b] - a
            p.inverse = inv[a:b] - a
            p.send_counts = [counts_h[r][s] for r in range(W)]
            p.recv_counts = [rc_h[r][s] for r in range(W)]
            nr = sum(p.recv_counts)
            p.recv_loc = recv_by_step[r0:r0 + nr]
            r0 += nr
            p.n = b - a
            plans.append(p)
        return plans

    def fetch(self, plan, gather_fn, permute=True):
        """gather_fn(local_rows) -> [m, W] rows of this rank's shard; returns [n, W] in the original id order, or
        (``permute=False``) as received, grouped by owner: row k of the request then sits at ``plan.inverse[k]`` -- consumers
        that gather through an id list take ``plan.inverse`` as that list and save the [n, W] index copy."""
        mine = gather_fn(plan.recv_loc)
        got = self._a2a(mine, plan.recv_counts, plan.send_counts, tag="rows")
        return got[plan.inverse] if permute else got

    def fetch_async(self, plan, gather_fn):
        """Start a fetch (un-permuted, see ``fetch``) without making the current stream wait for it: the gather runs on the
        current stream, the all-to-all on the communicator's own stream, overlapping whatever is enqueued next.
        -> (rows as received, handle for ``fetch_wait``).  Only for tables that nothing enqueued in between modifies."""
        mine = gather_fn(plan.recv_loc)
        out = mine.new_empty((int(sum(plan.send_counts)),) + tuple(mine.shape[1:]))
        if self.world == 1:
            out.copy_(mine)
            return out, None
        row_bytes = mine.element_size() * (mine[0].numel() if mine.shape[0] else 1)
        to_peers = (int(sum(plan.recv_counts)) - int(plan.recv_counts[self.rank])) * row_bytes
        work = dist.all_to_all_single(out, mine.contiguous(), output_split_sizes=list(plan.send_counts), input_split_sizes=list(plan.recv_counts),
                                      group=self.group, async_op=True)
        return out, (work, mine, to_peers)

    def fetch_wait(self, handle):
        """Make the current stream wait for a ``fetch_async``; the timer brackets only the part of it that was NOT hidden."""
        if handle is None:
            return
        work, _, to_peers = handle
        with self.comm("rows", to_peers):
            work.wait()

    def push(self, plan, rows, scatter_fn, ordered=False):
        """rows: [n, W] per-id payload (row gradients) in the original id order -- or (``ordered``) already in send order,
        grouped by owner; scatter_fn(local_rows, payload) is called once on the owner side with everything this rank received."""
        recv = self._a2a(rows if ordered else rows[plan.order], plan.send_counts, plan.recv_counts, tag="grads")
        scatter_fn(plan.recv_loc, recv)


def shard_rows(table, world, rank):
    """Rows of a replicated [N, d] tensor owned by ``rank`` (ids rank, rank + world, ...)."""
    return table[rank::world].contiguous()


class ShardedSML(object):
    """Row-sharded state + the sharded MF step, transfer step, updata and candidate evaluation.
    ``transfer`` is a (replicated) ``ConvTransfer_com``; all tensors live on this rank's GPU."""

    def __init__(self, user_tab, item_tab, transfer, world=1, rank=0, group=None, mf_lr=0.01, l2=1e-6, tr_lr=0.001, tr_l2=1e-4,
                 ops=None):
        """``ops``: the local operator module (default: the CUDA library wrappers, sml_b200.ops).  The CPU gloo test injects a
        numpy stand-in with the same functions to exercise the exchange / lazy-Adam logic without a GPU."""
        if ops is None:
            from . import ops
        self.ops = ops
        if getattr(transfer, "variant", ops.VARIANT_COM) != ops.VARIANT_COM:
            # ConvTransfer (variant CONV) trains with the BPR loss and an L2-normalised user output
            # (model/conv_transfer.py:57-85); the sharded steps below hard-wire ConvTransfer_com's BCE path
            raise NotImplementedError("ShardedSML supports ConvTransfer_com only (the reference default, --transfer_type conv_com)")
        self.world, self.rank, self.group = world, rank, group
        self.ex = RowExchange(world, rank, group)
        z = torch.zeros_like
        self.user, self.item = user_tab, item_tab                       # local shards of the live MF tables (w_hat params)
        self.last_user, self.last_item = user_tab.clone(), item_tab.clone()
        self.user_hat, self.item_hat = user_tab.clone(), item_tab.clone()
        self.m_user, self.v_user, self.g_user = z(user_tab), z(user_tab), z(user_tab)
        self.m_item, self.v_item, self.g_item = z(item_tab), z(item_tab), z(item_tab)
        self.transfer = transfer
        self.m_theta, self.v_theta = z(transfer.theta), z(transfer.theta)
        dev = user_tab.device
        # MF Adam: the reference's dense update in its bit-identical row-lazy form (ops.adam_rows) -- per step only the
        # rows the exchange touched move, the whole shard once per flush() -- instead of 7 table sweeps per step
        self.mf_state, self.tr_state = ops.new_adam_state(dev, history=True), ops.new_adam_state(dev)
        self.stamp_user = ops.new_row_stamps(user_tab.shape[0], self.mf_state)
        self.stamp_item = ops.new_row_stamps(item_tab.shape[0], self.mf_state)
        self._pending = 0
        self.mf_lr, self.l2, self.tr_lr, self.tr_l2 = mf_lr, l2, tr_lr, tr_l2
        self.loss = torch.zeros(2, dtype=torch.float32, device=dev)
        self._ws = {}

    # ------------------------------------------------------------------ helpers
    def _workspace(self, B):
        if B not in self._ws:
            self._ws[B] = torch.zeros(int(self.ops.lib().sml_step_workspace_bytes(B)), dtype=torch.uint8, device=self.user.device)
        return self._ws[B]

    def _global_batch(self, B):
        if self.world == 1:
            return B
        t = torch.tensor([B], dtype=torch.int64, device=self.user.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def _prefetch(self, plans, hat_u, hat_i):
        """Start the row exchange of a later step (snapshot tables only: nothing may write them before the step runs)."""
        ops = self.ops
        pu, pi = plans
        ru, hu = self.ex.fetch_async(pu, lambda loc: ops.gather_pairs(self.last_user, hat_u, loc))
        ri, hi = self.ex.fetch_async(pi, lambda loc: ops.gather_pairs(self.last_item, hat_i, loc))
        return ru, ri, hu, hi

    def _forward_backward(self, user, item, neg, hat_u, hat_i, want_theta_grad, live=False, plans=None, rows=None):
        """Exchange rows, run forward + loss + gradients on the received [last | hat] pairs.  ``live``: hat_* are the
        live MF shards, whose requested rows first catch up with the zero-gradient Adam steps they missed.  ``rows``: the
        result of an earlier ``_prefetch`` of this step's plans (the exchange then overlapped the previous step)."""
        ops = self.ops
        B = user.numel()
        pu, pi = plans if plans is not None else (self.ex.plan(user), self.ex.plan(torch.cat([item, neg])))
        if live:
            ops.adam_rows(self.user, self.m_user, self.v_user, None, self.stamp_user, pu.recv_loc, self.mf_state, apply=False)
            ops.adam_rows(self.item, self.m_item, self.v_item, None, self.stamp_item, pi.recv_loc, self.mf_state, apply=False)
        # the received [last | hat] pairs stay in arrival order (grouped by owner): the step kernels gather them through the
        # exchange plan's inverse permutation, like they gather table rows through batch ids (no [n, 128] index copy)
        if rows is not None:
            ru, ri, hu, hi = rows
            self.ex.fetch_wait(hu); self.ex.fetch_wait(hi)
        else:
            ru = self.ex.fetch(pu, lambda loc: ops.gather_pairs(self.last_user, hat_u, loc), permute=False)      # [B, 128]
            ri = self.ex.fetch(pi, lambda loc: ops.gather_pairs(self.last_item, hat_i, loc), permute=False)      # [2B, 128]
        iu, ii = pu.inverse.contiguous(), pi.inverse.contiguous()
        total, rp, rn = ops.step_rows(B)
        d_rows = torch.empty(total, 64, dtype=torch.float32, device=user.device)
        if want_theta_grad:
            self.transfer.theta_grad.zero_()
        a = ops.make_step_args(user=iu, item=ii[:B], neg=ii[B:], last_user=ru, last_item=ri, hat_user=ru[:, 64:], hat_item=ri[:, 64:],
                               theta=self.transfer.theta, variant=self.transfer.variant, loss=ops.LOSS_BCE,
                               g_theta=self.transfer.theta_grad if want_theta_grad else None, loss_out=self.loss,
                               workspace=self._workspace(B), table_pitch=128, n_users=B, n_items=2 * B, d_rows_by_id=True)
        ops.run_mf_grads(a, d_rows=d_rows)               # d_rows[:B] / d_rows[rp:rp + 2B] come out in send order
        return pu, pi, d_rows, rp

    # ------------------------------------------------------------------ the two hot loops
    def _epoch_plans(self, user, item, neg, B):
        """Per-step exchange plans and global-batch scales for an epoch of this rank's triples (one host sync)."""
        n = user.numel()
        sizes = [min(B, n - o) for o in range(0, n, B)]
        if self.world > 1:
            ns = torch.tensor([len(sizes), -len(sizes)], dtype=torch.int64, device=user.device)
            dist.all_reduce(ns, op=dist.ReduceOp.MAX, group=self.group)
            if int(ns[0]) != -int(ns[1]):
                raise ValueError("sharded epoch: ranks disagree on the step count (%d..%d); give every rank the same number "
                                 "of batches (pad or trim the per-rank triples)" % (-int(ns[1]), int(ns[0])))
            t = torch.tensor(sizes, dtype=torch.int64, device=user.device)
            dist.all_reduce(t, group=self.group)                   # global batch of every step
        pu = self.ex.plan_epoch(user, sizes)
        it = torch.cat([torch.cat([item[o:o + b], neg[o:o + b]]) for o, b in zip(range(0, n, B), sizes)]) if n else item
        pi = self.ex.plan_epoch(it, [2 * b for b in sizes])
        glob = t.tolist() if self.world > 1 else sizes
        return sizes, pu, pi, [b / g for b, g in zip(sizes, glob)]

    def mf_epoch(self, user, item, neg, B):
        """HOT LOOP A over this rank's triples, ``B`` per step: the exchange is planned once for the whole epoch, the
        steps enqueue without host synchronisation.  Returns the summed (global-mean) step losses as a device scalar."""
        sizes, pu, pi, scales = self._epoch_plans(user, item, neg, B)
        total = torch.zeros((), dtype=torch.float32, device=user.device)
        for s, b in enumerate(sizes):
            o = s * B
            total += self.mf_step(user[o:o + b], item[o:o + b], neg[o:o + b], plans=(pu[s], pi[s]), scale=scales[s])
        return total

    def tr_epoch(self, user, item, neg, B):
        """HOT LOOP B over this rank's triples (see mf_epoch).  The transfer step reads snapshot tables that no step writes, so
        the row exchange of step s + 1 is started before step s computes and runs under it (model/transfer.py:701-728 has no
        such dependency either: only theta changes between steps)."""
        sizes, pu, pi, scales = self._epoch_plans(user, item, neg, B)
        total = torch.zeros((), dtype=torch.float32, device=user.device)
        nxt = self._prefetch((pu[0], pi[0]), self.user_hat, self.item_hat) if sizes else None
        for s, b in enumerate(sizes):
            o = s * B
            cur, nxt = nxt, (self._prefetch((pu[s + 1], pi[s + 1]), self.user_hat, self.item_hat) if s + 1 < len(sizes) else None)
            total += self.tr_step(user[o:o + b], item[o:o + b], neg[o:o + b], plans=(pu[s], pi[s]), scale=scales[s], rows=cur)
        return total

    def mf_step(self, user, item, neg, plans=None, scale=None):
        """HOT LOOP A body (model/transfer.py:463-511) on this rank's slice of the global batch."""
        ops = self.ops
        B = user.numel()
        if scale is None:
            scale = B / self._global_batch(B)                  # BCE is a mean over the GLOBAL batch
        if self._pending >= ops.ADAM_HISTORY - 2:
            self.flush()                                       # the step-scalar history ring is about to wrap
        ops.adam_tick(self.mf_state, self.mf_lr)
        self._pending += 1
        pu, pi, d_rows, rp = self._forward_backward(user, item, neg, self.user, self.item, False, live=True, plans=plans)
        self.ex.push(pu, d_rows[:B], lambda loc, g: ops.scatter_grads(self.g_user, self.user, loc, g.contiguous(), scale, self.l2), ordered=True)
        self.ex.push(pi, d_rows[rp:rp + 2 * B], lambda loc, g: ops.scatter_grads(self.g_item, self.item, loc, g.contiguous(), scale, self.l2),
                     ordered=True)
        ops.adam_rows(self.user, self.m_user, self.v_user, self.g_user, self.stamp_user, pu.recv_loc, self.mf_state, apply=True)
        ops.adam_rows(self.item, self.m_item, self.v_item, self.g_item, self.stamp_item, pi.recv_loc, self.mf_state, apply=True)
        return self.loss[0] * scale

    def flush(self):
        """Bring every row of the live shards up to the current Adam step (before they are read as a whole)."""
        if self._pending:
            self.ops.adam_flush(self.user, self.m_user, self.v_user, self.stamp_user, self.mf_state)
            self.ops.adam_flush(self.item, self.m_item, self.v_item, self.stamp_item, self.mf_state)
            self._pending = 0

    def tr_step(self, user, item, neg, plans=None, scale=None, rows=None):
        """HOT LOOP B body (model/transfer.py:701-728): theta gradients, all-reduced, replicated Adam."""
        ops = self.ops
        B = user.numel()
        if scale is None:
            scale = B / self._global_batch(B)
        self._forward_backward(user, item, neg, self.user_hat, self.item_hat, True, plans=plans, rows=rows)
        g = self.transfer.theta_grad
        if scale != 1.0:
            g.mul_(scale)
        if self.world > 1:
            with self.ex.comm("theta", 2 * (self.world - 1) * g.numel() * 4 // self.world):    # ring all-reduce volume per rank
                dist.all_reduce(g, group=self.group)
        ops.adam_tick(self.tr_state, self.tr_lr)
        ops.adam_dense(self.transfer.theta, self.m_theta, self.v_theta, g, self.tr_state, weight_decay=self.tr_l2)
        return self.loss[0] * scale

    # ------------------------------------------------------------------ row-local pieces
    def save_last(self):
        self.flush()
        self.last_user.copy_(self.user); self.last_item.copy_(self.item)

    def save_hat(self):
        self.flush()
        self.user_hat.copy_(self.user); self.item_hat.copy_(self.item)

    def updata(self):
        """model/transfer.py:884-902 on the local shard: no communication."""
        ops = self.ops
        self.flush()            # nothing pending survives the overwrite below; stamps end at the current step
        th = self.transfer.theta
        ops.transfer_forward(self.last_user, self.user_hat, th[:ops.NET_STRIDE], variant=self.transfer.variant, out=self.user)
        ops.transfer_forward(self.last_item, self.item_hat, th[ops.NET_STRIDE:], variant=self.transfer.variant, out=self.item)

    def eval_candidates(self, rows, topK):
        """Candidate evaluation of this rank's slice of a test file: user rows come through the exchange, the
        (small) item table is all-gathered once; returns global (hits, ndcg_sum, n)."""
        ops = self.ops
        self.flush()
        dev = rows.device
        n = rows.shape[0]
        pu = self.ex.plan(rows[:, 0].contiguous())
        ur = self.ex.fetch(pu, lambda loc: self.user[loc])                   # [n, 64]; plain row gather (plumbing)
        if self.world > 1:
            n_items = torch.tensor([self.item.shape[0]], dtype=torch.int64, device=dev)
            parts = [torch.empty_like(n_items) for _ in range(self.world)]
            dist.all_gather(parts, n_items, group=self.group)
            mx = max(int(p.item()) for p in parts)
            pad = torch.zeros(mx, 64, dtype=torch.float32, device=dev); pad[:self.item.shape[0]] = self.item
            allp = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(allp, pad, group=self.group)
            full = torch.stack(allp, 1).reshape(-1, 64)                     # row id = local * world + rank
        else:
            full = self.item
        local_rows = rows.clone()
        local_rows[:, 0] = torch.arange(n, device=dev)
        gt, eq = ops.eval_candidates(ur.contiguous(), full.contiguous(), local_rows)
        hits, nd = ops.eval_reduce(gt, eq, topK, batch=max(n, 1))
        out = torch.stack([hits.sum().float(), nd.sum(), torch.tensor(float(n), device=dev)])
        if self.world > 1:
            dist.all_reduce(out, group=self.group)
        return out

    def eval_fullcat(self, pairs, topK, chunk=1 << 16):
        """Full-catalog recall / NDCG@K (BASELINE.json config 5) of this rank's slice of evaluated (user, positive item)
        pairs against the WHOLE row-sharded catalog: user rows and positive-item rows come through the exchange, the
        positive score is computed once (ops.fullcat_pos_scores: the score GEMM's own arithmetic), every rank ranks all pairs against its own item
        shard with the tcgen05 score GEMM (ops.fullcat_ranks) and the per-pair counts add up with one all-reduce.
        Returns global (hits, ndcg_sum, n) like eval_candidates."""
        ops = self.ops
        self.flush()
        dev = pairs.device
        W, R = self.world, self.rank
        n = pairs.shape[0]
        nmax = n
        if W > 1:
            t = torch.tensor([n], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            nmax = int(t.item())
        items_pk = ops.pack_rows(self.item)                      # this rank's item shard as a tensor-core operand
        out = torch.zeros(3, dtype=torch.float32, device=dev)
        for c0 in range(0, max(nmax, 1), chunk):
            sl = pairs[c0:c0 + chunk]
            m = sl.shape[0]
            mmax = min(chunk, nmax - c0)
            users, pos = sl[:, 0].contiguous(), sl[:, 1].contiguous()
            ur = self.ex.fetch(self.ex.plan(users), lambda loc: self.user[loc])
            pr = self.ex.fetch(self.ex.plan(pos), lambda loc: self.item[loc])
            ar = torch.arange(m, dtype=torch.int64, device=dev)
            sp = ops.fullcat_pos_scores(ur.contiguous(), pr.contiguous()) if m else torch.empty(0, device=dev)
            if W > 1:       # every rank ranks every pair of the chunk against its own item shard
                pad_u = torch.zeros(mmax, 64, device=dev); pad_u[:m] = ur
                pad_s = torch.full((mmax,), float("inf"), device=dev); pad_s[:m] = sp
                pad_p = torch.full((mmax,), -1, dtype=torch.int64, device=dev); pad_p[:m] = pos
                all_u = torch.empty(W * mmax, 64, device=dev); dist.all_gather_into_tensor(all_u, pad_u, group=self.group)
                all_s = torch.empty(W * mmax, device=dev); dist.all_gather_into_tensor(all_s, pad_s, group=self.group)
                all_p = torch.empty(W * mmax, dtype=torch.int64, device=dev); dist.all_gather_into_tensor(all_p, pad_p, group=self.group)
            else:
                all_u, all_s, all_p = ur.contiguous(), sp, pos
            if all_u.shape[0] == 0:
                continue
            local_pos = torch.where((all_p >= 0) & (all_p % W == R), all_p // W, torch.full_like(all_p, -1))
            ar_all = torch.arange(all_u.shape[0], dtype=torch.int64, device=dev)
            gt, eq = ops.fullcat_ranks(all_u, None, ar_all, local_pos, items_packed=items_pk, n_items=self.item.shape[0], s_pos=all_s)
            if W > 1:
                cnt = torch.stack([gt, eq])
                dist.all_reduce(cnt, group=self.group)
                gt, eq = cnt[0, R * mmax:R * mmax + m].contiguous(), cnt[1, R * mmax:R * mmax + m].contiguous()
            if m:
                hits, nd = ops.eval_reduce(gt, eq, topK, batch=m)
                out += torch.stack([hits.sum().float(), nd.sum(), torch.tensor(float(m), device=dev)])
        if W > 1:
            dist.all_reduce(out, group=self.group)
        return out
