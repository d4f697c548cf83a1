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


class ExchangePlan(object):
    __slots__ = ("order", "inverse", "send_counts", "recv_counts", "recv_loc", "n")


class RowExchange(object):
    """ids -> rows -> gradients exchange for one sharded table."""

    def __init__(self, world, rank, group=None):
        self.world, self.rank, self.group = world, rank, group

    def _a2a(self, send, send_counts, recv_counts):
        out = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
        if self.world == 1:
            out.copy_(send)
            return out
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

    def fetch(self, plan, gather_fn):
        """gather_fn(local_rows) -> [m, W] rows of this rank's shard; returns [n, W] in the original id order."""
        mine = gather_fn(plan.recv_loc)
        got = self._a2a(mine, plan.recv_counts, plan.send_counts)
        return got[plan.inverse]

    def push(self, plan, rows, scatter_fn):
        """rows: [n, W] per-id payload (row gradients) in the original id order; scatter_fn(local_rows, payload)
        is called once on the owner side with everything this rank received."""
        recv = self._a2a(rows[plan.order], plan.send_counts, plan.recv_counts)
        scatter_fn(plan.recv_loc, recv)


def shard_rows(table, world, rank):
    """Rows of a replicated [N, d] tensor owned by ``rank`` (ids rank, rank + world, ...)."""
    return table[rank::world].contiguous()


class ShardedSML(object):
    """Row-sharded state + the sharded MF step, transfer step, updata and candidate evaluation.
    ``transfer`` is a (replicated) ``ConvTransfer_com``; all tensors live on this rank's GPU."""

    def __init__(self, user_tab, item_tab, transfer, world=1, rank=0, group=None, mf_lr=0.01, l2=1e-6, tr_lr=0.001, tr_l2=1e-4):
        from . import ops
        self.ops = ops
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
        self.mf_state, self.tr_state = ops.new_adam_state(dev), ops.new_adam_state(dev)
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

    def _forward_backward(self, user, item, neg, hat_u, hat_i, want_theta_grad):
        """Exchange rows, run forward + loss + gradients on the received [last | hat] pairs."""
        ops = self.ops
        B = user.numel()
        pu = self.ex.plan(user)
        pi = self.ex.plan(torch.cat([item, neg]))
        ru = self.ex.fetch(pu, lambda loc: ops.gather_pairs(self.last_user, hat_u, loc))      # [B, 128]
        ri = self.ex.fetch(pi, lambda loc: ops.gather_pairs(self.last_item, hat_i, loc))      # [2B, 128]
        ar = torch.arange(2 * B, dtype=torch.int64, device=user.device)
        total, rp, rn = ops.step_rows(B)
        d_rows = torch.empty(total, 64, dtype=torch.float32, device=user.device)
        if want_theta_grad:
            self.transfer.theta_grad.zero_()
        a = ops.make_step_args(user=ar[:B], item=ar[:B], neg=ar[B:], last_user=ru, last_item=ri, hat_user=ru[:, 64:], hat_item=ri[:, 64:],
                               theta=self.transfer.theta, variant=self.transfer.variant, loss=ops.LOSS_BCE,
                               g_theta=self.transfer.theta_grad if want_theta_grad else None, loss_out=self.loss,
                               workspace=self._workspace(B), table_pitch=128, n_users=B, n_items=2 * B)
        ops.run_mf_grads(a, d_rows=d_rows)
        return pu, pi, d_rows, rp

    # ------------------------------------------------------------------ the two hot loops
    def mf_step(self, user, item, neg):
        """HOT LOOP A body (model/transfer.py:463-511) on this rank's slice of the global batch."""
        ops = self.ops
        B = user.numel()
        scale = B / self._global_batch(B)                      # BCE is a mean over the GLOBAL batch
        pu, pi, d_rows, rp = self._forward_backward(user, item, neg, self.user, self.item, False)
        self.ex.push(pu, d_rows[:B], lambda loc, g: ops.scatter_grads(self.g_user, self.user, loc, g.contiguous(), scale, self.l2))
        self.ex.push(pi, d_rows[rp:rp + 2 * B], lambda loc, g: ops.scatter_grads(self.g_item, self.item, loc, g.contiguous(), scale, self.l2))
        ops.adam_tick(self.mf_state, self.mf_lr)
        ops.adam_dense(self.user, self.m_user, self.v_user, self.g_user, self.mf_state)
        ops.adam_dense(self.item, self.m_item, self.v_item, self.g_item, self.mf_state)
        return self.loss[0] * scale

    def tr_step(self, user, item, neg):
        """HOT LOOP B body (model/transfer.py:701-728): theta gradients, all-reduced, replicated Adam."""
        ops = self.ops
        B = user.numel()
        scale = B / self._global_batch(B)
        self._forward_backward(user, item, neg, self.user_hat, self.item_hat, True)
        g = self.transfer.theta_grad
        if scale != 1.0:
            g.mul_(scale)
        if self.world > 1:
            dist.all_reduce(g, group=self.group)
        ops.adam_tick(self.tr_state, self.tr_lr)
        ops.adam_dense(self.transfer.theta, self.m_theta, self.v_theta, g, self.tr_state, weight_decay=self.tr_l2)
        return self.loss[0] * scale

    # ------------------------------------------------------------------ row-local pieces
    def save_last(self):
        self.last_user.copy_(self.user); self.last_item.copy_(self.item)

    def save_hat(self):
        self.user_hat.copy_(self.user); self.item_hat.copy_(self.item)

    def updata(self):
        """model/transfer.py:884-902 on the local shard: no communication."""
        ops = self.ops
        th = self.transfer.theta
        ops.transfer_forward(self.last_user, self.user_hat, th[:ops.NET_STRIDE], variant=self.transfer.variant, out=self.user)
        ops.transfer_forward(self.last_item, self.item_hat, th[ops.NET_STRIDE:], variant=self.transfer.variant, out=self.item)

    def eval_candidates(self, rows, topK):
        """Candidate evaluation of this rank's slice of a test file: user rows come through the exchange, the
        (small) item table is all-gathered once; returns global (hits, ndcg_sum, n)."""
        ops = self.ops
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
