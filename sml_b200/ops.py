"""Tensor-level wrappers over the C ABI (one function per entry point of include/sml_b200.h).

All tensors are CUDA tensors owned by the caller; nothing here computes on the CPU and
nothing falls back to PyTorch operators: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import (D, NET_STRIDE, VARIANT_COM, VARIANT_CONV, LOSS_BCE, LOSS_BPR, OPT_ADAM_DENSE_EXACT, OPT_ADAM_SPARSE, StepArgs,
                   check, lib, ptr, stream)

EVAL_BATCH = 1024      # evaluation2.test_model's DataLoader batch (model/transfer.py:431-435 of the reference)


def _f32(t, name):
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32" % name)
    return t


def _i64(t, name):
    if t.dtype != torch.int64:
        raise TypeError("%s must be int64" % name)
    return t


# ------------------------------------------------------------------ evaluation
EVAL_PREFILTER = os.environ.get("SML_EVAL_PREFILTER", "0") == "1"


def eval_candidates(user_tab, item_tab, rows, prefilter=None):
    """rows: int64 [n, 1+C] (user, positive, negatives...) -> (gt int32[n], eq int32[n]).
    ``prefilter`` (default off, SML_EVAL_PREFILTER=1 turns it on): bf16 pre-filter with exact fp32 fallback -- identical
    counts from about half the L2 bytes (sml_eval_candidates_prefilter), but measured no faster: it trades the L1/L2 limit of
    the fp32 kernel for an instruction-issue limit (profiles/r01_kernels_ncu.md)."""
    _f32(user_tab, "user_tab"); _f32(item_tab, "item_tab"); _i64(rows, "rows")
    n, w = rows.shape
    gt = torch.empty(n, dtype=torch.int32, device=rows.device)
    eq = torch.empty(n, dtype=torch.int32, device=rows.device)
    if EVAL_PREFILTER if prefilter is None else prefilter:
        scratch = _workspace(lib().sml_eval_prefilter_bytes(item_tab.shape[0]), rows.device, "eval_bf16")
        check(lib().sml_eval_candidates_prefilter(ptr(user_tab), ptr(item_tab), item_tab.shape[0], user_tab.shape[1], ptr(scratch),
                                                  ptr(rows), n, rows.stride(0), w - 1, ptr(gt), ptr(eq), stream()),
              "eval_candidates_prefilter")
        return gt, eq
    check(lib().sml_eval_candidates(ptr(user_tab), ptr(item_tab), user_tab.shape[1], ptr(rows), n, rows.stride(0), w - 1,
                                    ptr(gt), ptr(eq), stream()), "eval_candidates")
    return gt, eq


def eval_reduce(gt, eq, topk, batch=EVAL_BATCH, tie_loses=True):
    """-> (hits int32[nb], ndcg float32[nb]) per batch of `batch` consecutive rows."""
    n = gt.numel()
    nb = (n + batch - 1) // batch
    hits = torch.zeros(nb, dtype=torch.int32, device=gt.device)
    ndcg = torch.zeros(nb, dtype=torch.float32, device=gt.device)
    check(lib().sml_eval_reduce(ptr(gt), ptr(eq), n, batch, topk, int(bool(tie_loses)), ptr(hits), ptr(ndcg), stream()),
          "eval_reduce")
    return hits, ndcg


def pair_scores(user_tab, item_tab, user, item, norm=False):
    n = user.numel()
    out = torch.empty(n, dtype=torch.float32, device=user_tab.device)
    check(lib().sml_pair_scores(ptr(user_tab), ptr(item_tab), user_tab.shape[1], ptr(_i64(user, "user")), ptr(_i64(item, "item")),
                                n, int(bool(norm)), ptr(out), stream()), "pair_scores")
    return out


# ------------------------------------------------------------------ transfer forward
_ws_cache = {}


def _workspace(nbytes, device, tag):
    """Zero-initialised scratch, grown on demand and reused (keyed by device + tag)."""
    key = (device.index, tag)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def transfer_forward(x_t, x_hat, theta_net, variant=VARIANT_COM, ids=None, normalize_out=False, out=None, tensor_cores=None):
    """out[n] = one_transfer(stack(x_t[i_n], x_hat[i_n])), i_n = ids[n] or n.  theta_net: flat fp32 [NET_STRIDE].
    (``tensor_cores`` is ignored: the GEMM path is a process-wide choice, SML_GEMM=simt in the environment.)"""
    _f32(x_t, "x_t"); _f32(x_hat, "x_hat"); _f32(theta_net, "theta_net")
    n = x_t.shape[0] if ids is None else ids.numel()
    if out is None:
        out = torch.empty(n, D, dtype=torch.float32, device=x_t.device)
    l = lib()
    ws = _workspace(l.sml_transfer_fwd_workspace_bytes(n), x_t.device, "fwd")
    fn = l.sml_transfer_fwd
    check(fn(ptr(x_t), ptr(x_hat), ptr(ids), n, x_t.shape[1], variant, ptr(theta_net), int(bool(normalize_out)), ptr(out),
             ptr(ws), ws.numel(), stream()), "transfer_fwd")
    return out


# ------------------------------------------------------------------ optimizer
ADAM_HISTORY = 4096


def new_adam_state(device, history=False):
    """[step, (step_size, sqrt(bc2)) packed as floats, history flag, spare] (+ ADAM_HISTORY slots holding the float pair
    of every recent step when ``history``: what the row-lazy update replays) -- see sml_adam_tick."""
    st = torch.zeros(4 + (ADAM_HISTORY if history else 0), dtype=torch.int64, device=device)
    if history:
        st[2] = 1
    return st


STAMP_IDLE = 0x7fffffff


def new_row_stamps(n_rows, state, idle=True):
    """int32 [n_rows] 'last Adam step applied' stamps for a table whose exp_avg / exp_avg_sq are all zero (``idle``:
    SML_STAMP_IDLE, rows no gradient has reached yet are skipped by flushes on the stamp alone); ``idle=False``: the state's
    current step, for moments that are already non-zero (no sync)."""
    if idle:
        return torch.full((n_rows,), STAMP_IDLE, dtype=torch.int32, device=state.device)
    return state[0].to(torch.int32).expand(n_rows).contiguous()


def adam_tick(state, lr, beta1=0.9, beta2=0.999):
    check(lib().sml_adam_tick(ptr(state), lr, beta1, beta2, stream()), "adam_tick")


def adam_dense(p, m, v, g, state, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, zero_grad=True):
    check(lib().sml_adam_dense(ptr(p), ptr(m), ptr(v), ptr(g), p.numel(), ptr(state), beta1, beta2, eps, weight_decay,
                               int(bool(zero_grad)), stream()), "adam_dense")


def _i32t(t, name):
    if t.dtype != torch.int32 or not t.is_contiguous():
        raise TypeError("%s must be a contiguous int32 tensor" % name)
    return t


def adam_rows(p, m, v, g, stamp, ids, state, apply, beta1=0.9, beta2=0.999, eps=1e-8):
    """Row-lazy exact Adam on the listed rows: ``apply=False`` replays the zero-gradient steps they missed up to t-1,
    ``apply=True`` also applies step t with the gradient rows of ``g`` (re-zeroed)."""
    check(lib().sml_adam_rows(ptr(p), ptr(m), ptr(v), ptr(g), ptr(_i32t(stamp, "stamp")), ptr(_i64(ids, "ids")), ids.numel(), p.shape[0],
                              ptr(state), int(bool(apply)), beta1, beta2, eps, stream()), "adam_rows")


def adam_flush(p, m, v, stamp, state, beta1=0.9, beta2=0.999, eps=1e-8):
    """Every row of the table up to the state's current step (row-lazy exact Adam)."""
    check(lib().sml_adam_flush(ptr(p), ptr(m), ptr(v), ptr(_i32t(stamp, "stamp")), p.shape[0], ptr(state), beta1, beta2, eps, stream()),
          "adam_flush")


# ------------------------------------------------------------------ steps
def step_workspace(batch, device, tag="step"):
    return _workspace(lib().sml_step_workspace_bytes(batch), device, tag)


def step_rows(batch):
    """-> (total padded rows, first positive-item row, first negative-item row) of the per-step matrices."""
    rp, rn = C.c_int64(), C.c_int64()
    total = lib().sml_step_rows(batch, C.byref(rp), C.byref(rn))
    return int(total), int(rp.value), int(rn.value)


def make_step_args(*, user, item, neg, last_user, last_item, hat_user, hat_item, theta, variant=VARIANT_COM, loss=LOSS_BCE,
                   g_user=None, g_item=None, m_user=None, v_user=None, m_item=None, v_item=None, adam_state=None,
                   lr=0.0, l2=0.0, g_theta=None, m_theta=None, v_theta=None, loss_out=None, workspace=None, batch=None,
                   table_pitch=0, n_users=None, n_items=None, stamp_user=None, stamp_item=None, adaptive_beta=0.0, clip_max_norm=0.0,
                   d_rows_by_id=False):
    a = StepArgs()
    B = user.numel() if batch is None else int(batch)
    a.user, a.item, a.neg, a.batch = ptr(_i64(user, "user")), ptr(_i64(item, "item")), ptr(_i64(neg, "neg")), B
    _p = (lambda t: ptr(t)) if table_pitch in (0, D) else (lambda t: t.data_ptr())     # pitched views are not contiguous
    a.last_user, a.last_item = _p(_f32(last_user, "last_user")), _p(_f32(last_item, "last_item"))
    a.hat_user, a.hat_item = _p(_f32(hat_user, "hat_user")), _p(_f32(hat_item, "hat_item"))
    a.n_users = hat_user.shape[0] if n_users is None else n_users
    a.n_items = hat_item.shape[0] if n_items is None else n_items
    a.table_pitch = table_pitch
    a.theta, a.variant, a.loss = ptr(_f32(theta, "theta")), variant, loss
    a.g_user, a.g_item = ptr(g_user), ptr(g_item)
    a.m_user, a.v_user, a.m_item, a.v_item = ptr(m_user), ptr(v_user), ptr(m_item), ptr(v_item)
    a.adam_state, a.lr, a.l2 = ptr(adam_state), float(lr), float(l2)
    a.g_theta, a.m_theta, a.v_theta = ptr(g_theta), ptr(m_theta), ptr(v_theta)
    a.stamp_user = ptr(stamp_user if stamp_user is None else _i32t(stamp_user, "stamp_user"))
    a.stamp_item = ptr(stamp_item if stamp_item is None else _i32t(stamp_item, "stamp_item"))
    a.adaptive_beta, a.clip_max_norm = float(adaptive_beta), float(clip_max_norm)
    a.d_rows_by_id = int(bool(d_rows_by_id))
    if workspace is None:
        workspace = step_workspace(B, user.device)
    a.loss_out, a.workspace, a.workspace_bytes = ptr(loss_out), ptr(workspace), workspace.numel()
    # keep the tensors alive as long as the struct
    a._keep = (user, item, neg, last_user, last_item, hat_user, hat_item, theta, g_user, g_item, m_user, v_user, m_item,
               v_item, adam_state, g_theta, m_theta, v_theta, loss_out, workspace, stamp_user, stamp_item)
    a._lazy = (hat_user, m_user, v_user, stamp_user, hat_item, m_item, v_item, stamp_item, adam_state) if stamp_user is not None else None
    return a


def mf_step(args, flush=True):
    """One MF step.  With row stamps in ``args`` (row-lazy Adam) the tables are flushed afterwards unless ``flush=False``
    (then call ``adam_flush`` before reading them)."""
    check(lib().sml_mf_step(C.byref(args), stream()), "mf_step")
    if flush and args._lazy is not None:
        hu, mu, vu, su, hi, mi, vi, si, st = args._lazy
        adam_flush(hu, mu, vu, su, st)
        adam_flush(hi, mi, vi, si, st)


def tr_step(args):
    check(lib().sml_tr_step(C.byref(args), stream()), "tr_step")


def mf_epoch(args, n_total):
    """args.user/item/neg point at n_total triples; args.batch is the nominal batch size."""
    check(lib().sml_mf_epoch(C.byref(args), n_total, stream()), "mf_epoch")


def tr_epoch(args, n_total):
    check(lib().sml_tr_epoch(C.byref(args), n_total, stream()), "tr_epoch")


def run_mf_grads(args, d_rows=None, scores=None):
    check(lib().sml_run_mf_grads(C.byref(args), ptr(d_rows), ptr(scores), stream()), "run_mf_grads")


def pack_rows(tab, ids=None):
    """[*, 64] rows (optionally gathered by ids) -> packed tensor-core operand (uint8 tensor)."""
    n = tab.shape[0] if ids is None else ids.numel()
    out = torch.empty(int(lib().sml_packed_rows_bytes(n)), dtype=torch.uint8, device=tab.device)
    check(lib().sml_pack_rows(ptr(_f32(tab, "tab")), ptr(ids), n, tab.shape[1], ptr(out), stream()), "pack_rows")
    return out


def fullcat_ranks(user_tab, item_tab, users, pos_items, items_packed=None, item_id0=0, n_items=None, gt=None, eq=None, s_pos=None):
    """Full-catalog rank counts of (users[n], pos_items[n]) against every row of item_tab (or of a pre-packed
    item shard).  Returns (gt, eq) int32 [n]; pass gt/eq to accumulate over several item shards.  ``s_pos``: the
    positive scores when the positive item rows are not in ``item_tab`` (row-sharded catalog); ``pos_items`` are then ids
    in the shard's numbering (or -1: positive not in this shard)."""
    n = users.numel()
    up = pack_rows(user_tab, users)
    if s_pos is None:
        # the positive's score through the same tcgen05 3xTF32 arithmetic as the catalog scores: ties are exact
        s_pos = torch.empty(n, dtype=torch.float32, device=users.device)
        pp = pack_rows(item_tab, pos_items)
        check(lib().sml_fullcat_pos_scores(ptr(up), ptr(pp), n, ptr(s_pos), stream()), "fullcat_pos_scores")
    if items_packed is None:
        items_packed = pack_rows(item_tab)
        n_items = item_tab.shape[0]
    if gt is None:
        gt = torch.zeros(n, dtype=torch.int32, device=users.device)
        eq = torch.zeros(n, dtype=torch.int32, device=users.device)
    check(lib().sml_fullcat_rank(ptr(up), ptr(items_packed), ptr(s_pos), ptr(_i64(pos_items, "pos_items")), n, n_items, item_id0,
                                 ptr(gt), ptr(eq), stream()), "fullcat_rank")
    return gt, eq


def fullcat_pos_scores(user_rows, pos_rows):
    """[n, 64] user rows x [n, 64] item rows -> the n pair scores, computed by the full-catalog score GEMM's own arithmetic."""
    n = user_rows.shape[0]
    s = torch.empty(n, dtype=torch.float32, device=user_rows.device)
    up, pp = pack_rows(user_rows), pack_rows(pos_rows)          # both operands alive until the launch is enqueued
    check(lib().sml_fullcat_pos_scores(ptr(up), ptr(pp), n, ptr(s), stream()), "fullcat_pos_scores")
    return s


def fullcat_topk(user_tab, item_tab, users, k, items_packed=None, item_id0=0, n_items=None, exclude=None):
    """Full-catalog top-k: for every id of ``users`` the k (<= 64) highest-scoring items of ``item_tab`` (or of a pre-packed
    shard): (scores [n, k] descending, ids [n, k] int64).  ``exclude``: optional int64 [n] item id to skip per user."""
    n = users.numel()
    up = pack_rows(user_tab, users)
    if items_packed is None:
        items_packed = pack_rows(item_tab)
        n_items = item_tab.shape[0]
    scores = torch.empty(n, k, dtype=torch.float32, device=users.device)
    ids = torch.empty(n, k, dtype=torch.int64, device=users.device)
    ws = _workspace(lib().sml_fullcat_topk_workspace_bytes(n, n_items, k), users.device, "fullcat_topk")
    check(lib().sml_fullcat_topk(ptr(up), ptr(items_packed), ptr(exclude if exclude is None else _i64(exclude, "exclude")), n, n_items, item_id0,
                                 k, ptr(scores), ptr(ids), ptr(ws), ws.numel(), stream()), "fullcat_topk")
    return scores, ids


def philox_negatives(users, item_all, keys, span, seed, offset=0):
    """Device-side rejection sampler (one negative per entry of users); all arguments are int64 CUDA tensors."""
    neg = torch.empty_like(users)
    check(lib().sml_philox_negatives(ptr(_i64(users, "users")), users.numel(), ptr(_i64(item_all, "item_all")), item_all.numel(),
                                     ptr(_i64(keys, "keys")), keys.numel(), int(span), int(seed), int(offset), ptr(neg), stream()),
          "philox_negatives")
    return neg


def gather_pairs(last, hat, loc):
    """Owner side of the row exchange: [n, 128] = [last[loc] | hat[loc]]."""
    n = loc.numel()
    out = torch.empty(n, 2 * D, dtype=torch.float32, device=last.device)
    check(lib().sml_gather_pairs(ptr(last), ptr(hat), ptr(_i64(loc, "loc")), n, D, ptr(out), stream()), "gather_pairs")
    return out


def scatter_grads(g, hat, loc, d_rows, scale=1.0, l2=0.0):
    """Owner side: g[loc] += scale * d_rows + l2 * hat[loc]."""
    check(lib().sml_scatter_grads(ptr(g), ptr(hat), ptr(_i64(loc, "loc")), ptr(d_rows), loc.numel(), D, scale, l2, stream()),
          "scatter_grads")


def plain_mf_grads(user_tab, item_tab, user, item, neg, g_user, g_item, loss_out, loss=LOSS_BCE, l2_u=0.0, l2_i=0.0,
                   item_bias=None, g_item_bias=None):
    ws = _workspace(256 + 2 * 2048 * 4, user_tab.device, "plain_mf")
    check(lib().sml_plain_mf_grads(ptr(user_tab), ptr(item_tab), ptr(item_bias), ptr(_i64(user, "user")), ptr(_i64(item, "item")),
                                   ptr(_i64(neg, "neg")), user.numel(), user_tab.shape[1], loss, l2_u, l2_i, ptr(g_user),
                                   ptr(g_item), ptr(g_item_bias), ptr(loss_out), ptr(ws), ws.numel(), stream()),
          "plain_mf_grads")


def new_list_heads(n_rows, device):
    """int32 [n_rows] per-row occurrence-list heads for plain_mf_step (all -1 between steps)."""
    return torch.full((n_rows,), -1, dtype=torch.int32, device=device)


def plain_mf_step(user_tab, item_tab, m_user, v_user, m_item, v_item, head_user, head_item, user, item, neg, adam_state, lr,
                  loss_out, loss=LOSS_BCE, l2_u=0.0, l2_i=0.0, optimizer=OPT_ADAM_DENSE_EXACT, stamp_user=None, stamp_item=None):
    """One fused plain-MF step (gather - dot - loss - row gradients + L2 - Adam on the batch rows) in one kernel.
    OPT_ADAM_DENSE_EXACT needs ``adam_state = new_adam_state(dev, history=True)`` and row stamps; flush with
    ``adam_flush`` before reading the tables as a whole."""
    B = user.numel()
    ws = _workspace(lib().sml_plain_mf_step_workspace_bytes(B), user_tab.device, "plain_mf_step")
    check(lib().sml_plain_mf_step(ptr(_f32(user_tab, "user_tab")), ptr(_f32(item_tab, "item_tab")), ptr(m_user), ptr(v_user), ptr(m_item),
                                  ptr(v_item), ptr(stamp_user), ptr(stamp_item), ptr(_i32t(head_user, "head_user")),
                                  ptr(_i32t(head_item, "head_item")), ptr(_i64(user, "user")), ptr(_i64(item, "item")),
                                  ptr(_i64(neg, "neg")), B, user_tab.shape[1], loss, float(l2_u), float(l2_i), ptr(adam_state), float(lr),
                                  optimizer, ptr(loss_out), ptr(ws), ws.numel(), stream()), "plain_mf_step")
