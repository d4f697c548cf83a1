"""CPU stand-in for the subset of sml_b200.ops that ShardedSML's MF path calls, built on the numpy oracle -- TEST
INFRASTRUCTURE (tests/test_shard_gloo.py).  It lets the world-size-2 gloo test drive the real exchange plan, the row-lazy
Adam bookkeeping (tick / catch-up / apply / flush order, stamps, duplicate ids) and the global-batch scaling of
``ShardedSML`` without a GPU; the arithmetic of each local operator is the oracle's."""
import numpy as np
import torch

from oracle import sml_oracle as O

ADAM_HISTORY = 4096
LOSS_BCE = 0
VARIANT_COM = 0
NET_STRIDE = 197344


def _np(t):
    return t.detach().numpy()


class _Lib(object):
    @staticmethod
    def sml_step_workspace_bytes(B):
        return 8


def lib():
    return _Lib


def step_rows(B):
    Bp = -(-B // 128) * 128
    total = Bp + -(-2 * B // 128) * 128
    return total, Bp, Bp + B


def new_adam_state(device, history=False):
    return torch.zeros(4, dtype=torch.int64)


def new_row_stamps(n_rows, state):
    return torch.full((n_rows,), int(state[0]), dtype=torch.int32)


_LR = {}


def adam_tick(state, lr, beta1=0.9, beta2=0.999):
    state[0] += 1
    _LR[state.data_ptr()] = lr


def _replay(p, m, v, g_row, row, step, lr):
    pr, mr, vr = p[row:row + 1], m[row:row + 1], v[row:row + 1]
    O.adam_step(pr, np.zeros_like(pr) if g_row is None else g_row[None, :], mr, vr, step, lr)


def adam_rows(p, m, v, g, stamp, ids, state, apply, **kw):
    t, lr = int(state[0]), _LR[state.data_ptr()]
    P, M, V = _np(p), _np(m), _np(v)
    for row in np.unique(_np(ids)):
        for s in range(int(stamp[row]) + 1, t):
            _replay(P, M, V, None, row, s, lr)
        if apply:
            if int(stamp[row]) < t:
                _replay(P, M, V, _np(g)[row].copy(), row, t, lr)
                g[row] = 0
            stamp[row] = t
        else:
            stamp[row] = max(int(stamp[row]), t - 1)


def adam_flush(p, m, v, stamp, state, **kw):
    t, lr = int(state[0]), _LR[state.data_ptr()]
    P, M, V = _np(p), _np(m), _np(v)
    for row in range(p.shape[0]):
        for s in range(int(stamp[row]) + 1, t + 1):
            _replay(P, M, V, None, row, s, lr)
        stamp[row] = t


def gather_pairs(last, hat, loc):
    return torch.cat([last[loc], hat[loc]], 1)


def scatter_grads(g, hat, loc, rows, scale, l2):
    g.index_add_(0, loc, scale * rows + l2 * hat[loc])


def make_step_args(**kw):
    return kw


THETA = [None, None]      # (user net, item net) oracle parameter dicts, set by the test


# flat theta layout of include/sml_b200.h (SML_OFF_*): views into one fp32 block [user net | item net]
_OFF = {"conv1.weight": 0, "conv1.bias": 32, "conv2.weight": 64, "conv2.bias": 128, "fc1.weight": 160, "fc1.bias": 164000,
        "fc2.weight": 164512, "fc2.bias": 197280}


def flat_theta(tu, ti):
    """-> (flat torch tensor [2 * NET_STRIDE], (user dict, item dict) of numpy VIEWS into it)."""
    flat = torch.zeros(2 * NET_STRIDE, dtype=torch.float32)
    buf = flat.numpy()
    views = []
    for n, th in enumerate((tu, ti)):
        d = {}
        for k in O.THETA_KEYS:
            lo = n * NET_STRIDE + _OFF[k]
            buf[lo:lo + th[k].size] = th[k].ravel()
            d[k] = buf[lo:lo + th[k].size].reshape(th[k].shape)
        views.append(d)
    return flat, views


def adam_dense(p, m, v, g, state, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, zero_grad=True):
    O.adam_step(_np(p), _np(g), _np(m), _np(v), int(state[0]), _LR[state.data_ptr()], weight_decay=weight_decay)
    if zero_grad:
        g.zero_()


def run_mf_grads(a, d_rows=None, scores=None):
    """a: the dict of make_step_args; tables are [n, 128] = [last | hat] pairs (table_pitch 128)."""
    assert a["table_pitch"] == 128
    u, i, j = _np(a["user"]), _np(a["item"]), _np(a["neg"])
    ru, ri = _np(a["last_user"]), _np(a["last_item"])
    r = O.run_mf_forward_backward(THETA[0], THETA[1], ru[u, :64], ru[u, 64:], ri[i, :64], ri[i, 64:],
                                  ri[j, :64], ri[j, 64:], BCE=True, variant="com")
    B = len(u)
    _, rp, rn = step_rows(B)
    if a.get("d_rows_by_id"):                          # the row gathered through id k writes d_rows[k] / d_rows[rp + k]
        d_rows[torch.from_numpy(u)] = torch.from_numpy(r["d_u_hat"])
        d_rows[rp + torch.from_numpy(i)] = torch.from_numpy(r["d_i_hat"]); d_rows[rp + torch.from_numpy(j)] = torch.from_numpy(r["d_j_hat"])
    else:
        d_rows[:B] = torch.from_numpy(r["d_u_hat"]); d_rows[rp:rp + B] = torch.from_numpy(r["d_i_hat"])
        d_rows[rn:rn + B] = torch.from_numpy(r["d_j_hat"])
    a["loss_out"][0] = float(r["loss"])
    if a["g_theta"] is not None:                       # accumulate like the kernels do (the caller zeroed it)
        gt = _np(a["g_theta"])
        for n, g in enumerate((r["g_user"], r["g_item"])):
            for k in O.THETA_KEYS:
                lo = n * NET_STRIDE + _OFF[k]
                gt[lo:lo + g[k].size] += np.asarray(g[k], dtype=np.float32).ravel()


# ---- full-catalog evaluation stand-ins (ShardedSML.eval_fullcat) -------------------------------------------
def fullcat_pos_scores(user_rows, pos_rows):
    return (user_rows * pos_rows).sum(-1)


def pair_scores(user_tab, item_tab, user, item, norm=False):
    return (user_tab[user] * item_tab[item]).sum(-1)


def pack_rows(tab, ids=None):
    return tab if ids is None else tab[ids]


def fullcat_ranks(user_tab, item_tab, users, pos_items, items_packed=None, item_id0=0, n_items=None, gt=None, eq=None, s_pos=None):
    """Counts over the rows of ``items_packed`` (this rank's shard); pos_items = local row of the positive or -1."""
    s = user_tab[users] @ items_packed[:n_items].T                      # [n, n_items]
    keep = torch.ones_like(s, dtype=torch.bool)
    has = pos_items >= 0
    keep[torch.nonzero(has).squeeze(1), pos_items[has] - item_id0] = False
    g = ((s > s_pos[:, None]) & keep).sum(1).to(torch.int32)
    e = ((s == s_pos[:, None]) & keep).sum(1).to(torch.int32)
    return g, e


def eval_reduce(gt, eq, topk, batch=1024, tie_loses=True):
    rank = (gt + eq).to(torch.int64)
    hit = rank < topk
    nd = torch.where(hit, 1.0 / torch.log2(rank.float() + 2.0), torch.zeros(()))
    nb = -(-gt.numel() // batch)
    return (torch.stack([hit[b * batch:(b + 1) * batch].sum() for b in range(nb)]).to(torch.int32),
            torch.stack([nd[b * batch:(b + 1) * batch].sum() for b in range(nb)]))
