"""CPU oracle for the SML per-period retraining hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the arithmetic the reference (zyang1580/SML,
mounted read-only at /root/reference while the repo is built) performs on the
hot path.  It is *not* product code: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it, and only as the checker.  The product path (``sml_b200``) never imports it
and has no CPU fallback.

Parity pin: the reference ships no golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the
build container by ``oracle/gen_golden.py`` (which imports /root/reference on
CPU torch) and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py``
checks every function below against those fixtures.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Arithmetic is fp32 by default (``dtype=np.float32``) and can
be switched to fp64 for tolerance arbitration.
"""
from __future__ import annotations

import numpy as np

GELU_ALPHA = 1.702
ADAM_BETA1 = 0.9
ADAM_BETA2 = 0.999
ADAM_EPS = 1e-8

THETA_KEYS = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
              "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")


# --------------------------------------------------------------------------
# activation  (model/conv_transfer.py:9-10)
# --------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gelu(x):
    """``x * sigmoid(1.702 x)`` -- model/conv_transfer.py:9-10."""
    return x * sigmoid(x.dtype.type(GELU_ALPHA) * x)


def gelu_grad(x):
    s = sigmoid(x.dtype.type(GELU_ALPHA) * x)
    return s + x * x.dtype.type(GELU_ALPHA) * s * (1 - s)


# --------------------------------------------------------------------------
# one_transfer  (model/conv_transfer.py:18-50)
# --------------------------------------------------------------------------
def init_theta(rng, d=64, rows=3, dtype=np.float32):
    """Random parameters with the reference's tensor shapes
    (model/conv_transfer.py:23-34): conv1 (10,1,rows,1), conv2 (5,10,1,1),
    fc1 (512, 5*d), fc2 (d, 512).  Uniform(-1/sqrt(fan_in), +) like torch."""
    def u(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(dtype)
    return {
        "conv1.weight": u((10, 1, rows, 1), rows), "conv1.bias": u((10,), rows),
        "conv2.weight": u((5, 10, 1, 1), 10), "conv2.bias": u((5,), 10),
        "fc1.weight": u((512, 5 * d), 5 * d), "fc1.bias": u((512,), 5 * d),
        "fc2.weight": u((d, 512), 512), "fc2.bias": u((d,), 512),
    }


def one_transfer_forward(theta, x, keep=False):
    """model/conv_transfer.py:37-50.  ``x``: [N, R, d] (R = conv kernel height).
    conv1 is a (R x 1) kernel => per latent dim k a 1->10 channel map over the R
    stacked rows; conv2 is 1x1, 10->5; flatten is channel-major (m*d + k)
    (``x.view(-1, hidden_dim*out_channel2)`` at :43 on a [N,5,1,d] tensor)."""
    dt = x.dtype
    N, R, d = x.shape
    W1 = theta["conv1.weight"].reshape(10, R).astype(dt)
    b1 = theta["conv1.bias"].astype(dt)
    W2 = theta["conv2.weight"].reshape(5, 10).astype(dt)
    b2 = theta["conv2.bias"].astype(dt)
    Wf1 = theta["fc1.weight"].astype(dt)
    bf1 = theta["fc1.bias"].astype(dt)
    Wf2 = theta["fc2.weight"].astype(dt)
    bf2 = theta["fc2.bias"].astype(dt)
    zc1 = np.einsum("cr,nrk->nck", W1, x) + b1[None, :, None]       # :38
    h1 = gelu(zc1)                                                    # :40
    zc2 = np.einsum("mc,nck->nmk", W2, h1) + b2[None, :, None]      # :42
    a = gelu(zc2).reshape(N, 5 * d)                                   # :43-44
    z1 = a @ Wf1.T + bf1                                              # :47
    f = gelu(z1)                                                      # :48
    y = f @ Wf2.T + bf2                                               # :49
    if keep:
        return y, dict(x=x, zc1=zc1, h1=h1, zc2=zc2, a=a, z1=z1, f=f)
    return y


def one_transfer_backward(theta, cache, dy):
    """Autograd of model/conv_transfer.py:37-50 written out by hand.
    Returns (dx [N,R,d], dtheta dict with the reference's tensor shapes)."""
    x, zc1, h1, zc2, a, z1, f = (cache[k] for k in ("x", "zc1", "h1", "zc2", "a", "z1", "f"))
    dt = x.dtype
    N, R, d = x.shape
    W1 = theta["conv1.weight"].reshape(10, R).astype(dt)
    W2 = theta["conv2.weight"].reshape(5, 10).astype(dt)
    Wf1 = theta["fc1.weight"].astype(dt)
    Wf2 = theta["fc2.weight"].astype(dt)
    g = {}
    g["fc2.weight"] = dy.T @ f
    g["fc2.bias"] = dy.sum(0)
    df = dy @ Wf2
    dz1 = df * gelu_grad(z1)
    g["fc1.weight"] = dz1.T @ a
    g["fc1.bias"] = dz1.sum(0)
    da = dz1 @ Wf1
    dzc2 = da.reshape(N, 5, d) * gelu_grad(zc2)
    g["conv2.weight"] = np.einsum("nmk,nck->mc", dzc2, h1).reshape(5, 10, 1, 1)
    g["conv2.bias"] = dzc2.sum((0, 2))
    dh1 = np.einsum("mc,nmk->nck", W2, dzc2)
    dzc1 = dh1 * gelu_grad(zc1)
    g["conv1.weight"] = np.einsum("nck,nrk->cr", dzc1, x).reshape(10, 1, R, 1)
    g["conv1.bias"] = dzc1.sum((0, 2))
    dx = np.einsum("cr,nck->nrk", W1, dzc1)
    return dx, g


# --------------------------------------------------------------------------
# ConvTransfer_com / ConvTransfer  (model/conv_transfer.py:52-135)
# --------------------------------------------------------------------------
def com_stack(x_t, x_hat):
    """model/conv_transfer.py:93-103: x_com = x_t * detach(x_hat) / ||x_t||_2
    (no eps => NaN rows when x_t is all-zero); stack [x_t, x_hat, x_com]."""
    x_com = x_t * x_hat
    nrm = np.sqrt((x_t ** 2).sum(-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        x_com = x_com / nrm[:, None]
    return np.stack([x_t, x_hat, x_com], axis=1)


def conv_transfer_com_forward(theta_net, x_t, x_hat, keep=False):
    """ConvTransfer_com.forward for one net (``type`` picks the net in the
    reference, model/conv_transfer.py:104-110)."""
    return one_transfer_forward(theta_net, com_stack(x_t, x_hat), keep=keep)


def conv_transfer_forward(theta_net, x_t, x_hat, is_user, keep=False):
    """ConvTransfer.forward, model/conv_transfer.py:57-69: 2-row stack; the
    *user* output is divided by its detached L2 norm (:62-63)."""
    x = np.stack([x_t, x_hat], axis=1)
    out = one_transfer_forward(theta_net, x, keep=keep)
    y, cache = out if keep else (out, None)
    if is_user:
        nrm = np.sqrt((y ** 2).sum(-1))
        yn = y / nrm[:, None]
        if keep:
            cache["out_norm"] = nrm
        y = yn
    return (y, cache) if keep else y


def run_mf_forward_backward(theta_user, theta_item, u_last, u_hat, i_last, i_hat, j_last, j_hat,
                            BCE=True, variant="com", norm=False):
    """ConvTransfer_com.run_MF (model/conv_transfer.py:113-135) or
    ConvTransfer.run_MF (:72-85, always BPR-sum) + hand-written autograd.

    Returns dict(loss, s_pos, s_neg, d_u_hat, d_i_hat, d_j_hat, g_user, g_item):
    row gradients flow only through the x_hat channel (x_com is built from
    ``x_hat.data.detach()``, :93); theta gradients for both nets (the item net
    sees the positive and the negative rows)."""
    if variant == "com":
        uo, cu = conv_transfer_com_forward(theta_user, u_last, u_hat, keep=True)
        io, ci = conv_transfer_com_forward(theta_item, i_last, i_hat, keep=True)
        jo, cj = conv_transfer_com_forward(theta_item, j_last, j_hat, keep=True)
    else:
        uo, cu = conv_transfer_forward(theta_user, u_last, u_hat, True, keep=True)
        io, ci = conv_transfer_forward(theta_item, i_last, i_hat, False, keep=True)
        jo, cj = conv_transfer_forward(theta_item, j_last, j_hat, False, keep=True)
        BCE = False
    dt = uo.dtype
    B = uo.shape[0]
    s_pos = (uo * io).sum(-1)                                          # :120
    s_neg = (uo * jo).sum(-1)                                          # :121
    if BCE:
        sp, sn = sigmoid(s_pos), sigmoid(s_neg)
        eps = dt.type(1e-15)
        pos_loss = -np.mean(np.log(sp + eps))                          # :124
        neg_loss = -np.mean(np.log((dt.type(1) - sn) + eps))           # :125
        loss = pos_loss + neg_loss
        ds_pos = -(sp * (1 - sp)) / (sp + eps) / dt.type(B)
        ds_neg = (sn * (1 - sn)) / ((dt.type(1) - sn) + eps) / dt.type(B)
    else:
        score = s_pos - s_neg                                          # :128
        un = None
        if norm:
            un = np.sqrt((uo ** 2).sum(-1))                            # :130 (NOT detached)
            score = score / un
        # -sum(logsigmoid(score))                                      # :134
        loss = np.sum(np.logaddexp(dt.type(0), -score))
        dscore = -sigmoid(-score)
        if norm:
            raise NotImplementedError("norm=True is never used by the SML path (main_yelp.py:104)")
        ds_pos, ds_neg = dscore, -dscore
    d_uo = ds_pos[:, None] * io + ds_neg[:, None] * jo
    d_io = ds_pos[:, None] * uo
    d_jo = ds_neg[:, None] * uo
    if variant != "com":
        # y = z / ||z||.detach()  => dz = dy / ||z||
        d_uo = d_uo / cu["out_norm"][:, None]
    dxu, gu = one_transfer_backward(theta_user, cu, d_uo)
    dxi, gi = one_transfer_backward(theta_item, ci, d_io)
    dxj, gj = one_transfer_backward(theta_item, cj, d_jo)
    g_item = {k: gi[k] + gj[k] for k in gi}
    return dict(loss=loss, s_pos=s_pos, s_neg=s_neg, u_new=uo, i_new=io, j_new=jo,
                d_u_hat=dxu[:, 1, :], d_i_hat=dxi[:, 1, :], d_j_hat=dxj[:, 1, :],
                g_user=gu, g_item=g_item)


# --------------------------------------------------------------------------
# torch.optim.Adam  (model/transfer.py:392-393; torch single-tensor formula)
# --------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr, weight_decay=0.0, beta1=ADAM_BETA1, beta2=ADAM_BETA2, eps=ADAM_EPS):
    """One dense Adam update, in place; ``step`` is the 1-based step number.
    Mirrors torch/optim/adam.py::_single_tensor_adam (amsgrad off, coupled L2)."""
    dt = p.dtype.type
    if weight_decay != 0.0:
        g = g + dt(weight_decay) * p
    m += (g - m) * dt(1.0 - beta1)                 # exp_avg.lerp_(grad, 1-beta1)
    v *= dt(beta2)
    v += dt(1.0 - beta2) * g * g                   # addcmul_
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / dt(np.sqrt(bc2)) + dt(eps)
    p -= dt(step_size) * (m / denom)
    return p, m, v


# --------------------------------------------------------------------------
# SML MF step  (model/transfer.py:463-511)
# --------------------------------------------------------------------------
def sml_mf_step(state, theta_user, theta_item, user, item, neg, lr, l2, step, variant="com"):
    """One body of HOT LOOP A.  ``state`` holds numpy arrays that are updated
    in place: last_user, last_item (w_{t-1}, constant), user_tab, item_tab
    (MFbase latent tables = w_hat params), m_user, v_user, m_item, v_item.
    Dense Adam: *every* row is updated (zero gradient for untouched rows),
    model/transfer.py:392 with nn.Embedding(sparse=False) (model/MF.py:21-24)."""
    r = run_mf_forward_backward(theta_user, theta_item,
                                state["last_user"][user], state["user_tab"][user],
                                state["last_item"][item], state["item_tab"][item],
                                state["last_item"][neg], state["item_tab"][neg],
                                BCE=True, variant=variant)
    dt = state["user_tab"].dtype.type
    wu, wi, wj = state["user_tab"][user], state["item_tab"][item], state["item_tab"][neg]
    l2loss = dt(0.5) * np.sum(wu ** 2 + wi ** 2 + wj ** 2)             # :486
    loss = r["loss"] + dt(l2) * l2loss                                   # :488
    gu = np.zeros_like(state["user_tab"])
    gi = np.zeros_like(state["item_tab"])
    np.add.at(gu, user, r["d_u_hat"] + dt(l2) * wu)
    np.add.at(gi, item, r["d_i_hat"] + dt(l2) * wi)
    np.add.at(gi, neg, r["d_j_hat"] + dt(l2) * wj)
    adam_step(state["user_tab"], gu, state["m_user"], state["v_user"], step, lr)
    adam_step(state["item_tab"], gi, state["m_item"], state["v_item"], step, lr)
    return loss, gu, gi


# --------------------------------------------------------------------------
# SML transfer step  (model/transfer.py:701-728)
# --------------------------------------------------------------------------
def sml_tr_step(theta_user, theta_item, opt, tabs, user, item, neg, lr, wd, step, variant="com"):
    """One body of HOT LOOP B: run_MF on snapshot rows, theta-only gradients,
    Adam with coupled L2 ``wd`` (=TR_l2).  ``opt`` = {"user": {key: (m, v)},
    "item": {...}}; thetas updated in place."""
    r = run_mf_forward_backward(theta_user, theta_item,
                                tabs["last_user"][user], tabs["user_hat"][user],
                                tabs["last_item"][item], tabs["item_hat"][item],
                                tabs["last_item"][neg], tabs["item_hat"][neg],
                                BCE=True, variant=variant)
    for name, th, g in (("user", theta_user, r["g_user"]), ("item", theta_item, r["g_item"])):
        for k in THETA_KEYS:
            m, v = opt[name][k]
            adam_step(th[k], g[k].reshape(th[k].shape).astype(th[k].dtype), m, v, step, lr, weight_decay=wd)
    return r["loss"], r["g_user"], r["g_item"]


# --------------------------------------------------------------------------
# plain MF steps (baselines / MF2)
# --------------------------------------------------------------------------
def plain_mf_bce_grads(user_tab, item_tab, user, item, neg, l2_u, l2_i):
    """model/baseline.py:188-201: MFbasemode.forward twice, -(mean log sig(s+) +
    mean log(1 - sig(s-))) + L2; returns (loss, dense grad user, dense grad item)."""
    dt = user_tab.dtype.type
    wu, wi, wj = user_tab[user], item_tab[item], item_tab[neg]
    B = len(user)
    sp_, sn_ = (wu * wi).sum(-1), (wu * wj).sum(-1)
    sp, sn = sigmoid(sp_), sigmoid(sn_)
    eps = dt(1e-15)
    bce = np.mean(np.log(sp + eps)) + np.mean(np.log((dt(1.0) - sn) + eps))
    l2 = dt(l2_u) * dt(0.5) * np.sum(wu ** 2) + dt(l2_i) * dt(0.5) * (np.sum(wi ** 2) + np.sum(wj ** 2))
    loss = -bce + l2
    dsp = -(sp * (1 - sp)) / (sp + eps) / dt(B)
    dsn = (sn * (1 - sn)) / ((dt(1.0) - sn) + eps) / dt(B)
    gu = np.zeros_like(user_tab)
    gi = np.zeros_like(item_tab)
    np.add.at(gu, user, dsp[:, None] * wi + dsn[:, None] * wj + dt(l2_u) * wu)
    np.add.at(gi, item, dsp[:, None] * wu + dt(l2_i) * wi)
    np.add.at(gi, neg, dsn[:, None] * wu + dt(l2_i) * wj)
    return loss, gu, gi


def plain_mf_bpr_grads(user_tab, item_tab, user_bias, item_bias, user, item, neg):
    """MF2.forward train branch, model/MF.py:129-147: score = (b_u + b_i + <u,i>) -
    (b_u + b_j + <u,j>); loss = -sum(logsigmoid(score)).  (Its 'l2loss' output is
    a separate return value the caller may ignore; not differentiated here.)
    Returns (loss, gu, gi, g_item_bias); the user bias cancels."""
    dt = user_tab.dtype.type
    wu, wi, wj = user_tab[user], item_tab[item], item_tab[neg]
    score = ((user_bias[user, 0] + item_bias[item, 0] + (wu * wi).sum(-1))
             - (user_bias[user, 0] + item_bias[neg, 0] + (wu * wj).sum(-1)))
    loss = np.sum(np.logaddexp(dt(0), -score))
    ds = -sigmoid(-score)
    gu = np.zeros_like(user_tab)
    gi = np.zeros_like(item_tab)
    gb = np.zeros_like(item_bias)
    np.add.at(gu, user, ds[:, None] * (wi - wj))
    np.add.at(gi, item, ds[:, None] * wu)
    np.add.at(gi, neg, -ds[:, None] * wu)
    np.add.at(gb[:, 0], item, ds)
    np.add.at(gb[:, 0], neg, -ds)
    return loss, gu, gi, gb


# --------------------------------------------------------------------------
# candidate-list evaluation  (model/MF.py:45-80, evalution/evaluation2.py:8-26)
# --------------------------------------------------------------------------
def candidate_scores(user_tab, item_tab, rows):
    """model/MF.py:46-56: column 0 = user id, columns 1.. = candidate item ids
    (candidate 0 = the positive)."""
    u = user_tab[rows[:, 0]]
    c = item_tab[rows[:, 1:]]
    return np.einsum("nd,ncd->nc", u, c)


def candidate_ranks(scores):
    """Rank position of candidate 0 in a descending sort of each row:
    ``gt`` = #{j : s_j > s_0}, ``eq`` = #{j != 0 : s_j == s_0}.  torch.topk's tie
    order is unspecified (SURVEY.md section 7 hard part 3); the CPU reference was
    observed to place index 0 *after* its ties, so rank = gt + eq.  NaN sorts
    as larger than every number (torch.topk semantics)."""
    s0 = scores[:, :1]
    rest = scores[:, 1:]
    n0, nr = np.isnan(s0), np.isnan(rest)
    with np.errstate(invalid="ignore"):
        gt = ((rest > s0) | (nr & ~n0)).sum(-1)
        eq = ((rest == s0) | (nr & n0)).sum(-1)
    return gt.astype(np.int32), eq.astype(np.int32)


def mf_test(user_tab, item_tab, rows, topK):
    """MFbasemode.test (model/MF.py:45-80) -> (n_hits, ndcg_sum, hit_row_idx)."""
    gt, eq = candidate_ranks(candidate_scores(user_tab, item_tab, rows))
    rank = gt + eq
    hit = rank < topK
    ndcg = (np.float32(1.0) / np.log2(rank[hit].astype(np.float32) + np.float32(2.0))).sum(dtype=np.float32)
    return float(hit.sum()), ndcg, np.nonzero(hit)[0]


def test_model(user_tab, item_tab, rows, topK, batch=1024):
    """evaluation2.test_model (evalution/evaluation2.py:8-26) -> (recall, ndcg)."""
    hits, nd = 0.0, 0.0
    for s in range(0, rows.shape[0], batch):
        h, n, _ = mf_test(user_tab, item_tab, rows[s:s + batch], topK)
        hits += h
        nd += float(n)
    return hits / rows.shape[0], nd / rows.shape[0]


def legacy_rec_ndcg(scores, n_pos, topK):
    """evalution/evaluation.py:34-60 + evalution_function.py:26-39,86-95 for one
    user: positives occupy indices [0, n_pos); Rec = hits/n_pos,
    NDCG = sum 1/log2(rank+2) / IDCG(n_pos) over the top-K list."""
    order = np.argsort(-scores, kind="stable")[:topK]
    pos_rank = np.nonzero(order < n_pos)[0]
    idcg = (1.0 / np.log2(np.arange(n_pos, dtype=np.float32) + 2)).sum(dtype=np.float32)
    if len(pos_rank) == 0:
        return 0.0, 0.0
    dcg = (1.0 / np.log2(pos_rank.astype(np.float32) + 2)).sum(dtype=np.float32)
    return len(pos_rank) / n_pos, float(dcg / idcg)


# --------------------------------------------------------------------------
# updata()  (model/transfer.py:884-902)
# --------------------------------------------------------------------------
def updata(theta_user, theta_item, last_user, user_hat, last_item, item_hat, variant="com"):
    """w_t = Transfer(w_{t-1}, w_hat) over every row of both tables."""
    if variant == "com":
        return (conv_transfer_com_forward(theta_user, last_user, user_hat),
                conv_transfer_com_forward(theta_item, last_item, item_hat))
    return (conv_transfer_forward(theta_user, last_user, user_hat, True),
            conv_transfer_forward(theta_item, last_item, item_hat, False))
