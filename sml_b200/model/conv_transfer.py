"""Convolutional transfer networks -- drop-in for the reference's model/conv_transfer.py.

Class names, constructor signatures, sub-module names and state_dict keys follow the reference
(``user_transfer.fc1.weight`` (512, 320) ...).  The parameters of both nets live in ONE flat fp32
buffer (layout in include/sml_b200.h); the ``nn.Parameter`` objects are views into it, so the CUDA
kernels read theta through a single pointer while PyTorch-side code (state_dict, load_state_dict,
``.parameters()``, direct ``.data`` edits) keeps working.  All arithmetic runs in libsml_b200.so.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .._lib import (NET_STRIDE, OFF_C1W, OFF_C1B, OFF_C2W, OFF_C2B, OFF_F1W, OFF_F1B, OFF_F2W, OFF_F2B, VARIANT_COM,
                    VARIANT_CONV, LOSS_BCE, LOSS_BPR)

_SEGMENTS = (("conv1.weight", OFF_C1W), ("conv1.bias", OFF_C1B), ("conv2.weight", OFF_C2W), ("conv2.bias", OFF_C2B),
             ("fc1.weight", OFF_F1W), ("fc1.bias", OFF_F1B), ("fc2.weight", OFF_F2W), ("fc2.bias", OFF_F2B))


def Gelu(x):
    """reference: model/conv_transfer.py:9-10 (kept for API parity; elementwise torch expression)."""
    return x * torch.sigmoid(1.702 * x)


class one_transfer(nn.Module):
    """reference: model/conv_transfer.py:18-50.  Parameter container; same construction order as
    the reference (conv1, conv2, fc1, fc2) so the same torch seed gives the same initial theta."""

    def __init__(self, input_dim, out_dim, kernel=2):
        super(one_transfer, self).__init__()
        if input_dim != 64 or out_dim != 64:
            raise ValueError("sml_b200 kernels are specialised for the reference's default laten=64 "
                             "(main_yelp.py:41); got in=%d out=%d" % (input_dim, out_dim))
        self.hidden_dim = input_dim
        self.out_channel = 10
        self.conv1 = nn.Conv2d(1, self.out_channel, (kernel, 1), stride=1)
        self.out_channel2 = 5
        self.conv2 = nn.Conv2d(self.out_channel, self.out_channel2, (1, 1), stride=1)
        self.fc1 = nn.Linear(input_dim * self.out_channel2, 512)
        self.fc2 = nn.Linear(512, out_dim)
        self.kernel = kernel
        print("kernel:", kernel)          # the reference prints this (:36)

    def named_segments(self):
        mods = dict(conv1=self.conv1, conv2=self.conv2, fc1=self.fc1, fc2=self.fc2)
        for name, off in _SEGMENTS:
            m, p = name.split(".")
            yield name, off, getattr(mods[m], p)

    def forward(self, x):
        raise RuntimeError("one_transfer.forward is fused into ConvTransfer(_com).forward in sml_b200; "
                           "call the owning module")


class _FlatTransfer(nn.Module):
    """Shared machinery: flat theta / grad buffers and kernel dispatch."""
    variant = VARIANT_COM
    kernel_rows = 3

    def __init__(self, in_dim, out_dim):
        super(_FlatTransfer, self).__init__()
        self.user_transfer = one_transfer(in_dim, out_dim, kernel=self.kernel_rows)
        self.item_transfer = one_transfer(in_dim, out_dim, kernel=self.kernel_rows)
        self._flat = None
        self._gflat = None
        self._flatten()

    # -- flat storage ---------------------------------------------------------------------
    def _flatten(self):
        dev = self.user_transfer.fc1.weight.device
        flat = torch.zeros(2 * NET_STRIDE, dtype=torch.float32, device=dev)
        for base, net in ((0, self.user_transfer), (NET_STRIDE, self.item_transfer)):
            for _, off, p in net.named_segments():
                seg = flat[base + off: base + off + p.numel()].view(p.shape)
                seg.copy_(p.data)
                p.data = seg
        self._flat = flat
        self._gflat = None

    def _apply(self, fn, *a, **k):
        super(_FlatTransfer, self)._apply(fn, *a, **k)
        self._flatten()       # .cuda()/.to() re-materialise parameters; re-alias them into one block
        return self

    def load_state_dict(self, *a, **k):
        r = super(_FlatTransfer, self).load_state_dict(*a, **k)
        self._check_alias()
        return r

    def _check_alias(self):
        p = self.user_transfer.fc1.weight
        if p.data.data_ptr() != self._flat.data_ptr() + 4 * OFF_F1W:
            self._flatten()

    @property
    def theta(self):
        """flat fp32 [2 * NET_STRIDE] = [user net | item net] (device tensor, aliased by the parameters)."""
        self._check_alias()
        return self._flat

    @property
    def theta_grad(self):
        if self._gflat is None or self._gflat.device != self._flat.device:
            self._gflat = torch.zeros_like(self._flat)
        return self._gflat

    def grad_views(self):
        """{state_dict key: view of the flat gradient buffer}."""
        out = {}
        g = self.theta_grad
        for prefix, base, net in (("user_transfer.", 0, self.user_transfer), ("item_transfer.", NET_STRIDE, self.item_transfer)):
            for name, off, p in net.named_segments():
                out[prefix + name] = g[base + off: base + off + p.numel()].view(p.shape)
        return out

    def _net(self, type):
        if type == "user":
            return self.theta[:NET_STRIDE]
        if type == "item":
            return self.theta[NET_STRIDE:]
        raise TypeError("convtransfer has not this type")      # conv_transfer.py:67,109

    # -- run_MF -----------------------------------------------------------------------------
    def _run_mf(self, rows, loss_kind, need_grad):
        u_last, u_hat, i_last, i_hat, j_last, j_hat = (r.detach().contiguous().float() for r in rows)
        B = u_last.shape[0]
        dev = u_last.device
        ar = torch.arange(2 * B, dtype=torch.int64, device=dev)
        last_item = torch.cat([i_last, j_last]); hat_item = torch.cat([i_hat, j_hat])
        loss_out = torch.zeros(2, dtype=torch.float32, device=dev)
        total, rp, rn = ops.step_rows(B)
        d_rows = torch.zeros(total, 64, dtype=torch.float32, device=dev) if need_grad else None
        args = ops.make_step_args(user=ar[:B], item=ar[:B], neg=ar[B:], last_user=u_last, last_item=last_item,
                                  hat_user=u_hat, hat_item=hat_item, theta=self.theta, variant=self.variant, loss=loss_kind,
                                  g_theta=self.theta_grad if need_grad else None, loss_out=loss_out,
                                  workspace=ops.step_workspace(B, dev, "run_mf"))
        if need_grad:
            self.theta_grad.zero_()
        ops.run_mf_grads(args, d_rows=d_rows)
        if d_rows is not None:
            d_rows = torch.cat([d_rows[:B], d_rows[rp:rp + B], d_rows[rn:rn + B]])
        return loss_out[0], d_rows


class _RunMF(torch.autograd.Function):
    """loss = run_MF(rows...; theta): the kernels compute the loss together with dL/d(x_hat rows)
    and dL/d(theta); backward only scales them by the incoming gradient."""

    @staticmethod
    def forward(ctx, module, loss_kind, need_grad, u_last, u_hat, i_last, i_hat, j_last, j_hat, *params):
        loss, d_rows = module._run_mf((u_last, u_hat, i_last, i_hat, j_last, j_hat), loss_kind, need_grad)
        B = u_last.shape[0]
        ctx.module = module
        ctx.B = B
        if d_rows is not None:
            ctx.save_for_backward(d_rows, module.theta_grad.clone())
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        d_rows, gtheta = ctx.saved_tensors
        B = ctx.B
        m = ctx.module
        grads = [None, None, None, None, gout * d_rows[:B], None, gout * d_rows[B:2 * B], None, gout * d_rows[2 * B:]]
        for base, net in ((0, m.user_transfer), (NET_STRIDE, m.item_transfer)):
            for _, off, p in net.named_segments():
                grads.append(gout * gtheta[base + off: base + off + p.numel()].view(p.shape))
        return tuple(grads)


class ConvTransfer(_FlatTransfer):
    """reference: model/conv_transfer.py:52-85 (2-row stack, user output L2-normalised, BPR loss)."""
    variant = VARIANT_CONV
    kernel_rows = 2

    def __init__(self, in_dim, out_dim):
        super(ConvTransfer, self).__init__(in_dim, out_dim)

    def forward(self, x_t, x_hat, type):
        net = self._net(type)
        return ops.transfer_forward(x_t.detach().contiguous(), x_hat.detach().contiguous(), net, variant=self.variant,
                                    normalize_out=(type == "user"))

    def run_MF(self, user_weight_last, user_weight_hat, item_weight_last, item_weight_hat, negitem_weight_last,
               negitem_weight_hat, norm=False):
        if norm:
            raise NotImplementedError("norm=True is not used by the SML path (main_yelp.py:104 default False)")
        return _RunMF.apply(self, LOSS_BPR, torch.is_grad_enabled(), user_weight_last, user_weight_hat, item_weight_last, item_weight_hat,
                            negitem_weight_last, negitem_weight_hat, *self.parameters())


class ConvTransfer_com(_FlatTransfer):
    """reference: model/conv_transfer.py:87-135 -- the transfer SML uses (``--transfer_type conv_com``)."""
    variant = VARIANT_COM
    kernel_rows = 3

    def __init__(self, in_dim, out_dim):
        super(ConvTransfer_com, self).__init__(in_dim, out_dim)

    def forward(self, x_t, x_hat, type):
        """w = Transfer(x_t, x_hat) for every row (conv_transfer.py:92-110); used by updata()."""
        net = self._net(type)
        return ops.transfer_forward(x_t.detach().contiguous(), x_hat.detach().contiguous(), net, variant=self.variant)

    def run_MF(self, user_weight_last, user_weight_hat, item_weight_last, item_weight_hat, negitem_weight_last,
               negitem_weight_hat, norm=False, adpative=False, BCE=True):
        """conv_transfer.py:113-135.  Differentiable w.r.t. the three *_hat row tensors and theta."""
        if norm and not BCE:
            raise NotImplementedError("norm=True is not used by the SML path (main_yelp.py:104 default False)")
        return _RunMF.apply(self, LOSS_BCE if BCE else LOSS_BPR, torch.is_grad_enabled(), user_weight_last, user_weight_hat, item_weight_last,
                            item_weight_hat, negitem_weight_last, negitem_weight_hat, *self.parameters())
