"""CPU port of the reference's hot path in plain PyTorch (TEST / BASELINE INFRASTRUCTURE ONLY).

/root/reference cannot travel to the GPU box, so the "reference arm" and the ``cpu_baseline`` leg
of bench.py time this port on the box's host cores instead ("kind": "port").  It executes the same
stock-PyTorch operator sequence the reference does -- nn.Embedding gathers, F.conv2d, F.linear,
autograd, dense torch.optim.Adam, torch.topk -- so its timings are representative of the
reference's own CPU path; tests/test_oracle_golden.py pins it to the reference's outputs.
Reference lines followed: model/conv_transfer.py:37-50,92-135; model/transfer.py:463-511,701-728,
884-902; model/MF.py:45-80; evalution/evaluation2.py:8-26.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _gelu(x):
    return x * torch.sigmoid(1.702 * x)                         # conv_transfer.py:9-10


class Net(torch.nn.Module):
    """one_transfer with kernel rows R (conv_transfer.py:18-50)."""

    def __init__(self, theta, rows=3):
        super().__init__()
        self.p = torch.nn.ParameterDict({k.replace(".", "_"): torch.nn.Parameter(torch.as_tensor(np.array(v))) for k, v in theta.items()})
        self.rows = rows

    def forward(self, x):                                       # x: [N,1,R,64]
        p = self.p
        x = _gelu(F.conv2d(x, p["conv1_weight"], p["conv1_bias"]))
        x = F.conv2d(x, p["conv2_weight"], p["conv2_bias"])
        x = _gelu(x.view(-1, 320))
        x = _gelu(F.linear(x, p["fc1_weight"], p["fc1_bias"]))
        return F.linear(x, p["fc2_weight"], p["fc2_bias"])


def com_forward(net, x_t, x_hat):                               # conv_transfer.py:92-110
    x_com = torch.mul(x_t, x_hat.detach()) / (x_t ** 2).sum(dim=-1).sqrt().unsqueeze(-1)
    x = torch.cat((x_t, x_hat, x_com), dim=-1).view(-1, 1, 3, x_t.shape[-1])
    return net(x)


def run_mf(user_net, item_net, ul, uh, il, ih, jl, jh):         # conv_transfer.py:113-126 (BCE)
    u, i, j = com_forward(user_net, ul, uh), com_forward(item_net, il, ih), com_forward(item_net, jl, jh)
    sp, sn = torch.mul(u, i).sum(dim=-1), torch.mul(u, j).sum(dim=-1)
    return -torch.mean(torch.log(torch.sigmoid(sp) + 1e-15)) - torch.mean(torch.log(1 - torch.sigmoid(sn) + 1e-15))


class Port(object):
    """State of one SML run on the CPU: MF tables (nn.Embedding, dense Adam), snapshots, two nets."""

    def __init__(self, user0, item0, theta_user, theta_item, mf_lr=0.01, l2=1e-6, tr_lr=0.001, tr_l2=1e-4, device="cpu"):
        """device="cuda": the same stock-PyTorch operator sequence on the GPU (cuBLAS / cuDNN / ATen kernels, TF32 off) --
        the "existing Blackwell implementation" bar of BASELINE.md section 3."""
        self.device = torch.device(device)
        self.user = torch.nn.Embedding.from_pretrained(torch.as_tensor(np.array(user0)), freeze=False).to(self.device)
        self.item = torch.nn.Embedding.from_pretrained(torch.as_tensor(np.array(item0)), freeze=False).to(self.device)
        self.user_net, self.item_net = Net(theta_user).to(self.device), Net(theta_item).to(self.device)
        self.last_user = self.user.weight.data.clone(); self.last_item = self.item.weight.data.clone()
        self.user_hat = self.user.weight.data.clone(); self.item_hat = self.item.weight.data.clone()
        self.mf_opt = torch.optim.Adam([self.user.weight, self.item.weight], lr=mf_lr, weight_decay=0)
        self.tr_opt = torch.optim.Adam(list(self.user_net.parameters()) + list(self.item_net.parameters()), lr=tr_lr, weight_decay=tr_l2)
        self.l2 = l2

    def mf_step(self, u, i, j):                                 # model/transfer.py:463-511
        u, i, j = (torch.as_tensor(x).long().to(self.device) for x in (u, i, j))
        self.mf_opt.zero_grad(); self.tr_opt.zero_grad()
        wu, wi, wj = self.user(u), self.item(i), self.item(j)
        loss = run_mf(self.user_net, self.item_net, self.last_user[u], wu, self.last_item[i], wi, self.last_item[j], wj)
        loss = loss + self.l2 * (0.5 * torch.sum(wu ** 2 + wi ** 2 + wj ** 2))
        loss.backward()
        self.mf_opt.step()
        return loss.detach()            # stays on the device like the reference's loss_all accumulation (model/transfer.py:501,722)

    def tr_step(self, u, i, j):                                 # model/transfer.py:701-728
        u, i, j = (torch.as_tensor(x).long().to(self.device) for x in (u, i, j))
        self.tr_opt.zero_grad()
        loss = run_mf(self.user_net, self.item_net, self.last_user[u], self.user_hat[u], self.last_item[i], self.item_hat[i],
                      self.last_item[j], self.item_hat[j])
        loss.backward()
        self.tr_opt.step()
        return loss.detach()            # stays on the device like the reference's loss_all accumulation (model/transfer.py:501,722)

    def save_last(self):                                        # model/transfer.py:925-927
        self.last_user.copy_(self.user.weight.data); self.last_item.copy_(self.item.weight.data)

    def save_hat(self):                                         # :928-933
        self.user_hat.copy_(self.user.weight.data); self.item_hat.copy_(self.item.weight.data)

    def updata(self, max_rows=None):                            # :884-902 (autograd recording on, as the reference)
        nu = self.last_user.shape[0] if max_rows is None else min(max_rows, self.last_user.shape[0])
        ni = self.last_item.shape[0] if max_rows is None else min(max_rows, self.last_item.shape[0])
        wu = com_forward(self.user_net, self.last_user[:nu], self.user_hat[:nu])
        wi = com_forward(self.item_net, self.last_item[:ni], self.item_hat[:ni])
        self.user.weight.data[:nu].copy_(wu); self.item.weight.data[:ni].copy_(wi)
        return nu + ni

    @torch.no_grad()
    def test(self, rows, topK):                                 # model/MF.py:45-80
        rows = torch.as_tensor(rows).long().to(self.device)
        ue = self.user(rows[:, 0]).unsqueeze(1)
        sc = torch.mul(ue, self.item(rows[:, 1:])).sum(-1)
        _, rank = torch.topk(sc, topK)
        pos = (rank < 1).nonzero()
        nd = (1 / torch.log2(pos[:, 1].float() + 2)).sum() if pos.shape[0] else torch.tensor(0.0)
        return float(pos.shape[0]), float(nd)

    def test_model(self, rows, topK, batch=1024):               # evalution/evaluation2.py:8-26
        h = n = 0.0
        for s in range(0, len(rows), batch):
            a, b = self.test(rows[s:s + batch], topK)
            h += a; n += b
        return h / len(rows), n / len(rows)
