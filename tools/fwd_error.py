"""Relative error of the transfer forward (conv -> fc1 -> GELU -> fc2) against the fp64 numpy oracle."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sml_oracle as O  # noqa: E402 (checker)
from sml_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
th = O.init_theta(np.random.default_rng(3))
f = torch.zeros(_lib.NET_STRIDE)
for k, off in (("conv1.weight", _lib.OFF_C1W), ("conv1.bias", _lib.OFF_C1B), ("conv2.weight", _lib.OFF_C2W), ("conv2.bias", _lib.OFF_C2B),
               ("fc1.weight", _lib.OFF_F1W), ("fc1.bias", _lib.OFF_F1B), ("fc2.weight", _lib.OFF_F2W), ("fc2.bias", _lib.OFF_F2B)):
    f[off:off + th[k].size] = torch.from_numpy(th[k].reshape(-1))
n = 4096
xt = rng.standard_normal((n, 64)).astype(np.float32); xh = rng.standard_normal((n, 64)).astype(np.float32)
y = ops.transfer_forward(torch.from_numpy(xt).to(dev), torch.from_numpy(xh).to(dev), f.to(dev)).cpu().numpy().astype(np.float64)
th64 = {k: v.astype(np.float64) for k, v in th.items()}
ref = O.conv_transfer_com_forward(th64, xt.astype(np.float64), xh.astype(np.float64))
ref32 = O.conv_transfer_com_forward(th, xt, xh).astype(np.float64)
print("%s: max rel err vs fp64 %.3g (numpy fp32 oracle vs fp64: %.3g)" % (os.environ.get("SML_GEMM", "tcgen05"),
      np.abs(y - ref).max() / np.abs(ref).max(), np.abs(ref32 - ref).max() / np.abs(ref).max()))
