"""Fused transfer forward (k_transfer_fused) against the three-kernel path and an fp64 restatement; timing at table scale.
    python tools/fused_fwd_check.py [rows ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sml_oracle as O  # noqa: E402  (checker only)
from sml_b200 import ops  # noqa: E402
from sml_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
th = O.init_theta(np.random.default_rng(1))
flat = torch.zeros(ops.NET_STRIDE, device=dev)
from sml_b200._lib import OFF_C1W, OFF_C1B, OFF_C2W, OFF_C2B, OFF_F1W, OFF_F1B, OFF_F2W, OFF_F2B  # noqa: E402
for k, off in (("conv1.weight", OFF_C1W), ("conv1.bias", OFF_C1B), ("conv2.weight", OFF_C2W), ("conv2.bias", OFF_C2B),
               ("fc1.weight", OFF_F1W), ("fc1.bias", OFF_F1B), ("fc2.weight", OFF_F2W), ("fc2.bias", OFF_F2B)):
    v = torch.from_numpy(th[k].reshape(-1)).to(dev)
    flat[off:off + v.numel()] = v
res = {}
n = 1000
xt = rng.standard_normal((n, 64)).astype(np.float32); xh = rng.standard_normal((n, 64)).astype(np.float32)
y64 = O.conv_transfer_com_forward({k: v.astype(np.float64) for k, v in th.items()}, xt.astype(np.float64), xh.astype(np.float64))
T = lambda a: torch.from_numpy(a).to(dev)
for name, mask in (("fused", 0), ("three_kernel", 2048)):
    lib().sml_debug_set_mask(mask)
    y = ops.transfer_forward(T(xt), T(xh), flat).cpu().numpy().astype(np.float64)
    res["err_vs_fp64_" + name] = float(np.abs(y - y64).max() / np.abs(y64).max())
ids = torch.from_numpy(rng.integers(0, n, 777)).to(dev)
lib().sml_debug_set_mask(0)
ya = ops.transfer_forward(T(xt), T(xh), flat, ids=ids)
lib().sml_debug_set_mask(2048)
yb = ops.transfer_forward(T(xt), T(xh), flat, ids=ids)
res["gather_fused_vs_three_kernel"] = float((ya - yb).abs().max() / yb.abs().max())
for rows in [int(a) for a in sys.argv[1:]] or [59082, 122816, 4_000_000]:
    a = torch.randn(rows, 64, device=dev); b = torch.randn(rows, 64, device=dev); out = torch.empty_like(a)
    for name, mask in (("fused", 0), ("three_kernel", 2048)):
        lib().sml_debug_set_mask(mask)
        ops.transfer_forward(a, b, flat, out=out); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.transfer_forward(a, b, flat, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res["%s_%d" % (name, rows)] = dict(ms=ms, tflops_fp32_equiv=rows * 403456 / ms / 1e9, rows_per_s=rows / ms * 1e3)
lib().sml_debug_set_mask(0)
print(json.dumps(res, indent=1))
