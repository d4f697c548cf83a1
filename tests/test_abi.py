"""CPU: the C-ABI library loads and exports every symbol include/sml_b200.h declares; without a
GPU every entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from sml_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sml_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    l = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(l, s), "libsml_b200.so does not export %s" % s
    assert set(_lib.exported_symbols()) <= set(syms)


def test_abi_version_and_layout_constants():
    l = _lib.lib()
    assert l.sml_abi_version() == _lib.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "sml_b200.h")).read()
    for name in ("OFF_C1W", "OFF_C1B", "OFF_C2W", "OFF_C2B", "OFF_F1W", "OFF_F1B", "OFF_F2W", "OFF_F2B", "NET_STRIDE"):
        m = re.search(r"#define SML_%s (\d+)" % name, hdr)
        assert m and int(m.group(1)) == getattr(_lib, name), name
    assert _lib.OFF_F1B - _lib.OFF_F1W == 512 * 320 and _lib.OFF_F2B - _lib.OFF_F2W == 64 * 512
    assert all(getattr(_lib, n) % 32 == 0 for n in ("OFF_C1B", "OFF_C2W", "OFF_C2B", "OFF_F1W", "OFF_F1B", "OFF_F2W", "OFF_F2B", "NET_STRIDE"))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    l = _lib.lib()
    assert l.sml_device_check() != 0
    assert b"no CPU path" in l.sml_last_error() or b"CUDA" in l.sml_last_error()
    from sml_b200 import ops
    with pytest.raises(RuntimeError):
        ops.pair_scores(torch.zeros(2, 64), torch.zeros(2, 64), torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64))


def test_flat_theta_aliases_parameters():
    import torch
    from sml_b200.model.conv_transfer import ConvTransfer_com, ConvTransfer
    torch.manual_seed(0)
    m = ConvTransfer_com(64, 64)
    keys = list(m.state_dict().keys())
    assert keys[0] == "user_transfer.conv1.weight" and "item_transfer.fc2.bias" in keys and len(keys) == 16
    assert m.state_dict()["user_transfer.fc1.weight"].shape == (512, 320)
    assert m.state_dict()["user_transfer.conv1.weight"].shape == (10, 1, 3, 1)
    th = m.theta
    assert th.numel() == 2 * _lib.NET_STRIDE
    m.item_transfer.fc2.bias.data.fill_(3.0)
    assert float(th[_lib.NET_STRIDE + _lib.OFF_F2B]) == 3.0 and float(th[_lib.NET_STRIDE + _lib.OFF_F2B + 63]) == 3.0
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == 2 * 197279                       # 394 558 fp32 parameters
    assert float(th.abs().sum()) == pytest.approx(float(sum(p.detach().abs().sum() for p in m.parameters())), rel=1e-6)
    # same seed => same init as plain torch modules constructed in the reference's order
    torch.manual_seed(0)
    import torch.nn as nn
    c1 = nn.Conv2d(1, 10, (3, 1)); c2 = nn.Conv2d(10, 5, (1, 1)); f1 = nn.Linear(320, 512)
    assert torch.equal(m.user_transfer.conv1.weight.data, c1.weight.data) and torch.equal(m.user_transfer.fc1.weight.data, f1.weight.data)
    assert ConvTransfer(64, 64).state_dict()["user_transfer.conv1.weight"].shape == (10, 1, 2, 1)
    with pytest.raises(TypeError):
        m._net("nope")
