"""Shared helpers for the parity tests (TEST INFRASTRUCTURE)."""
import numpy as np

from oracle import sml_oracle as O

D = 64


def theta_from_chk(chk):
    """Rebuild (theta_user, theta_item) from the [seed, rows, chk_user, chk_item] record the
    golden generator stored, and verify the checksums (numpy Generator stream drift guard)."""
    seed, rows = int(chk[0]), int(chk[1])
    out = []
    for off in (0, 1):
        th = O.init_theta(np.random.default_rng(seed + off), d=D, rows=rows)
        s = float(sum(np.abs(v.astype(np.float64)).sum() for v in th.values()))
        assert abs(s - chk[2 + off]) <= 1e-9 * abs(s), "theta regeneration drifted from the fixture"
        out.append(th)
    return out[0], out[1]


def sample(a):
    a = np.asarray(a)
    return a[::5, ::7] if a.ndim == 2 and a.size > 20000 else a


def rel_err(a, b, atol=1e-6):
    """max |a-b| relative to max |b|; differences below ``atol`` count as zero (quantities that are
    exactly 0 in the reference, e.g. the BPR fc2.bias gradient, come out as ~1e-8 rounding noise)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = float(np.abs(a - b).max())
    if d < atol:
        return 0.0
    return d / max(float(np.abs(b).max()), 1e-30)
