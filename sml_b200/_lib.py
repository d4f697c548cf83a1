"""ctypes binding of libsml_b200.so (the C ABI declared in include/sml_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsml_b200.so")

ABI_VERSION = 4
D = 64
OFF_C1W, OFF_C1B, OFF_C2W, OFF_C2B = 0, 32, 64, 128
OFF_F1W, OFF_F1B, OFF_F2W, OFF_F2B = 160, 164000, 164512, 197280
NET_STRIDE = 197344
VARIANT_COM, VARIANT_CONV = 3, 2
LOSS_BCE, LOSS_BPR = 0, 1
OPT_ADAM_DENSE_EXACT, OPT_ADAM_SPARSE = 0, 1

_vp, _i64, _i32, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_size_t


class StepArgs(C.Structure):
    """Mirror of sml_step_args (include/sml_b200.h)."""
    _fields_ = [
        ("user", _vp), ("item", _vp), ("neg", _vp), ("batch", _i64),
        ("last_user", _vp), ("last_item", _vp), ("hat_user", _vp), ("hat_item", _vp),
        ("n_users", _i64), ("n_items", _i64),
        ("theta", _vp), ("variant", _i32), ("loss", _i32),
        ("g_user", _vp), ("g_item", _vp), ("m_user", _vp), ("v_user", _vp), ("m_item", _vp), ("v_item", _vp),
        ("adam_state", _vp), ("lr", _dbl), ("l2", _dbl),
        ("g_theta", _vp), ("m_theta", _vp), ("v_theta", _vp),
        ("loss_out", _vp), ("workspace", _vp), ("workspace_bytes", _sz), ("table_pitch", _i64),
        ("stamp_user", _vp), ("stamp_item", _vp),
        ("adaptive_beta", _dbl), ("clip_max_norm", _dbl),
        ("d_rows_by_id", _i32),
    ]


_PROTOS = {
    "sml_abi_version": (_i32, []),
    "sml_last_error": (C.c_char_p, []),
    "sml_device_check": (_i32, []),
    "sml_sm_count": (_i32, []),
    "sml_launch_count": (C.c_uint64, []),
    "sml_eval_candidates": (_i32, [_vp, _vp, _i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "sml_eval_prefilter_bytes": (_sz, [_i64]),
    "sml_eval_candidates_prefilter": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "sml_eval_reduce": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "sml_pair_scores": (_i32, [_vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _vp]),
    "sml_transfer_fwd_workspace_bytes": (_sz, [_i64]),
    "sml_transfer_fwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "sml_adam_tick": (_i32, [_vp, _dbl, _dbl, _dbl, _vp]),
    "sml_adam_dense": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _dbl, _dbl, _dbl, _dbl, _i32, _vp]),
    "sml_sumsq": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "sml_adam_dense_clipped": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _dbl, _dbl, _dbl, _dbl, _i32, _vp, _dbl, _vp]),
    "sml_adam_rows": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _dbl, _dbl, _dbl, _vp]),
    "sml_adam_flush": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _dbl, _dbl, _dbl, _vp]),
    "sml_step_workspace_bytes": (_sz, [_i64]),
    "sml_step_rows": (_i64, [_i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "sml_mf_step": (_i32, [C.POINTER(StepArgs), _vp]),
    "sml_tr_step": (_i32, [C.POINTER(StepArgs), _vp]),
    "sml_mf_epoch": (_i32, [C.POINTER(StepArgs), _i64, _vp]),
    "sml_tr_epoch": (_i32, [C.POINTER(StepArgs), _i64, _vp]),
    "sml_run_mf_grads": (_i32, [C.POINTER(StepArgs), _vp, _vp, _vp]),
    "sml_packed_rows_bytes": (_sz, [_i64]),
    "sml_pack_rows": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "sml_fullcat_rank": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "sml_fullcat_pos_scores": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "sml_fullcat_topk_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "sml_fullcat_topk": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "sml_philox_negatives": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "sml_gather_pairs": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "sml_scatter_grads": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _dbl, _dbl, _vp]),
    "sml_host_keyset_build": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64]),
    "sml_host_rejection_walk_hashed": (_i64, [_vp, _i64, _vp, _i64, C.POINTER(_i64), _vp, _vp, _i64, _i64, _vp]),
    "sml_host_rejection_walk": (_i64, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _vp]),
    "sml_debug_set_mask": (_i32, [_i32]),
    "sml_debug_mask": (_i32, []),
    "sml_debug_set_ksplit": (_i32, [_i32, _i32, _i32, _i32]),
    "sml_debug_ksplit": (_i32, [_i32]),
    "sml_debug_gemm": (_i32, [_vp, _vp, _vp, _vp, _vp] + [_i32] * 12 + [_vp]),
    "sml_plain_mf_step_workspace_bytes": (_sz, [_i64]),
    "sml_plain_mf_step": (_i32, [_vp] * 13 + [_i64, _i32, _i32, _dbl, _dbl, _vp, _dbl, _i32, _vp, _vp, _sz, _vp]),
    "sml_plain_mf_grads": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def lib():
    """The loaded shared library.  Raises RuntimeError (never falls back) if it cannot be used."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "sml_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C sml_b200/csrc`). There is no CPU/PyTorch fallback." % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(l, name)        # AttributeError => stale library; fail loudly
        fn.restype, fn.argtypes = res, args
    if l.sml_abi_version() != ABI_VERSION:
        raise RuntimeError("sml_b200: ABI mismatch: library %d, python %d" % (l.sml_abi_version(), ABI_VERSION))
    _lib = l
    return l


def exported_symbols():
    return list(_PROTOS)


def check(rc, what=""):
    if rc != 0:
        msg = lib().sml_last_error().decode("utf-8", "replace")
        raise RuntimeError("sml_b200 %s failed (code %d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  The tensor must be CUDA + contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("sml_b200: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("sml_b200: expected a contiguous tensor")
    return t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
