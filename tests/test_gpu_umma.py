"""GPU: the tcgen05 3xTF32 GEMM against an fp64 reference (torch.matmul on the same device; this is a
floating-point kernel, so a plain torch reference is the right yardstick) and against the SIMT fp32
kernel, for every operand-layout / epilogue combination the steps use, including ragged tiles."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

A_MK, A_MK_GELU, A_KM, A_KM_GELU = 0, 1, 2, 3
B_NK, B_KN, B_KN_GELU = 0, 1, 2
EPI_NONE, EPI_BIAS, EPI_MULG, EPI_ACC = 0, 1, 2, 3


def gelu(x):
    return x * torch.sigmoid(1.702 * x)


def gelu_grad(x):
    s = torch.sigmoid(1.702 * x)
    return s + x * 1.702 * s * (1 - s)


def run_gemm(A, B, M, N, K, a_mode, b_mode, epi, bn, tc, transpose=False, bias=None, aux=None, C0=None):
    from sml_b200._lib import lib, check, ptr, stream
    dev = A.device
    C = (C0.clone() if C0 is not None else torch.zeros((N, M) if transpose else (M, N), device=dev))
    ldc = C.shape[1]
    check(lib().sml_debug_gemm(ptr(A), ptr(B), ptr(bias), ptr(aux), ptr(C), M, N, K, A.shape[1], B.shape[1], ldc, a_mode, b_mode,
                               epi, int(transpose), bn, int(tc), stream()), "debug_gemm")
    return C


def reference(A, B, a_mode, b_mode, epi, bias, aux, C0, transpose):
    a = A.double(); b = B.double()
    if a_mode in (A_MK_GELU, A_KM_GELU):
        a = gelu(A).double()
    if b_mode == B_KN_GELU:
        b = gelu(B).double()
    a = a if a_mode in (A_MK, A_MK_GELU) else a.t()
    b = b.t() if b_mode == B_NK else b
    c = a @ b
    if epi == EPI_BIAS:
        c = c + bias.double()
    if epi == EPI_MULG:
        c = c * gelu_grad(aux).double()
    if transpose:
        c = c.t()
    if epi == EPI_ACC:
        c = c + C0.double()
    return c


CASES = [
    # (M, N, K, a_mode, b_mode, epi, bn, transpose)           the step GEMMs (model/transfer.py:463-511,701-728)
    (768, 512, 320, A_MK, B_NK, EPI_BIAS, 128, False),        # fc1 forward
    (768, 64, 512, A_MK_GELU, B_NK, EPI_BIAS, 64, False),     # fc2 forward
    (768, 512, 64, A_MK, B_KN, EPI_MULG, 128, False),         # dZ1 = (dY W2) * g'(Z1)
    (768, 320, 512, A_MK, B_KN, EPI_NONE, 64, False),         # dA = dZ1 W1
    (512, 320, 256, A_KM, B_KN, EPI_ACC, 64, False),          # dW1 += dZ1^T A
    (512, 64, 256, A_KM_GELU, B_KN, EPI_ACC, 64, True),       # dW2^T += g(Z1)^T dY, stored transposed
    (248, 512, 320, A_MK, B_NK, EPI_BIAS, 128, False),        # ragged last batch (75000 % 1024 = 248)
    (130, 70, 45, A_MK, B_NK, EPI_NONE, 64, False),           # ragged everything
    (130, 70, 45, A_KM, B_KN_GELU, EPI_NONE, 64, False),
    (1, 8, 4, A_MK, B_NK, EPI_NONE, 64, False),
    (3072, 512, 320, A_MK, B_NK, EPI_BIAS, 128, False),       # MF step size
]


@pytest.mark.parametrize("case", CASES)
def test_umma_gemm_matches_fp64(case):
    M, N, K, a_mode, b_mode, epi, bn, transpose = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    pad4 = lambda n: (n + 3) // 4 * 4
    a_shape = (M, pad4(K)) if a_mode in (A_MK, A_MK_GELU) else (K, pad4(M))
    b_shape = (N, pad4(K)) if b_mode == B_NK else (K, pad4(N))
    A = torch.randn(a_shape, generator=g).to(dev); B = (torch.randn(b_shape, generator=g) * 0.1).to(dev)
    Aeff = A[:, :K] if a_mode in (A_MK, A_MK_GELU) else A[:, :M]
    Beff = B[:, :K] if b_mode == B_NK else B[:, :N]
    bias = torch.randn(N, generator=g).to(dev) if epi == EPI_BIAS else None
    aux = torch.randn(M, N, generator=g).to(dev) if epi == EPI_MULG else None
    C0 = torch.randn((N, M) if transpose else (M, N), generator=g).to(dev) if epi == EPI_ACC else None
    ref = reference(Aeff, Beff, a_mode, b_mode, epi, bias, aux, C0, transpose)
    scale = float(ref.abs().max())
    got_tc = run_gemm(A, B, M, N, K, a_mode, b_mode, epi, bn, True, transpose, bias, aux, C0)
    err_tc = float((got_tc.double() - ref).abs().max()) / scale
    print("case", case, "tcgen05 3xTF32 rel err %.3g" % err_tc)
    assert err_tc < 1e-5, "tcgen05 3xTF32 rel err %.3g" % err_tc
    if not transpose and a_mode != A_KM_GELU:
        got_simt = run_gemm(A, B, M, N, K, a_mode, b_mode, epi, bn, False, False, bias, aux, C0)
        err_simt = float((got_simt.double() - ref).abs().max()) / scale
        print("      SIMT fp32 rel err %.3g" % err_simt)
        assert err_simt < 1e-5, "SIMT rel err %.3g" % err_simt
