"""Times the fc-layer GEMM kernels (tcgen05 3xTF32 vs SIMT fp32) on the step shapes with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200._lib import lib, check, ptr, stream  # noqa: E402

SHAPES = [  # name, M, N, K, a_mode, b_mode, epi, bn, transpose
    ("TR fc1 fwd", 768, 512, 320, 0, 0, 1, 128, 0), ("TR fc2 fwd", 768, 64, 512, 1, 0, 1, 64, 0),
    ("TR dZ1", 768, 512, 64, 0, 1, 2, 128, 0), ("TR dA", 768, 320, 512, 0, 1, 0, 64, 0),
    ("TR dW1", 512, 320, 512, 2, 1, 3, 64, 0), ("TR dW2^T", 512, 64, 512, 3, 1, 3, 64, 1),
    ("MF fc1 fwd", 3072, 512, 320, 0, 0, 1, 128, 0), ("MF dA", 3072, 320, 512, 0, 1, 0, 64, 0),
    ("updata fc1", 65536, 512, 320, 0, 0, 1, 128, 0), ("updata fc2", 65536, 64, 512, 1, 0, 1, 64, 0),
]


def main():
    dev = torch.device("cuda:0")
    l = lib()
    profile = os.environ.get("SML_PROFILE", "")          # under ncu: plain launches of the selected shapes, no graph
    for name, M, N, K, am, bm, epi, bn, tr in SHAPES:
        if profile and not any(k in name for k in profile.split(",")):
            continue
        A = torch.randn((M, K) if am in (0, 1) else (K, M), device=dev)
        B = torch.randn((N, K) if bm == 0 else (K, N), device=dev) * 0.1
        bias = torch.randn(N, device=dev); aux = torch.randn(M, N, device=dev)
        C = torch.zeros((N, M) if tr else (M, N), device=dev)
        res = []
        for tc in (1, 0):
            if tc == 0 and (tr or am == 3):
                res.append(float("nan")); continue
            def run():
                check(l.sml_debug_gemm(ptr(A), ptr(B), ptr(bias), ptr(aux), ptr(C), M, N, K, A.shape[1], B.shape[1], C.shape[1], am, bm,
                                       epi, tr, bn, tc, stream()))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            if profile:
                res.append(float("nan")); continue
            # a ctypes call costs ~15 us of Python: capture 20 launches in a CUDA graph so the GPU time is visible
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    run()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 20 * 1e3)
        fl = 2.0 * M * N * K
        print("%-12s M=%6d N=%4d K=%4d  tcgen05 %8.1f us (%6.1f TFLOP/s fp32-equiv)   simt %8.1f us (%5.1f TFLOP/s)"
              % (name, M, N, K, res[0], fl / res[0] / 1e6, res[1], fl / res[1] / 1e6))


if __name__ == "__main__":
    main()
