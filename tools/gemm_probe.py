"""Not product code: times v2f_gemm_tc on the small-M recurrent shapes (hot L2 and with an L2 flush between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visuelle2_multimodal_fusion_b200 import functional as Fv

shapes = [(128, 1536, 512), (128, 3072, 512), (128, 512, 1536), (256, 512, 512), (128, 512, 512), (128, 1536, 513 // 1 // 1 * 0 + 512),
          (128, 512, 3072), (6656, 1536, 512), (12800, 512, 2048), (12800, 512, 512)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kind in (1, 0):
    for (M, N, K) in shapes:
        dt = torch.float32 if kind == 1 else torch.bfloat16
        A = torch.randn(M, K, device="cuda").to(dt)
        B = torch.randn(N, K, device="cuda").to(dt)
        C = torch.empty(M, N, device="cuda")
        for _ in range(5):
            Fv.gemm_tc(kind, M, N, K, A, K, B, K, C, N)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        e0.record()
        for _ in range(n):
            Fv.gemm_tc(kind, M, N, K, A, K, B, K, C, N)
        e1.record()
        torch.cuda.synchronize()
        hot = e0.elapsed_time(e1) / n * 1e3
        cold = []
        for _ in range(20):
            flush.zero_()
            e0.record()
            Fv.gemm_tc(kind, M, N, K, A, K, B, K, C, N)
            e1.record()
            torch.cuda.synchronize()
            cold.append(e0.elapsed_time(e1) * 1e3)
        cold.sort()
        print(f"kind {kind} M={M:6d} N={N:5d} K={K:5d}  hot {hot:7.2f} us  ({2 * M * N * K / hot / 1e6:8.1f} TFLOP/s)   cold median {cold[10]:7.2f} us")
