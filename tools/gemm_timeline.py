import os, sys, ctypes, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visuelle2_multimodal_fusion_b200 import functional as Fv, _lib
L = _lib.lib()
for (M, N, K) in [(128, 1536, 512), (128, 512, 1536), (12800, 512, 512)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    for _ in range(5):
        Fv.gemm_tc(1, M, N, K, A, K, B, K, C, N, act=int(os.environ.get("ACT", "0")))
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 16)()
    L.v2f_debug_timeline(out)
    t = list(out)[:10]
    print(M, N, K, "ns since entry:", [x - t[0] for x in t])
