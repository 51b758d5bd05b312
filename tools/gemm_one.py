import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visuelle2_multimodal_fusion_b200 import functional as Fv
M, N, K = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (128, 1536, 512)
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
for _ in range(4):
    Fv.gemm_tc(1, M, N, K, A, K, B, K, C, N)
torch.cuda.synchronize()
