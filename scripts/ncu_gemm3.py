"""The three products of the DNN tower's first layer at the bench shape (ncu capture target for dfm_gemm3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.dnn import _gemm3

M, N, K = 65536, 256, 2496
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.02; dy = torch.randn(M, N, device="cuda")
outs = [torch.empty(M, N, device="cuda"), torch.empty(M, K, device="cuda"), torch.empty(N, K, device="cuda")]
for _ in range(2):
    _gemm3(0, x, w, outs[0], None, M, N, K)
    _gemm3(1, dy, w, outs[1], None, M, K, N)
    _gemm3(2, dy, x, outs[2], None, N, K, M)
torch.cuda.synchronize()
print("ok")
