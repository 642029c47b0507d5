"""Diagnostics for dfm_gemm3 (which of the 3xTF32 terms / operand layouts are right)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.dnn import _gemm3

def hi(x): return (x.view(torch.int32) & -8192).view(torch.float32)
def rel(a, b): return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()

def report(tag, got, A, B):     # A (M,K), B (N,K) logical
    Ah, Bh = hi(A).double(), hi(B).double()
    Al, Bl = (A - hi(A)).double(), (B - hi(B)).double()
    full = A.double() @ B.double().t()
    hh = Ah @ Bh.t()
    print(f"{tag}: vs full {rel(got, full):.2e} | vs hh {rel(got, hh):.2e} | vs hh+hl {rel(got, hh + Ah @ Bl.t()):.2e} | "
          f"vs hh+lh {rel(got, hh + Al @ Bh.t()):.2e} | vs 3-term {rel(got, hh + Ah @ Bl.t() + Al @ Bh.t()):.2e}", flush=True)

torch.manual_seed(0)
for nomask in ("", "1"):
  if nomask: os.environ["DFM_G3_SPLIT_B_IN_KERNEL"] = "1"
  else: os.environ.pop("DFM_G3_SPLIT_B_IN_KERNEL", None)
  print("==== DFM_G3_SPLIT_B_IN_KERNEL =", repr(nomask), flush=True)
  for (M, N, K) in [(128, 32, 32), (128, 64, 64), (256, 128, 256), (4096, 256, 2496), (1000, 2496, 256)]:
      x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda")
      try:
          y = _gemm3(0, x, w, torch.zeros(M, N, device="cuda"), None, M, N, K); torch.cuda.synchronize()
          report(f"mode0 NT M{M} N{N} K{K}", y, x, w)
      except Exception as e:
          print("mode0 failed", M, N, K, e)
      dy = torch.randn(M, N, device="cuda")
      try:
          dx = _gemm3(1, dy, w, torch.zeros(M, K, device="cuda"), None, M, K, N); torch.cuda.synchronize()
          report(f"mode1 NN M{M} N{K} K{N}", dx, dy, w.t().contiguous())
      except Exception as e:
          print("mode1 failed", M, N, K, e)
      try:
          dw = _gemm3(2, dy, x, torch.zeros(N, K, device="cuda"), None, N, K, M); torch.cuda.synchronize()
          report(f"mode2 TN M{N} N{K} K{M}", dw, dy.t().contiguous(), x.t().contiguous())
      except Exception as e:
          print("mode2 failed", M, N, K, e)
os.environ.pop("DFM_G3_SPLIT_B_IN_KERNEL", None)
# layout probe: A = shifted identity, B[n][k] = 100 n + k  ->  D[m][n] = B[n][m]
M, N, K = 128, 32, 32
A = torch.zeros(M, K, device="cuda"); A[torch.arange(K), torch.arange(K)] = 1.0
B = (100 * torch.arange(N, device="cuda")[:, None] + torch.arange(K, device="cuda")[None, :]).float()
D = _gemm3(0, A, B, torch.zeros(M, N, device="cuda"), None, M, N, K); torch.cuda.synchronize()
print("layout mode0: D[0:4,0:4]=", D[0:4, 0:4].tolist(), "expected", B.t()[0:4, 0:4].tolist())
print("rows 32..35:", D[32:36, 0:2].tolist(), " nonzero rows:", (D.abs().sum(1) > 0).nonzero().flatten().tolist()[:40])

# timing of the three big products of the DNN's first layer at the bench shape
import time
M, N, K = 65536, 256, 2496
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.02; dy = torch.randn(M, N, device="cuda")
outs = [torch.empty(M, N, device="cuda"), torch.empty(M, K, device="cuda"), torch.empty(N, K, device="cuda")]
calls = [lambda: _gemm3(0, x, w, outs[0], None, M, N, K), lambda: _gemm3(1, dy, w, outs[1], None, M, K, N),
         lambda: _gemm3(2, dy, x, outs[2], None, N, K, M)]
for name, fn in zip(("fwd  Y=XW^T ", "dX=dY W     ", "dW=dY^T X   "), calls):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{name} {ms * 1e3:8.1f} us  {3 * 2 * M * N * K / ms / 1e9:7.1f} TFLOP/s tf32-equivalent (3 products), {2 * M * N * K / ms / 1e9:7.1f} useful", flush=True)
