"""Timing + accuracy of dfm_gemm3 on the nine products of the DNN tower at the bench shape (B = 65536, 2496-256-128-64).
Environment switches are read by the library: DFM_G3_TS_ALL, DFM_G3_SPLIT_B_IN_KERNEL."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.dnn import _gemm3

def rel(a, b): return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()

torch.manual_seed(0)
print("env:", {k: v for k, v in os.environ.items() if k.startswith("DFM_G3")}, flush=True)
# accuracy on a mid-size problem of every mode (vs fp64)
M, N, K = 4096, 256, 2496
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.02; dy = torch.randn(M, N, device="cuda")
y = _gemm3(0, x, w, torch.empty(M, N, device="cuda"), None, M, N, K)
dx = _gemm3(1, dy, w, torch.empty(M, K, device="cuda"), None, M, K, N)
dw = _gemm3(2, dy, x, torch.empty(N, K, device="cuda"), None, N, K, M)
torch.cuda.synchronize()
print(f"max-norm rel err vs fp64: fwd {rel(y, x.double() @ w.double().t()):.2e}  dX {rel(dx, dy.double() @ w.double()):.2e}  "
      f"dW {rel(dw, dy.double().t() @ x.double()):.2e}", flush=True)
B = 65536
total = 0.0
for (din, dout) in [(2496, 256), (256, 128), (128, 64)]:
    x = torch.randn(B, din, device="cuda"); w = torch.randn(dout, din, device="cuda") * 0.02; dy = torch.randn(B, dout, device="cuda")
    outs = [torch.empty(B, dout, device="cuda"), torch.empty(B, din, device="cuda"), torch.empty(dout, din, device="cuda")]
    calls = [("fwd", lambda: _gemm3(0, x, w, outs[0], None, B, dout, din)), ("dX ", lambda: _gemm3(1, dy, w, outs[1], None, B, din, dout)),
             ("dW ", lambda: _gemm3(2, dy, x, outs[2], None, dout, din, B))]
    for name, fn in calls:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        total += ms
        fl = 2.0 * B * din * dout
        print(f"{din:5d}->{dout:4d} {name} {ms * 1e3:8.1f} us  {3 * fl / ms / 1e9:7.1f} TF/s tf32 (3 products)  {fl / ms / 1e9:6.1f} useful", flush=True)
print(f"sum of the nine products: {total * 1e3:.1f} us", flush=True)
