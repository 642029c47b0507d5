"""ncu target: xDeepFM-Criteo-shaped CIN forward + backward on tcgen05 (F=39, D=64, layers [128,128])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.cin import CIN
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cin = CIN(39, 64, [128, 128], True).cuda()
cin.precision = "tf32"
x = (torch.randn(B, 39, 64, device="cuda") * 0.5).requires_grad_(True)
g = torch.randn(B, cin.output_dim, device="cuda")
def step():
    cin.zero_grad(set_to_none=True); x.grad = None
    cin(x).backward(g)
for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
fwd = 2 * 64 * (128 * 39 * 39 + 128 * 64 * 39) * B
print(f"B={B} cin fwd+bwd {ms:.3f} ms  {3 * fwd / ms / 1e9:.1f} TFLOP/s algorithmic (6*D*sum(L*K) per sample, TF32 tcgen05)")
