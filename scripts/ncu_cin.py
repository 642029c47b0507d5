"""ncu target: xDeepFM-Criteo-shaped CIN forward on tcgen05 (F=39, D=64, layers [128,128])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.cin import CIN
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cin = CIN(39, 64, [128, 128], True).cuda()
cin.precision = "tf32"
x = torch.randn(B, 39, 64, device="cuda") * 0.5
with torch.no_grad():
    for _ in range(3):
        cin(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        cin(x)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
flops = 2 * 64 * (128 * 39 * 39 + 128 * 64 * 39) * B
print(f"B={B} cin fwd {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (TF32 tcgen05)")
