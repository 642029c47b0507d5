"""Short run of the field self-attention block at BASELINE config 3 (B=65536, F=16, D=16, 4 heads, A=64): ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.attention import MultiHeadSelfAttention

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
att = MultiHeadSelfAttention(16, 4, 64, 1, True).cuda()
x = torch.randn(B, 16, 16, device="cuda", requires_grad=True)
g = torch.randn(B, 16, 16, device="cuda")
def one():
    att.zero_grad(set_to_none=True); x.grad = None
    y = att(x)
    y.backward(g)

for _ in range(5):
    one()
torch.cuda.synchronize()
n = 10
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    with torch.no_grad():
        att(x)
b.record(); torch.cuda.synchronize()
fwd = a.elapsed_time(b) / n
a.record()
for _ in range(n):
    one()
b.record(); torch.cuda.synchronize()
both = a.elapsed_time(b) / n
flops = B * (2 * 16 * 16 * 3 * 64 + 4 * 4 * 16 * 16 * 16 + 2 * 16 * 64 * 16)
print(f"attention B={B}: fwd {fwd:.3f} ms ({flops / fwd / 1e9:.1f} TFLOP/s), fwd+bwd {both:.3f} ms")
