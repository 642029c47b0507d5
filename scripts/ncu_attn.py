"""Short run of the field self-attention block at BASELINE config 3 (B=65536, F=16, D=16, 4 heads, A=64): ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.attention import MultiHeadSelfAttention

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
att = MultiHeadSelfAttention(16, 4, 64, 1, True).cuda()
x = torch.randn(B, 16, 16, device="cuda", requires_grad=True)
g = torch.randn(B, 16, 16, device="cuda")
for _ in range(3):
    att.zero_grad(set_to_none=True); x.grad = None
    att(x).backward(g)
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
a.record(); y = att(x); b.record(); y.backward(g); c.record(); torch.cuda.synchronize()
flops = B * (2 * 16 * 16 * 3 * 64 + 4 * 4 * 16 * 16 * 16 + 2 * 16 * 64 * 16)
print(f"attention B={B}: fwd {a.elapsed_time(b):.3f} ms ({flops / a.elapsed_time(b) / 1e9:.1f} TFLOP/s), bwd {b.elapsed_time(c):.3f} ms")
