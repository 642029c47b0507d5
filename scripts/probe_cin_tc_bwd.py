"""Development check of the tcgen05 CIN backward (data gradients) against the fp32 path (+ timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.cin import CIN
torch.manual_seed(0)
def rel(a, b): return ((a - b).norm() / (b.norm() + 1e-12)).item()
def run(B, F, D, sizes, split=True, time_it=False):
    cin = CIN(F, D, sizes, split).cuda()
    x = (torch.randn(B, F, D, device="cuda") * 0.5)
    g = torch.randn(B, cin.output_dim, device="cuda")
    res = {}
    for prec in ("fp32", "tf32"):
        cin.precision = prec
        cin.zero_grad(set_to_none=True)
        xt = x.clone().requires_grad_(True)
        out = cin(xt)
        out.backward(g)
        torch.cuda.synchronize()
        res[prec] = (out.detach(), xt.grad.clone(), [c.weight.grad.clone() for c in cin.conv_layers])
    print(f"B={B} F={F} D={D} sizes={sizes}: out {rel(res['tf32'][0], res['fp32'][0]):.2e}  gx {rel(res['tf32'][1], res['fp32'][1]):.2e}  "
          f"gW {[round(rel(a, b), 5) for a, b in zip(res['tf32'][2], res['fp32'][2])]}  finite={bool(torch.isfinite(res['tf32'][1]).all())}", flush=True)
    if time_it:
        for prec in ("fp32", "tf32"):
            cin.precision = prec
            def step():
                cin.zero_grad(set_to_none=True)
                xt = x.clone().requires_grad_(True)
                cin(xt).backward(g)
            for _ in range(2): step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3): step()
            b.record(); torch.cuda.synchronize()
            print(f"   {prec}: fwd+bwd {a.elapsed_time(b) / 3:.3f} ms", flush=True)

run(4, 16, 16, [64])
run(300, 16, 16, [64])
run(130, 16, 16, [128, 128, 64])
run(65, 39, 64, [24, 20])
run(37, 7, 12, [9, 5, 3], split=False)
run(16384, 16, 16, [128, 128, 64], time_it=True)
run(4096, 39, 64, [128, 128], time_it=True)
