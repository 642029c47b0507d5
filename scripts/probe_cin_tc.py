"""Development check of the tcgen05 CIN forward against the fp32 CUDA-core path (+ timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.cin import CIN

torch.manual_seed(0)
def run(B, F, D, sizes, split=True, time_it=False):
    cin = CIN(F, D, sizes, split).cuda()
    x = (torch.randn(B, F, D, device="cuda") * 0.5)
    cin.precision = "fp32"
    ref = cin(x)
    cin.precision = "tf32"
    out = cin(x)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    print(f"B={B} F={F} D={D} sizes={sizes}: rel max err {err:.3e}  finite={bool(torch.isfinite(out).all())}", flush=True)
    if time_it:
        for prec in ("fp32", "tf32"):
            cin.precision = prec
            with torch.no_grad():
                for _ in range(2): cin(x)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5): cin(x)
                b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            prev, flops = F, 0
            for i, L in enumerate(sizes):
                flops += 2 * D * L * prev * F
                prev = (L - L // 2) if (split and i < len(sizes) - 1) else L
            print(f"   {prec}: {ms:.3f} ms  {flops * B / ms / 1e9:.1f} TFLOP/s", flush=True)
    return err

run(4, 16, 16, [64])
run(300, 16, 16, [64])
run(130, 16, 16, [128, 128, 64])
run(65, 39, 64, [24, 20])
run(37, 7, 12, [9, 5, 3], split=False)
run(4096, 16, 16, [128, 128, 64], time_it=True)
run(8192, 39, 64, [128, 128], time_it=True)
