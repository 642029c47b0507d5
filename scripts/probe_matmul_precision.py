"""What precision do torch's fp32 matmuls have on this box?  (diagnostic for the DNN-gradient parity test)"""
import os, torch
print("allow_tf32", torch.backends.cuda.matmul.allow_tf32, "precision", torch.get_float32_matmul_precision(),
      {k: v for k, v in os.environ.items() if "TF32" in k or "CUBLAS" in k})
torch.manual_seed(0)
for (M, N, K) in [(256, 2496, 8192), (256, 2496, 65536), (8192, 256, 2496)]:
    a = torch.randn(K, M, device="cuda") * 1e-3
    b = torch.randn(K, N, device="cuda")
    c = a.t() @ b
    ref = (a.double().t() @ b.double())
    err = (c.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"({M}x{K}) @ ({K}x{N}): max-norm rel err vs fp64 {err:.2e}")
