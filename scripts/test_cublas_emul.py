"""Does the system cuBLAS 12.9 (fp32 emulation through 9 BF16 tensor-core GEMMs) load under torch and help the DNN GEMMs?"""
import ctypes, os, sys, time
mode = sys.argv[1] if len(sys.argv) > 1 else "emul"
if mode != "stock":
    ctypes.CDLL("/usr/local/cuda/lib64/libcublasLt.so.12", mode=ctypes.RTLD_GLOBAL)
    ctypes.CDLL("/usr/local/cuda/lib64/libcublas.so.12", mode=ctypes.RTLD_GLOBAL)
if mode == "emul":
    os.environ["CUBLAS_EMULATE_SINGLE_PRECISION"] = "1"
    os.environ.setdefault("CUBLAS_EMULATION_STRATEGY", "performant")
import torch
print("mode", mode, "torch", torch.__version__, "cublas", torch.backends.cuda.preferred_blas_library())
lib = ctypes.CDLL("libcublas.so.12")
v = ctypes.c_int()
try:
    lib.cublasGetProperty.restype = ctypes.c_int
    for i, n in enumerate(("major", "minor", "patch")):
        lib.cublasGetProperty(i, ctypes.byref(v)); print(n, v.value, end="  ")
    print()
except Exception as e:
    print("version query failed", e)
B = 65536
torch.manual_seed(0)
x = torch.randn(B, 2496, device="cuda"); w = torch.randn(256, 2496, device="cuda") * 0.02
g = torch.randn(B, 256, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for name, fn in (("fwd x@w.T", lambda: x @ w.t()), ("dX g@w", lambda: g @ w), ("dW g.T@x", lambda: g.t() @ x)):
    ms = t(fn)
    out = fn()
    print(f"{name}: {ms:.3f} ms")
xs, ws = x[:2048], w
ref = (xs.double() @ ws.double().t())
err = ((xs @ ws.t()).double() - ref).abs().max().item() / ref.abs().max().item()
print(f"max-norm rel err vs fp64 (fwd, 2048 rows): {err:.3e}")
