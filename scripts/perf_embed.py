"""Quick device timing of the embedding+FM path on the Criteo shape (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.layers.embedding import FeatureEmbedding
from deepfm_b200.layers.fm import FMInteraction
from deepfm_b200.layers.l2 import l2_penalty
from deepfm_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mode = sys.argv[3] if len(sys.argv) > 3 else "row_sparse"
schema = W.criteo_schema(64, vocab_scale=scale)
t0 = time.time()
with torch.device("cuda"):
    emb = FeatureEmbedding(schema, 64)
emb.grad_mode = mode
print("init s", time.time() - t0, "rows", emb._row_base[-1])
batches = [W.synthetic_batch(schema, B, seed=s, device="cuda") for s in range(4)]
fm = FMInteraction()
g_flat = torch.randn(B, schema.total_embedding_dim, device="cuda")

def ev():
    return torch.cuda.Event(enable_timing=True)

def timeit(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def fwd(i):
    with torch.no_grad():
        emb(batches[i % 4])

def fwd_bwd(i):
    emb.zero_grad(set_to_none=True)
    fo, fe, fl = emb(batches[i % 4])
    loss = (fl * g_flat).sum() + fo.sum() + fm(fe).sum() + l2_penalty(emb, 1e-5) * 0
    loss.backward()

def fwd_bwd_nol2(i):
    emb.zero_grad(set_to_none=True)
    fo, fe, fl = emb(batches[i % 4])
    loss = (fl * g_flat).sum() + fo.sum() + fm(fe).sum()
    loss.backward()

bytes_fwd = B * (26 * 268 + 52 + 9984 + 8)
t = timeit(fwd)
print(f"K1 fwd: {t:.3f} ms  {bytes_fwd / t / 1e6:.0f} GB/s algorithmic  ({B / t / 1e3:.1f} M samples/s)")
t = timeit(fwd_bwd_nol2)
print(f"fwd+bwd ({mode}, no L2): {t:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        fwd_bwd_nol2(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
