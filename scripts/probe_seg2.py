"""Embedding fwd/bwd on the Criteo shape: per-kernel device time of the backward for the seg2 tuning switches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from deepfm_b200.layers.embedding import FeatureEmbedding
from deepfm_b200.layers.fm import FMInteraction
from deepfm_b200 import workloads as W

B = 65536
schema = W.criteo_schema(64)
with torch.device("cuda"):
    emb = FeatureEmbedding(schema, 64)
emb.grad_mode = "row_sparse"
fm = FMInteraction()
g_flat = torch.randn(B, schema.total_embedding_dim, device="cuda")
batches = [W.synthetic_batch(schema, B, seed=s, device="cuda") for s in range(3)]

def step(i):
    emb.zero_grad(set_to_none=True)
    fo, fe, fl = emb(batches[i % 3])
    torch.autograd.backward([fl, fo, fm(fe)], [g_flat, torch.ones_like(fo), torch.ones_like(fo)])

for cfg, dbg in (("0", "0"),):
    for i in range(3): step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3): step(i)
        torch.cuda.synchronize()
    rows = [(e.key[:60], e.device_time_total / e.count) for e in prof.key_averages() if "seg2" in e.key or "stitch" in e.key or "dense_stream" in e.key or "embed_fwd" in e.key]
    print("DFM_SEG2_CFG =", cfg, "DFM_SEG2_DBG =", dbg, " | ".join(f"{k}: {t:.1f} us" for k, t in rows), flush=True)
