"""Short run of the embedding+FM forward/backward on the Criteo shape: the ncu capture target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
for s in range(3):
    batch = W.synthetic_batch(schema, B, seed=s, device="cuda")
    emb.zero_grad(set_to_none=True)
    fo, fe, fl = emb(batch)
    torch.autograd.backward([fl, fo, fm(fe)], [g_flat, torch.ones_like(fo), torch.ones_like(fo)])
torch.cuda.synchronize()
print("ok")
