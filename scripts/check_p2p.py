"""torchrun target: the peer-memory exchange must give bit-identical results to the NCCL all-to-all path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from deepfm_b200 import models as M, workloads as W
from deepfm_b200.sharded import ShardedFeatureEmbedding, TorchDistComm

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
comm = TorchDistComm()
multihot = len(sys.argv) > 1 and sys.argv[1] == "multihot"
B = 4096
schema = W.criteo_multihot_schema(64, 16, vocab_scale=0.01) if multihot else W.criteo_schema(64, vocab_scale=0.05)
M.BaseCTRModel.embedding_factory = staticmethod(lambda s, fm_embed_dim: ShardedFeatureEmbedding(s, fm_embed_dim, world, rank, comm))
torch.manual_seed(1)
with torch.device(dev):
    model = M.create_model("xdeepfm" if multihot else "deepfm", schema, bench.bench_config("deepfm_criteo"))
model.eval()                       # no dropout: the two passes must be comparable bit for bit
batch = W.synthetic_batch(schema, B, seed=rank, device=dev)
y = W.synthetic_labels(B, seed=rank, device=dev)
bce = torch.nn.BCEWithLogitsLoss()
out = {}
for mode in ("0", "1", "1"):
    os.environ["DFM_SHARD_P2P"] = mode
    model.zero_grad(set_to_none=True)
    logits = model(batch).squeeze(1)
    loss = bce(logits, y) + model.get_l2_reg_loss()
    loss.backward()
    rg = model.embedding.row_grads
    nv = int(rg.counts[0].item())
    heads = rg.heads()
    out[mode] = (logits.detach().clone(), rg.sorted_keys[:nv].clone(), rg.row_grad2[heads].clone(), rg.row_grad1[heads].clone(),
                 [p.grad.clone() for p in model.dnn.parameters()])
    used = model.embedding._px not in (None, False) and mode == "1"
    print(f"rank {rank} mode {mode}: loss {loss.item():.6f} p2p_used {used} keys {nv} unique {heads.numel()}", flush=True)
a, b = out["0"], out["1"]
ok = torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) \
    and all(torch.equal(x, z) for x, z in zip(a[4], b[4]))
print(f"rank {rank}: p2p == nccl bit-identical: {ok}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
