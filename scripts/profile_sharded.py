"""torchrun target: per-kernel time of one sharded DeepFM step on rank 0 (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
import bench
from deepfm_b200 import models as M, workloads as W
from deepfm_b200.sharded import ShardedFeatureEmbedding, TorchDistComm, allreduce_dense

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
comm = TorchDistComm()
schema = W.criteo_schema(64)
M.BaseCTRModel.embedding_factory = staticmethod(lambda s, fm_embed_dim: ShardedFeatureEmbedding(s, fm_embed_dim, world, rank, comm))
torch.manual_seed(1)
with torch.device(dev):
    model = M.create_model("deepfm", schema, bench.bench_config())
model.train()
emb = model.embedding
tid = {id(p) for p in emb.table_parameters()}
dense = [p for p in model.parameters() if id(p) not in tid]
batch = W.synthetic_batch(schema, 65536, seed=rank, device=dev)
y = W.synthetic_labels(65536, seed=rank, device=dev)
bce = torch.nn.BCEWithLogitsLoss()

def step():
    model.zero_grad(set_to_none=True)
    loss = bce(model(batch).squeeze(1), y) + model.get_l2_reg_loss()
    loss.backward()
    allreduce_dense(dense, world)

for _ in range(3):
    step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
if rank == 0:
    print("ms/step wall", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
dist.destroy_process_group()
