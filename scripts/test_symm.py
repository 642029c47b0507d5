"""torchrun target: does torch symmetric memory (peer pointers + device barrier) work on this box?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
g = dist.group.WORLD
n = 64 * 1024 * 1024
t = symm.empty((n,), dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, g.group_name)
print(rank, "rendezvous ok; world", hdl.world_size, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
t.fill_(float(rank))
hdl.barrier()
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (n,), torch.float32)
src = torch.full((n,), 100.0 + rank, device=dev)
hdl.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    pt.copy_(src)          # peer write over NVLink
b.record()
hdl.barrier()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(rank, f"peer write {n * 4 / 1e6:.0f} MB in {ms:.3f} ms = {n * 4 / ms / 1e6:.0f} GB/s; my buffer now holds", t[0].item(), t[-1].item(), flush=True)
dist.destroy_process_group()
