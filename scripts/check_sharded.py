"""torchrun target: the real multi-rank sharded path against the unsharded module (deepfm_b200/sharded_check.py).

    torchrun --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 scripts/check_sharded.py [deepfm|xdeepfm_multihot] [nccl]

Exits non-zero unless logits are bit-identical and every gradient is within 2e-5 (both transports: peer memory and, with
the `nccl` argument or DFM_SHARD_P2P=0, the NCCL all-to-all)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from deepfm_b200 import workloads as W
from deepfm_b200.sharded import TorchDistComm
from deepfm_b200.sharded_check import check_against_unsharded

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
comm = TorchDistComm()
which = sys.argv[1] if len(sys.argv) > 1 else "deepfm"
if "nccl" in sys.argv[2:]:
    os.environ["DFM_SHARD_P2P"] = "0"
multihot = which == "xdeepfm_multihot"
B = 2048 if multihot else 8192
schema = W.criteo_multihot_schema(64, 16, vocab_scale=0.002) if multihot else W.criteo_schema(64, vocab_scale=0.02)
cfg = bench.bench_config("xdeepfm_criteo_multihot" if multihot else "deepfm_criteo")
batch = W.synthetic_batch(schema, B, seed=10 + rank, device=dev)
y = W.synthetic_labels(B, seed=10 + rank, device=dev)
res = check_against_unsharded("xdeepfm" if multihot else "deepfm", schema, cfg, batch, y, comm,
                              replicate_below=60 if multihot else 4096)
ok = res["logits_bit_identical"] and res["table_grad_max_rel_err"] <= 2e-5 and res["dense_grad_max_rel_err"] <= 2e-5 \
    and res["table_rows_checked"] > 0
if rank == 0:
    print(json.dumps({"check": which, "transport": "p2p" if res["p2p"] else "nccl", "ok": bool(ok), **res}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
