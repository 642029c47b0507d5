"""fwd+bwd samples/s of the three models on the ML-100K-shaped synthetic workload (configs 1-3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfm_b200.config import ExperimentConfig
from deepfm_b200.models import create_model
from deepfm_b200 import workloads as W
from torch.profiler import profile, ProfilerActivity

schema = W.ml100k_schema()
def run(name, B, cin_sizes=None, precision="fp32", prof=False):
    cfg = ExperimentConfig()
    if cin_sizes: cfg.cin.layer_sizes = cin_sizes
    else: cfg.cin.layer_sizes = [64]
    torch.manual_seed(0)
    model = create_model(name, schema, cfg).cuda().train()
    if name == "xdeepfm": model.cin.precision = precision
    batch = W.synthetic_batch(schema, B, seed=0, device="cuda")
    y = W.synthetic_labels(B, seed=0, device="cuda")
    bce = torch.nn.BCEWithLogitsLoss()
    def step():
        model.zero_grad(set_to_none=True)
        loss = bce(model(batch).squeeze(1), y) + model.get_l2_reg_loss()
        loss.backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{name:18s} B={B:6d} cin={cin_sizes} {precision}: {ms:8.3f} ms/step  {B / ms / 1e3:8.2f} M samples/s", flush=True)
    if prof:
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            step(); torch.cuda.synchronize()
        print(p.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))

run("deepfm", 4096); run("deepfm", 65536)
run("xdeepfm", 4096); run("xdeepfm", 65536, precision="tf32")
run("xdeepfm", 65536, [128, 128, 64], "tf32", prof=True)
run("attention_deepfm", 4096); run("attention_deepfm", 65536, prof=True)
