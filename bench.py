#!/usr/bin/env python
"""bench.py -- train samples/sec (fwd+bwd) of DeepFM on the Criteo-shaped synthetic workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], the one the metric is quoted on): DeepFM, 13 DENSE + 26 SPARSE
fields, embedding_dim = fm_embed_dim = 64, Criteo-Kaggle cardinalities (33.76 M rows, 8.6 GB of
tables), DNN [256,128,64] + BatchNorm + dropout 0.1, batch 65536 PER GPU (weak scaling).
A step is the reference trainer's forward + loss + backward (trainer.py:219-229):
    logits = model(batch); loss = BCEWithLogits(logits, y) + model.get_l2_reg_loss(); loss.backward()
`value`  : inputs already resident in HBM (4 rotating batches), device-timed with CUDA events.
`e2e`    : the same step through the public module API with the batch in pinned HOST memory, the
           host->device copies (issued one step ahead on a copy stream) and the loss.item() read of every step
           inside the timed region.
`roofline`: the fused embedding+FM forward kernel (K1), algorithmic bytes / CUDA-event time of the
           C-ABI call, against MEASURED_PEAKS.json.  `roofline_bwd`: the backward (K2) the same way; its
           algorithmic bytes use the number of UNIQUE rows the batch touched (counted by the kernel), the
           all-rows-unique bound of SURVEY 8(d) is given next to it.
`--workload`: the other BASELINE.json configs as secondary lines (same JSON shape, `config.workload` names
           them): xdeepfm_criteo_multihot (config 5; tables sized 125 M rows per GPU), xdeepfm_ml, attention_ml.
`cpu_baseline` / `--impl reference`: the reference's own PyTorch-CPU code path (oracle/torch_port.py:
           the same ATen ops in the same order; the reference package itself cannot travel to the
           GPU box) on a bounded sample of the same workload, all host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec (fwd+bwd)"
UNIT = "samples/s"
BATCH = 65536
EMBED_DIM = 64
WORKLOAD = "deepfm_criteo_13dense_26sparse_d64_b65536_per_gpu"
K1_BYTES_PER_SAMPLE = 26 * (8 + 4 * EMBED_DIM + 4) + 13 * 4 + 4 * 39 * EMBED_DIM + 8   # 17012
K1_DRAM_TRAFFIC = 746_396_672          # bytes per launch, ncu --set full (profiles/r2_ncu_full_kernels.csv: 116.8 MB read + 629.6 MB written)
K1_TRAFFIC_SOURCE = "profiles/r2_ncu_full_kernels.csv (dram__bytes_read+write, one launch)"
CPU_SAMPLE_BATCH = 8192
CPU_SAMPLE_MAX_VOCAB = 1_000_000


WORKLOADS = {
    # name: (model, description)
    "deepfm_criteo": ("deepfm", WORKLOAD),
    "xdeepfm_criteo_multihot": ("xdeepfm", "xdeepfm_criteo_13dense_26multihot_avg8_d64_cin128x128_b{B}_per_gpu"),
    "xdeepfm_ml": ("xdeepfm", "xdeepfm_ml100k_shape_cin128x128x64_b65536"),
    "xdeepfm_ml_yaml": ("xdeepfm", "xdeepfm_ml100k_shape_cin64_b65536 (configs/xdeepfm_movielens.yaml, batch raised to fill the GPU)"),
    "xdeepfm_ml_yaml_b4096": ("xdeepfm", "xdeepfm_ml100k_shape_cin64_b4096 (configs/xdeepfm_movielens.yaml as written)"),
    "deepfm_ml_b4096": ("deepfm", "deepfm_ml100k_shape_b4096 (configs/deepfm_movielens.yaml: BASELINE config 1, the reference's CPU case)"),
    "attention_ml": ("attention_deepfm", "attention_deepfm_ml100k_shape_4heads_a64_b65536"),
}
MULTIHOT_ROWS_PER_GPU = 125_000_000     # config 5: 1 B rows over 8 GPUs (32 GB of tables per GPU)
MULTIHOT_BATCH = 16384                  # per GPU: CIN [128,128] at F = 39, D = 64 is 197 MFLOP per sample


_RESULT_FD = None      # the real stdout; fd 1 is pointed at stderr while the bench runs (see main)


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def bench_config(workload: str = "deepfm_criteo"):
    from deepfm_b200.config import ExperimentConfig
    cfg = ExperimentConfig()
    if workload in ("deepfm_criteo", "xdeepfm_criteo_multihot"):
        cfg.feature.fm_embed_dim = EMBED_DIM
    if workload == "xdeepfm_ml":
        cfg.cin.layer_sizes = [128, 128, 64]          # configs/xdeepfm_movielens_cin_tuned.yaml
    if workload.startswith("xdeepfm_ml_yaml"):
        cfg.cin.layer_sizes = [64]                    # configs/xdeepfm_movielens.yaml:22-24
    return cfg


def workload_schema(workload: str, n_gpus: int):
    from deepfm_b200 import workloads as W
    if workload == "deepfm_criteo":
        return W.criteo_schema(EMBED_DIM), BATCH
    if workload == "xdeepfm_criteo_multihot":
        scale = MULTIHOT_ROWS_PER_GPU * n_gpus / float(sum(W.CRITEO_VOCAB))
        return W.criteo_multihot_schema(EMBED_DIM, max_length=16, vocab_scale=scale), MULTIHOT_BATCH
    return W.ml100k_schema(), cfg_batch(workload)


def base_line(args, n_gpus):
    return {
        "metric": METRIC, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "model": "DeepFM", "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus,
                   "fields": "13 dense + 26 sparse", "embed_dim": EMBED_DIM, "table_rows": 33762577,
                   "dnn": [256, 128, 64], "table_grad": "row_sparse (sorted unique rows), L2 value exact",
                   "cache": "working set >> L2: 8.6 GB tables, 654 MB embeddings written per step, 4 rotating batches",
                   "parallelism": f"dp{n_gpus}" if n_gpus == 1 else
                   f"dp{n_gpus} dense params + tables <= 4096 rows replicated (NCCL allreduce, overlapped) + large tables "
                   f"row-sharded over {n_gpus} ranks (ids by NCCL all-to-all, vectors and gradients as NVLink peer-memory "
                   f"stores fused into the gather / pack kernels)"},
    }


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions -- INLINE, from the benchmark's own host thread between
    steps (`poll()` every few steps: two NVML queries, ~50 us), while the launch queue is full and the GPU is busy.
    History: `nvidia-smi -lms 100` as a subprocess (the recipe's line) stalled free-running multi-rank steps for 10-120 ms at
    a time (spikes in the per-step times at N = 2 that vanish without the sampler); an in-process NVML thread polling every
    50 ms still produced an occasional burst.  A handful of inline samples per region has not.  Falls back to one-shot
    `nvidia-smi` calls when pynvml is missing."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.rows, self.mode = index, [], None
        self.nv = self.handle = self.get_reasons = None
        self.mx = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
            except Exception:
                self.handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
            self.nv, self.mode = nv, "nvml"
            self.poll()                 # the first query of each kind costs several ms (lazy NVML state): keep it out of the timed region
            self.rows.clear()
        except Exception:
            self.mode = "nvidia-smi"

    def poll(self):
        """One sample (call it between steps inside a timed region)."""
        if self.mode == "nvml":
            try:
                sm = self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)
                r = int(self.get_reasons(self.handle))
                self.rows.append([str(sm), str(self.mx)] + ["Active" if r & b else "Not Active" for _, b in self.BITS])
            except Exception:
                pass
        elif self.mode == "nvidia-smi":
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
                if out:
                    self.rows.append([c.strip() for c in out[-1].split(",")])
            except Exception:
                pass

    def stop(self):
        if self.mode is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"], "samples": 0}
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = [n for n, _ in self.BITS]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": ("NVML queries" if self.mode == "nvml" else "nvidia-smi calls") +
                " from the host loop between steps of the timed regions"}


# ------------------------------------------------------------------------------------ CPU arm
def _import_reference():
    """The UNMODIFIED reference package from baseline/_ref (see baseline/README.md), or None."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "deepfm")):
        return None
    for p in (os.path.join(ROOT, "baseline"), ref_dir):       # baseline/dacite.py: stand-in for the absent `dacite`
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import deepfm                                          # noqa: F401
        from deepfm import config as rcfg, models as rmodels
        from deepfm.data import schema as rschema
        return rcfg, rmodels, rschema
    except Exception as e:                                     # pragma: no cover - depends on the box
        print(f"[bench] reference import failed ({e}); using the torch-CPU port", file=sys.stderr)
        return None


def _ref_schema(rschema, schema):
    """Our DatasetSchema -> the reference's own dataclasses (same fields, same order)."""
    fields = {}
    for name, f in schema.fields.items():
        kind = f.feature_type.value if hasattr(f.feature_type, "value") else str(f.feature_type)
        fields[name] = rschema.FieldSchema(name, rschema.FeatureType(kind), vocabulary_size=f.vocabulary_size,
                                           embedding_dim=f.embedding_dim, max_length=f.max_length, combiner=f.combiner)
    return rschema.DatasetSchema(fields=fields, label_field="label")


def cpu_reference_run(steps: int, warmup: int, budget_s: float = 240.0, workload: str = "deepfm_criteo"):
    """The reference's own PyTorch-CPU code path on a bounded sample of the workload, all host cores: the unmodified
    modules from baseline/_ref when they are there (kind "reference"), else the torch-CPU port (kind "port")."""
    import psutil
    import torch
    from deepfm_b200 import workloads as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model_name = WORKLOADS[workload][0]
    cfg = bench_config(workload)
    free_gb = psutil.virtual_memory().available / 2 ** 30
    note = ""
    if workload == "deepfm_criteo":
        # full size = 8.6 GB of tables + the same again of dense autograd gradients (+ zeros): needs ~40 GB of host RAM
        full = free_gb >= 48 and not os.environ.get("DFM_BENCH_CPU_SMALL")
        schema = W.criteo_schema(EMBED_DIM) if full else W.criteo_schema(EMBED_DIM, max_vocab=CPU_SAMPLE_MAX_VOCAB)
        B = BATCH if full else CPU_SAMPLE_BATCH
        note = ("full 33.76 M-row tables, the bench batch" if full else
                f"tables capped at {CPU_SAMPLE_MAX_VOCAB} rows each (host has {free_gb:.0f} GB free)")
    elif workload == "xdeepfm_criteo_multihot":
        schema = W.criteo_multihot_schema(EMBED_DIM, max_length=16, vocab_scale=CPU_SAMPLE_MAX_VOCAB / 10131227.0)
        B = 512           # the reference materialises a (B, 39*39, 64) outer product per CIN layer
        note = "tables scaled to <= 1 M rows, batch bounded by the unfused CIN outer product"
    else:
        schema = W.ml100k_schema()      # fits the host at full size
        B = cfg_batch(workload)
        note = "full ML-100K-shaped schema"
    batch = W.synthetic_batch(schema, B, seed=0)
    labels = W.synthetic_labels(B, seed=0)
    ref = None if os.environ.get("DFM_BENCH_CPU_PORT") else _import_reference()
    if ref is not None:
        rcfg, rmodels, rschema = ref
        rc = rcfg.ExperimentConfig()
        rc.feature.fm_embed_dim = cfg.feature.fm_embed_dim
        rc.cin.layer_sizes = list(cfg.cin.layer_sizes)
        torch.manual_seed(0)
        model = rmodels.create_model(model_name, _ref_schema(rschema, schema), rc)
        model.train()
        bce = torch.nn.BCEWithLogitsLoss()

        def one_step():
            model.zero_grad(set_to_none=True)
            loss = bce(model(batch).squeeze(1), labels) + model.get_l2_reg_loss()     # trainer.py:219-229
            loss.backward()
        kind = "reference"
    else:
        from oracle import torch_port as TP
        model = TP.PortedModel(model_name, schema, cfg, seed=0)

        def one_step():
            for p in model.trainable():
                p.grad = None
            model.loss(batch, labels).backward()
        kind = "port"
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    ms = 1e3 * sum(times) / len(times)
    rows = sum(f.vocabulary_size for f in schema.fields.values())
    return {"value": B / (ms / 1e3), "ms_per_step": ms, "cores": cores, "steps": len(times), "kind": kind,
            "sample": f"batch {B}, {note} ({rows} rows), {len(times)} timed steps of fwd+loss+L2+backward, "
                      f"{'unmodified reference modules (baseline/_ref)' if kind == 'reference' else 'torch-CPU port of the reference'}, "
                      f"torch {torch.__version__} CPU, {cores} threads"}


def cfg_batch(workload: str) -> int:
    return 4096 if workload.endswith("_b4096") else BATCH


def run_reference(args, rank: int, n_gpus: int):
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, workload=args.workload)
    line = base_line(args, n_gpus)
    line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                 "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                 "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "dtype": "f32"})
    if args.workload != "deepfm_criteo":
        line["config"] = {"workload": WORKLOADS[args.workload][1].format(B=cfg_batch(args.workload)), "model": WORKLOADS[args.workload][0]}
    emit(line)


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank: int, local_rank: int, n_gpus: int):
    import torch
    import torch.distributed as dist
    from deepfm_b200 import _lib, workloads as W
    from deepfm_b200.models import create_model

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the deepfm_b200 kernels have no CPU fallback")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"        # NCCL prints its version banner on STDOUT: keep stdout to the one JSON line
    _lib.lib()
    from deepfm_b200.layers.dnn import DNN as _DNN
    _DNN.fused = args.dnn_gemm == "own"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if n_gpus > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)                      # identical replicas of the dense parameters
    wl = args.workload
    model_name = WORKLOADS[wl][0]
    schema, BATCH = workload_schema(wl, n_gpus)
    cfg = bench_config(wl)
    sharded = n_gpus > 1 and wl in ("deepfm_criteo", "xdeepfm_criteo_multihot")
    if sharded:
        # tables row-sharded over the ranks (all-to-all of looked-up vectors and their gradients),
        # everything else replicated and data-parallel (one flat NCCL allreduce)
        from deepfm_b200 import models as M
        from deepfm_b200.sharded import ShardedFeatureEmbedding, TorchDistComm
        comm = TorchDistComm()
        M.BaseCTRModel.embedding_factory = staticmethod(
            lambda schema, fm_embed_dim: ShardedFeatureEmbedding(schema, fm_embed_dim, n_gpus, rank, comm))
    parity = None
    if sharded and (args.check or os.environ.get("DFM_BENCH_CHECK")):
        # parity of the product multi-rank path (routing kernels, peer-memory exchange, owner-side backward, reducer)
        # against the unsharded model, on the same schema with the vocabularies scaled down so that every rank can
        # also hold the full model; per-rank batch 8192
        from deepfm_b200.sharded import TorchDistComm as _Comm
        from deepfm_b200.sharded_check import check_against_unsharded
        cschema = (W.criteo_schema(EMBED_DIM, vocab_scale=0.02) if wl == "deepfm_criteo"
                   else W.criteo_multihot_schema(EMBED_DIM, max_length=16, vocab_scale=0.002))
        cb = 8192 if wl == "deepfm_criteo" else 1024
        parity = check_against_unsharded(model_name, cschema, cfg, W.synthetic_batch(cschema, cb, seed=7 + rank, device=dev),
                                         W.synthetic_labels(cb, seed=7 + rank, device=dev), _Comm(),
                                         replicate_below=4096 if wl == "deepfm_criteo" else 60)
        parity["what"] = (f"sharded product path vs unsharded model, {wl} schema at reduced vocabularies, {cb} samples/rank: "
                          "logits bit-identical, gradients max-norm relative error")
    with torch.device(dev):
        model = create_model(model_name, schema, cfg)
    model.train()
    model.embedding.grad_mode = "row_sparse"
    if model_name == "xdeepfm" and args.cin_precision:
        model.cin.precision = args.cin_precision
    emb = model.embedding
    if hasattr(emb, "async_sort"):
        emb.async_sort = args.sort != "inline"
    ordered = emb._ordered_params()
    table_ids = {id(p) for p, is_table in zip(ordered, emb._param_is_table) if is_table}
    dense_params = [p for p in model.parameters() if id(p) not in table_ids]
    if n_gpus > 1:      # identical replicas of the data-parallel parameters
        for p in dense_params:
            dist.broadcast(p.data, src=0)
    n_batches = 4
    host = [W.synthetic_batch(schema, BATCH, seed=100 * rank + s) for s in range(n_batches)]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    host_y = [W.synthetic_labels(BATCH, seed=100 * rank + s).pin_memory() for s in range(n_batches)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    devy = [y.to(dev) for y in host_y]
    bce = torch.nn.BCEWithLogitsLoss()

    reducer = None
    if n_gpus > 1:
        from deepfm_b200.sharded import DenseGradReducer
        emb_ids = {id(p) for p in emb.parameters()}
        early = [p for p in dense_params if id(p) not in emb_ids]
        late = [p for p in dense_params if id(p) in emb_ids]
        # allreduce of the DNN grads overlaps the table backward (launched by the sharded embedding's backward)
        reducer = DenseGradReducer(early, late, n_gpus, embedding=emb if sharded else None)

    all_params = list(model.parameters())

    def allreduce_dense():
        if reducer is not None:
            reducer.finish()

    def step(batch, labels, next_batch=None, next_ready=None):
        if next_batch is not None and sharded:
            # input pipeline: route the NEXT batch on a side stream, under this step
            model.embedding.queue_prefetch(next_batch, ready_event=next_ready)
        for p in all_params:                       # == optimizer.zero_grad(set_to_none=True): no module-tree walk per step
            p.grad = None
        logits = model(batch).squeeze(1)
        loss = bce(logits, labels) + model.get_l2_reg_loss()
        loss.backward()
        allreduce_dense()
        return loss

    def prepare(batch):
        # device-side input pipeline: the NEXT batch's row keys are emitted and sorted on a side
        # stream while this step computes (the sort depends on the ids only), so its backward starts at the segmented
        # reduction.  One sort per step still runs -- overlapped, not skipped; its duration is reported (sort_ms).
        if not sharded and args.sort == "ahead":
            model.embedding.prepare(batch)

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = None

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_step = [] if os.environ.get("DFM_BENCH_STEP_TIMES") else None
        a.record()
        every = max(steps // 4, 1)
        for i in range(steps):
            fn(i)
            if rank == 0 and sampler is not None and i % every == every - 1:
                sampler.poll()                         # clocks / throttle reasons while the region runs (GPU busy)
            if per_step is not None:
                e_ = torch.cuda.Event(enable_timing=True)
                e_.record()
                per_step.append(e_)
        b.record()
        barrier()
        if per_step is not None:
            ts = [a.elapsed_time(e_) for e_ in per_step]
            print(f"[bench rank {rank}] step end times (ms): " + " ".join(f"{t:.2f}" for t in ts) + " | deltas: " +
                  " ".join(f"{y - x:.2f}" for x, y in zip([0.0] + ts[:-1], ts)), file=sys.stderr, flush=True)
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if n_gpus > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    W_, K_ = max(args.warmup, 3), args.steps
    dbg = (lambda m: print(f"[bench rank {rank}] {m}", file=sys.stderr, flush=True)) if os.environ.get("DFM_BENCH_DEBUG") else (lambda m: None)
    dbg("model and batches ready")
    prepare(devb[0])
    for i in range(W_):
        prepare(devb[(i + 1) % n_batches])
        step(devb[i % n_batches], devy[i % n_batches], devb[(i + 1) % n_batches])
        dbg(f"warmup {i} done")
    if rank == 0 and not os.environ.get("DFM_BENCH_NO_SAMPLER"):
        sampler = ClockSampler(local_rank)
        sampler.start()
    model.embedding.profile_events = {}
    ahead = int(os.environ.get("DFM_BENCH_MAX_AHEAD", "-1"))      # >= 0: the host enqueues at most this many steps ahead
    done_events = []

    def value_step(i):
        if ahead >= 0 and len(done_events) > ahead:
            done_events.pop(0).synchronize()
        prepare(devb[(W_ + i + 1) % n_batches])
        out = step(devb[(W_ + i) % n_batches], devy[(W_ + i) % n_batches], devb[(W_ + i + 1) % n_batches])
        if ahead >= 0:
            ev_ = torch.cuda.Event()
            ev_.record()
            done_events.append(ev_)
        return out

    if os.environ.get("DFM_BENCH_CPROFILE") and rank == 0:      # development aid: where the HOST time of a step goes
        import cProfile, pstats
        pr = cProfile.Profile()
        barrier()
        pr.enable()
        for i in range(10):
            value_step(i)
        torch.cuda.synchronize()
        pr.disable()
        with open(os.environ["DFM_BENCH_CPROFILE"], "w") as fh:
            st = pstats.Stats(pr, stream=fh)
            st.sort_stats("tottime").print_stats(60)
            st.sort_stats("cumulative").print_stats(80)
    elif os.environ.get("DFM_BENCH_CPROFILE"):
        barrier()
        for i in range(10):
            value_step(i)
    total_ms = timed(value_step, K_)
    ev = model.embedding.profile_events
    model.embedding.profile_events = None
    k1_ms = k2_ms = sort_ms = None
    if ev.get("fwd"):
        k1_ms = sum(a.elapsed_time(b) for a, b in ev["fwd"]) / len(ev["fwd"])
        k2_ms = sum(a.elapsed_time(b) for a, b in ev["bwd"]) / len(ev["bwd"])
    if ev.get("sort"):
        sort_ms = sum(a.elapsed_time(b) for a, b in ev["sort"]) / len(ev["sort"])
    # K2 as ONE serial stream (sort inside the backward, nothing concurrent): the cost of the embedding backward itself.
    # In the timed steps above the sort runs on a side stream under the DNN forward, where the span between its events
    # is stretched by the kernels it shares the SMs with -- that span is reported, but is not what the sort costs.
    k2_serial_ms = None
    if k1_ms and not sharded and getattr(emb, "async_sort", False):
        emb.async_sort = False
        emb.profile_events = {}
        for i in range(6):
            step(devb[i % n_batches], devy[i % n_batches])
        torch.cuda.synchronize()
        pairs = emb.profile_events.get("bwd", [])[1:]
        if pairs:
            k2_serial_ms = sum(a.elapsed_time(b) for a, b in pairs) / len(pairs)
        emb.profile_events = None
        emb.async_sort = True

    # end to end: pinned host batch -> device copies -> step -> loss.item(), every step.  The copies of step i+1 are
    # issued on a second stream before step i is computed (a two-deep input pipeline), so they travel while the GPU
    # works; each step still waits for ITS inputs to land and reads ITS loss back on the host.
    # The batch crosses the bus as ONE packed pinned buffer (deepfm_b200.pipeline: PackedBatchLayout / DeviceStagingRing):
    # one cudaMemcpyAsync per step into a 3-deep ring of device buffers whose column views were built once.
    from deepfm_b200.pipeline import DeviceStagingRing, PackedBatchLayout
    layout = PackedBatchLayout(host[0], host_y[0])
    packed = [layout.pack(b, y) for b, y in zip(host, host_y)]
    ring = DeviceStagingRing(layout, dev, depth=3)
    main_stream = torch.cuda.current_stream(dev)

    def issue_copy(i):
        batch, labels, ev = ring.stage(i, packed[i % n_batches])
        if not sharded and args.sort == "ahead":      # the input pipeline sorts the batch's keys right behind its copy
            with torch.cuda.stream(ring.stream):
                model.embedding.prepare(batch, stream=ring.stream)
                ev = torch.cuda.Event()
                ev.record(ring.stream)
        return batch, labels, ev

    pending = [issue_copy(0)]

    def e2e_step(i):
        batch, labels, ev = pending.pop()
        main_stream.wait_event(ev)
        if sharded:                                # the next batch's ids are routed behind their copy, under this step
            nxt = issue_copy(i + 1)
            loss = step(batch, labels, nxt[0], nxt[2])
        else:                                      # enqueue this step first, then the copy of step i + 1 travels under it
            loss = step(batch, labels)
            nxt = issue_copy(i + 1)
        ring.release(i)
        pending.append(nxt)
        return loss.item()

    for i in range(2):
        e2e_step(i)
    e2e_ms = timed(lambda i: e2e_step(i + 2), K_)
    clocks = sampler.stop() if sampler is not None else None
    h2d = layout.payload_bytes

    # launch count of OUR kernels in one step (profiled outside the timed region)
    # (every rank runs the step -- it contains collectives -- only rank 0 profiles it)
    launches = None
    if rank == 0:
        from torch.profiler import ProfilerActivity, profile
        if os.environ.get("DFM_BENCH_TRACE"):     # development aid: host + device timeline of two steady-state steps
            step(devb[0], devy[0], devb[1])
            acts = [ProfilerActivity.CUDA] if os.environ.get("DFM_BENCH_TRACE_CUDA_ONLY") else [ProfilerActivity.CUDA, ProfilerActivity.CPU]
            with profile(activities=acts) as tprof:
                sync_each = bool(os.environ.get("DFM_BENCH_TRACE_E2E"))     # .item() after every step, like the e2e region
                l1 = step(devb[1], devy[1], devb[2])
                if sync_each:
                    l1.item()
                l2 = step(devb[2], devy[2], devb[3])
                if sync_each:
                    l2.item()
                else:                                  # four more: the host runs ahead, the last steps are steady state
                    for j in range(4):
                        step(devb[(3 + j) % 4], devy[(3 + j) % 4], devb[(4 + j) % 4])
                torch.cuda.synchronize()
            tprof.export_chrome_trace(os.environ["DFM_BENCH_TRACE"])
            step(devb[3], devy[3])
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(devb[0], devy[0])
            torch.cuda.synchronize()
        launches = sum(e.count for e in prof.key_averages() if "dfm::" in e.key or "DeviceRadixSort" in e.key)
        gemm_us = sum(e.device_time_total for e in prof.key_averages() if "gemm3_kernel" in e.key)
        cin_tc_us = sum(e.device_time_total for e in prof.key_averages() if "cin_tc" in e.key)
        if args.profile_step:      # per-kernel device time of one warm step (concurrent streams as they run), for profiles/
            with open(args.profile_step, "w") as fh:
                fh.write(f"# torch.profiler, one warm step of: bench.py --workload {wl} --gpus {n_gpus} --sort {args.sort}\n")
                fh.write("kernel,launches,total_us,avg_us,share_pct\n")
                evs = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
                tot = sum(e.device_time_total for e in evs) or 1.0
                for e in evs:
                    if e.device_time_total > 0:
                        fh.write(f"\"{e.key[:110]}\",{e.count},{e.device_time_total:.1f},{e.device_time_total / e.count:.2f},"
                                 f"{100.0 * e.device_time_total / tot:.2f}\n")
    else:
        if os.environ.get("DFM_BENCH_TRACE"):
            for i in range(3):
                step(devb[i], devy[i], devb[i + 1])
            if not os.environ.get("DFM_BENCH_TRACE_E2E"):
                for j in range(4):
                    step(devb[(3 + j) % 4], devy[(3 + j) % 4], devb[(4 + j) % 4])
            step(devb[3], devy[3])
        step(devb[0], devy[0])
        torch.cuda.synchronize()

    if rank != 0:
        if n_gpus > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # K2 algorithmic bytes with the MEASURED unique-row count U of the last batch:
    #   g_flat read once 4T B/sample, sorted key + payload read by the reduction 8 B/slot, per unique row one
    #   table-row read (FM fold + L2) and one gradient-row write (4d each) + the two first-order scalars;
    #   the radix sort's own passes (CUB, 4 x (read + write) of 8 B/slot) are reported separately.
    n_slots, n_unique = BATCH * 26, None
    rg = getattr(model.embedding, "row_grads", None)
    if rg is not None:
        n_slots, n_unique = [int(v) for v in rg.counts.tolist()]
    T_ = schema.total_embedding_dim
    k2_bytes_bound = 33_500 * BATCH                               # all rows unique
    k2_bytes = BATCH * 4 * T_ + n_slots * 8 + (n_unique if n_unique is not None else n_slots) * (8 * EMBED_DIM + 8)
    k2_sort_bytes = n_slots * 8 * 2 * 4
    ms_step = total_ms / K_
    line = base_line(args, n_gpus)
    from deepfm_b200 import fp32_emulation
    if args.dnn_gemm == "own":
        line["config"]["dnn_gemm"] = ("own kernels: dfm_gemm3 (tcgen05.mma kind::tf32, 3xTF32 split, fp32 in / fp32 out, TMA operands, "
                                      "TMEM accumulators promoted to fp32 registers every 128 k) + fused BatchNorm/ReLU/dropout passes")
    elif fp32_emulation._state["enabled"]:
        line["config"]["dnn_gemm"] = (f"library: {fp32_emulation.status()}, cuBLAS {fp32_emulation.cublas_version()}; "
                                      "fp32 in / fp32 out, max-norm rel err vs fp64 4.4e-7 (native SIMT sgemm: 1.8e-6)")
    else:
        line["config"]["dnn_gemm"] = f"library: torch-bundled cuBLAS {fp32_emulation.cublas_version()} SIMT sgemm"
    dnn_note = line["config"]["dnn_gemm"]
    if wl != "deepfm_criteo":
        line["config"] = {"workload": WORKLOADS[wl][1].format(B=BATCH), "model": model_name, "batch_per_gpu": BATCH,
                          "global_batch": BATCH * n_gpus, "embed_dim": cfg.feature.fm_embed_dim,
                          "table_rows": int(sum(f.vocabulary_size for f in schema.fields.values())),
                          "cin": getattr(model, "cin", None) and model.cin.layer_sizes,
                          "cin_precision": getattr(model, "cin", None) and model.cin.precision,
                          "parallelism": line["config"]["parallelism"] if sharded else f"dp{n_gpus}",
                          "dnn_gemm": dnn_note,
                          "note": "secondary line: not the workload BASELINE.json's metric is quoted on"}
    line.update({
        "value": BATCH * n_gpus / (ms_step * 1e-3), "ms_per_step": ms_step, "warmup": W_,
        "clocks": clocks,
        "e2e": {"value": BATCH * n_gpus / (e2e_ms / K_ * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K_,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": (launches or 0) * K_,
    })
    if k1_ms and wl == "deepfm_criteo":
        k1_bytes = K1_BYTES_PER_SAMPLE * BATCH
        k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
        line["roofline"] = {"kernel": "dfm::embed_fwd_kernel<4> (K1: gather+pool+FM forward)", "bound": "hbm",
                            "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                            "traffic": K1_DRAM_TRAFFIC, "ms": k1_ms, "algorithmic_bytes": k1_bytes,
                            # the same launch on the bytes that actually crossed HBM (hot rows are L2 hits)
                            "frac_on_dram_traffic": K1_DRAM_TRAFFIC / (k1_ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": peak_src,
                            "traffic_source": K1_TRAFFIC_SOURCE}
        # K2 = everything the embedding backward costs per step: the key sort (run ahead by the input pipeline on a
        # side stream, timed there) + the in-backward kernels (segmented reduction, stitch, DENSE-field streams).
        k2_total = k2_serial_ms if k2_serial_ms else k2_ms + (sort_ms or 0.0)
        k2_gbs = k2_bytes / (k2_total * 1e-3) / 1e9
        line["roofline_bwd"] = {"kernel": "K2 = key sort + seg2 segmented reduction + stitch + dense_stream (dfm_embed_bwd)",
                                "bound": "hbm", "achieved": k2_gbs, "peak": hbm_peak, "unit": "GB/s",
                                "frac": k2_gbs / hbm_peak, "ms": k2_total,
                                "ms_definition": "sort + reduction as one serial stream, CUDA events around dfm_embed_bwd "
                                                 "(6 extra steps after the timed region with the sort inside the backward)",
                                "ms_in_backward_timed_steps": k2_ms, "ms_sort_span_on_side_stream_timed_steps": sort_ms,
                                "frac_in_backward_only": k2_bytes / (k2_ms * 1e-3) / 1e9 / hbm_peak,
                                "algorithmic_bytes": k2_bytes, "id_slots": n_slots, "unique_rows": n_unique,
                                "sort_bytes_not_counted": k2_sort_bytes,
                                "all_rows_unique_bound": {"algorithmic_bytes": k2_bytes_bound,
                                                          "frac": k2_bytes_bound / (k2_total * 1e-3) / 1e9 / hbm_peak}}
        path_ms = k1_ms + k2_total
        line["roofline_path"] = {"what": "embedding + FM path = K1 + K2 (north_star target: >= 0.70 of HBM peak)", "bound": "hbm",
                                 "achieved": (k1_bytes + k2_bytes) / (path_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": (k1_bytes + k2_bytes) / (path_ms * 1e-3) / 1e9 / hbm_peak, "ms": path_ms,
                                 "algorithmic_bytes": k1_bytes + k2_bytes}
    else:
        line["roofline"] = {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                            "traffic": None, "note": "per-kernel roofline is reported by the N=1 run (unsharded K1)"}
        if k1_ms:
            line["roofline"]["k1_ms"], line["roofline"]["k2_ms"] = k1_ms, k2_ms
    # tensor-bound pieces, from the per-kernel device times of one warm step (torch.profiler, outside the timed regions)
    bf16_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1344.1)))
    tf32_peak = 0.5 * bf16_peak
    if args.dnn_gemm == "own" and launches and gemm_us > 0 and hasattr(model, "dnn"):
        dims, width = [], model.dnn.mlp[0].in_features
        for u in cfg.dnn.hidden_units:
            dims.append((width, u))
            width = u
        useful = 3 * 2.0 * BATCH * sum(a_ * b_ for a_, b_ in dims)              # forward + dX + dW of every Linear
        line["roofline_dnn"] = {"kernel": "dfm::g3::gemm3_kernel (tcgen05 kind::tf32, 3xTF32: 3 tensor-core products per fp32 product)",
                                "bound": "tensor", "achieved": 3 * useful / (gemm_us * 1e-6) / 1e12, "peak": tf32_peak,
                                "unit": "TFLOP/s", "frac": 3 * useful / (gemm_us * 1e-6) / 1e12 / tf32_peak,
                                "useful_fp32_tflops": useful / (gemm_us * 1e-6) / 1e12, "ms": gemm_us * 1e-3, "launches": 3 * len(dims),
                                "peak_source": "tf32 dense = 1/2 of the measured sustained bf16 rate (MEASURED_PEAKS.json)"}
    if launches and cin_tc_us > 0 and getattr(model, "cin", None) is not None:
        F_, D_ = schema.num_fields, cfg.feature.fm_embed_dim
        ks, prev = [], F_
        for i_, L_ in enumerate(model.cin.layer_sizes):
            ks.append(L_ * prev * F_)
            prev = model.cin.next_sizes[i_]
        cin_flops = 6.0 * D_ * sum(ks) * BATCH                                   # fwd + bwd = 6 D sum(L_i K_i) per sample
        line["roofline_cin"] = {"kernel": "cin_tc_fwd / cin_tc_bwd_data / cin_tc_dw (tcgen05 kind::tf32)", "bound": "tensor",
                                "achieved": cin_flops / (cin_tc_us * 1e-6) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                                "frac": cin_flops / (cin_tc_us * 1e-6) / 1e12 / tf32_peak, "ms": cin_tc_us * 1e-3,
                                "peak_source": "tf32 dense = 1/2 of the measured sustained bf16 rate (MEASURED_PEAKS.json)"}
    if parity is not None:
        line["parity"] = parity
    if n_gpus == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=3, warmup=1, budget_s=60.0, workload=wl)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    emit(line)
    if n_gpus > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--check", action="store_true",
                    help="N > 1: before timing, run the real sharded path against the unsharded model on a reduced copy of "
                         "the workload (deepfm_b200/sharded_check.py) and report it as `parity` in the JSON line")
    ap.add_argument("--profile-step", default=None, help="write a per-kernel device-time table of one warm step to this file")
    ap.add_argument("--sort", default="side", choices=["side", "ahead", "inline"],
                    help="where the backward's key sort runs: on a side stream right behind K1 (default), one step ahead in "
                         "the input pipeline (FeatureEmbedding.prepare), or inside the backward")
    ap.add_argument("--workload", default="deepfm_criteo", choices=sorted(WORKLOADS))
    ap.add_argument("--dnn-gemm", default="own", choices=["own", "emulated", "native"],
                    help="DNN tower: the repo's own kernels (tcgen05 3xTF32 GEMMs + fused BatchNorm/activation/dropout, "
                         "default), or the library routes kept for comparison: cuBLAS 12.9 FP32 emulation (BF16x9) / "
                         "torch's bundled cuBLAS SIMT sgemm")
    ap.add_argument("--cin-precision", default="tf32", choices=["fp32", "tf32"],
                    help="xDeepFM workloads: CIN contraction on tcgen05 (tf32) or CUDA cores (fp32)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = world if world > 1 else 1
    if args.gpus != n_gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it (the driver normally does this itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--impl", args.impl,
               "--workload", args.workload, "--cin-precision", args.cin_precision, "--dnn-gemm", args.dnn_gemm,
               "--sort", args.sort] + (["--check"] if args.check else [])
        sys.exit(subprocess.call(cmd))
    # Libraries (NCCL's version banner, for one) write to fd 1: point it at stderr for the whole run and keep the
    # real stdout for the single JSON line.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl != "reference" and args.dnn_gemm == "emulated":
        from deepfm_b200 import fp32_emulation          # before anything imports torch
        fp32_emulation.enable()
    if args.impl == "reference":
        run_reference(args, rank, n_gpus)
    else:
        run_ours(args, rank, local_rank, n_gpus)


if __name__ == "__main__":
    main()
