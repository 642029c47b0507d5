"""Parity of the REAL multi-rank sharded path against the unsharded module.

Run under torchrun (``scripts/check_sharded.py``) or through ``bench.py --check``: every rank holds the same full model
(same seed), shards it with ``load_from_full``, runs the product path -- ``_ShardedEmbedFn`` with its routing kernels,
the peer-memory (or NCCL) exchange, the owner-side backward, ``DenseGradReducer`` -- on ITS batch, and compares with
the unsharded model:

  logits      of the rank's samples: bit-identical (the forward is per-sample; BatchNorm uses its running statistics
              and dropout is off in this check, so the unsharded model on the same samples is the same arithmetic)
  table grads of the rows this rank owns: the unsharded model's row-sparse gradient of the GLOBAL-mean loss on the
              concatenated batch (all ranks' samples), restricted to those rows -- max-norm relative error <= 2e-5
              (same terms, summed per (sender, row) first and then over senders instead of in one pass)
  data-parallel grads (DNN / CIN / head / DENSE-field Linears / replicated small tables) after the reducer:
              the same global-mean-loss gradients, <= 2e-5.
"""

from __future__ import annotations

import os

from typing import Dict

import torch
import torch.distributed as dist


def _rel(a: torch.Tensor, b: torch.Tensor, floor: float = 1e-7) -> float:
    if a.numel() == 0:
        return 0.0
    return float(((a.double() - b.double()).abs().max() / (b.double().abs().max() + floor)).item())


def check_against_unsharded(model_name: str, schema, cfg, batch: Dict[str, torch.Tensor], labels: torch.Tensor, comm,
                            replicate_below: int = 4096, cin_precision: str = "fp32") -> Dict[str, float]:
    """Collective.  ``batch`` / ``labels``: this rank's samples on its CUDA device.  Returns the error summary."""
    from . import models as M
    from .sharded import DenseGradReducer, ShardedFeatureEmbedding
    world, rank = comm.world, comm.rank
    dev = labels.device
    lam = float(cfg.feature.embedding_l2_reg)
    prev_factory = M.BaseCTRModel.embedding_factory
    try:
        M.BaseCTRModel.embedding_factory = None
        torch.manual_seed(4321)
        with torch.device(dev):
            full = M.create_model(model_name, schema, cfg)
        M.BaseCTRModel.embedding_factory = staticmethod(
            lambda s, fm_embed_dim: ShardedFeatureEmbedding(s, fm_embed_dim, world, rank, comm, replicate_below=replicate_below))
        with torch.device(dev):
            shard = M.create_model(model_name, schema, cfg)
    finally:
        M.BaseCTRModel.embedding_factory = prev_factory
    for p in full.parameters():                      # identical replicas of the full model on every rank
        dist.broadcast(p.data, src=0)
    with torch.no_grad():                            # non-trivial padding rows / biases
        for p in full.embedding.parameters():
            p.add_(0.01 * torch.sin(torch.arange(p.numel(), device=dev, dtype=torch.float32)).view_as(p))
    shard.embedding.load_from_full(full.embedding)
    sd = {k: v for k, v in full.state_dict().items() if not k.startswith("embedding.")}
    shard.load_state_dict(sd, strict=False)
    if model_name == "xdeepfm":
        full.cin.precision = shard.cin.precision = cin_precision
    full.eval()
    shard.eval()                                     # running-statistics BatchNorm, no dropout: per-sample arithmetic
    full.embedding.grad_mode = "row_sparse"
    shard.embedding.grad_mode = "row_sparse"
    emb = shard.embedding
    ordered = emb._ordered_params()
    table_ids = {id(p) for p, t in zip(ordered, emb._param_is_table) if t}
    dense_params = [p for p in shard.parameters() if id(p) not in table_ids]
    emb_ids = {id(p) for p in emb.parameters()}
    reducer = DenseGradReducer([p for p in dense_params if id(p) not in emb_ids], [p for p in dense_params if id(p) in emb_ids], world,
                               embedding=emb)       # the product wiring: launched from the sharded backward, flat late bucket
    bce = torch.nn.BCEWithLogitsLoss()
    out: Dict[str, float] = {}

    # ---- the product path on this rank's batch: once cold (routing inside the forward, ids by NCCL all-to-all), then the
    # way the training loop runs it -- the batch routed ahead by prefetch() (ids exchanged as peer stores when the
    # peer-memory buffers exist); the second pass is the one compared below, the first must agree with it bit for bit
    shard.zero_grad(set_to_none=True)
    logits0 = shard(batch).squeeze(1)
    (bce(logits0, labels) + shard.get_l2_reg_loss()).backward()
    reducer.finish()
    shard.zero_grad(set_to_none=True)
    emb.prefetch(batch)
    logits = shard(batch).squeeze(1)
    (bce(logits, labels) + shard.get_l2_reg_loss()).backward()
    reducer.finish()
    out["prefetched_equals_cold"] = float(torch.equal(logits.detach(), logits0.detach()))
    # ---- unsharded model: this rank's samples (logits) ...
    with torch.no_grad():
        ref_logits = full(batch).squeeze(1)
    out["logits_bit_identical"] = float(torch.equal(logits.detach(), ref_logits))
    out["logits_max_rel_err"] = _rel(logits.detach(), ref_logits)
    # ---- ... and the concatenated batch (gradients of the global-mean loss)
    whole = {}
    for k, v in batch.items():
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v.contiguous())
        whole[k] = torch.cat(parts)
    lparts = [torch.empty_like(labels) for _ in range(world)]
    dist.all_gather(lparts, labels.contiguous())
    full.zero_grad(set_to_none=True)
    (bce(full(whole).squeeze(1), torch.cat(lparts)) + full.get_l2_reg_loss()).backward()
    # table gradients of the rows this rank owns
    mine = emb.row_grads.per_table() if emb.row_grads is not None else {}
    ref = full.embedding.row_grads.per_table()
    worst_tab, n_rows = 0.0, 0
    names = emb.field_names
    for f in emb._table_idx:
        name = names[f]
        rows_r, g2_r, g1_r = ref.get(name, (torch.empty(0, dtype=torch.long, device=dev),) * 3)
        own = (rows_r + f) % world == rank
        gid = rows_r[own]
        want2, want1 = g2_r[own], g1_r[own]
        if gid.numel() and int(gid[0].item()) == 0:          # the padding row takes no lookup gradient on either side;
            gid, want2, want1 = gid[1:], want2[1:], want1[1:]    # the unsharded row-sparse list carries its L2 term only
        got = mine.get(name)
        if got is None:
            assert gid.numel() == 0, f"rank {rank}: table {name} produced no gradient rows"
            continue
        lrows, g2, g1 = got
        order = torch.argsort(lrows)
        lrows, g2, g1 = lrows[order], g2[order], g1[order]
        keep = lrows * world + ((rank - f) % world) != 0
        lrows, g2, g1 = lrows[keep], g2[keep], g1[keep]
        assert torch.equal(lrows, gid // world), f"rank {rank}: table {name}: touched rows differ"
        worst_tab = max(worst_tab, _rel(g2, want2), _rel(g1, want1))
        n_rows += int(lrows.numel())
    out["table_grad_max_rel_err"] = worst_tab
    out["table_rows_checked"] = float(n_rows)
    # data-parallel parameters
    worst_dense = 0.0
    fparams = dict(full.named_parameters())
    for k, p in shard.named_parameters():
        if id(p) in table_ids or p.grad is None:
            continue
        fg = fparams[k].grad
        if fg is None:      # a replicated small table: dense (V, d) gradient here, touched rows only in the unsharded
            parts = k.split(".")                                   # embedding.<view>_embeddings.<field>.weight
            rows_r, g2_r, g1_r = ref.get(parts[2], (None, None, None))
            if rows_r is None or rows_r.numel() == 0:
                continue
            want = g2_r if parts[1].startswith("second") else g1_r[:, None]
            worst_dense = max(worst_dense, _rel(p.grad[rows_r], want, floor=1e-6))
            continue
        worst_dense = max(worst_dense, _rel(p.grad, fg, floor=1e-6))
    out["dense_grad_max_rel_err"] = worst_dense
    out["p2p"] = float(emb._px not in (None, False))
    # worst over ranks
    t = torch.tensor([1.0 - min(out["logits_bit_identical"], out["prefetched_equals_cold"]), out["logits_max_rel_err"],
                      out["table_grad_max_rel_err"], out["dense_grad_max_rel_err"]], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([out["table_rows_checked"]], device=dev, dtype=torch.float64)
    dist.all_reduce(n)
    return {"logits_bit_identical": bool(t[0].item() == 0.0), "logits_max_rel_err": t[1].item(),
            "table_grad_max_rel_err": t[2].item(), "dense_grad_max_rel_err": t[3].item(),
            "table_rows_checked": int(n.item()), "p2p": bool(out["p2p"]), "world": world,
            "ids_by_peer_stores": bool(out["p2p"]) and os.environ.get("DFM_SHARD_P2P_IDS", "1") != "0"}
