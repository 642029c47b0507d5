"""On-device evaluation (reference: deepfm/training/trainer.py:244-332 and
deepfm/training/metrics.py:9-111).

The reference copies every batch of probabilities to the host, calls sklearn for AUC / log-loss and groups the scores
per user in a Python loop (1 positive + 999 negatives per user in the leave-one-out protocol, so the loop runs over
millions of rows).  Here the scores stay on the device; every metric is a handful of sort / segment operations over
the whole evaluation set and ONE host read at the end:

  auc       Mann-Whitney U with average ranks for tied scores == sklearn.metrics.roc_auc_score (trapezoid over ties)
  logloss   mean binary cross-entropy of probabilities clipped to [1e-7, 1 - 1e-7] (metrics.py:14-18)
  HR@K, NDCG@K  per user with at least one positive and one negative (trainer.py:318-326): the user's rows are ranked
            by score (descending, ties in row order -- numpy's stable argsort; the reference's default argsort leaves tie
            order unspecified) and the FIRST positive's 1-based rank r gives hit = r <= K, gain = 1 / log2(r + 1)
            (metrics.py:96-109), averaged over those users.

The functions take torch tensors on any device (the CPU tests compare them with the reference's own formulas).
"""

from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch


def auc(labels: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
    """ROC AUC as a 0-d float64 tensor on the inputs' device (NaN when only one class is present)."""
    labels = labels.reshape(-1).to(torch.float64)
    scores = scores.reshape(-1)
    n = scores.numel()
    order = torch.argsort(scores, stable=True)
    s = scores[order]
    lab = labels[order]
    # average 1-based rank of every run of tied scores
    new_run = torch.ones(n, dtype=torch.bool, device=s.device)
    if n > 1:
        new_run[1:] = s[1:] != s[:-1]
    run_id = torch.cumsum(new_run.to(torch.int64), 0) - 1
    n_runs = int(run_id[-1].item()) + 1 if n else 0
    counts = torch.zeros(n_runs, dtype=torch.float64, device=s.device).index_add_(0, run_id, torch.ones(n, dtype=torch.float64, device=s.device))
    ends = torch.cumsum(counts, 0)
    avg_rank = ends - (counts - 1.0) / 2.0
    ranks = avg_rank[run_id]
    n_pos = lab.sum()
    n_neg = n - n_pos
    u = (ranks * lab).sum() - n_pos * (n_pos + 1.0) / 2.0
    return u / (n_pos * n_neg)


def logloss(labels: torch.Tensor, probs: torch.Tensor) -> torch.Tensor:
    p = probs.reshape(-1).to(torch.float64).clamp(1e-7, 1 - 1e-7)
    y = labels.reshape(-1).to(torch.float64)
    return -(y * torch.log(p) + (1 - y) * torch.log1p(-p)).mean()


def ranking_metrics(user_ids: torch.Tensor, labels: torch.Tensor, scores: torch.Tensor,
                    ks: Sequence[int] = (5, 10, 20)) -> Dict[str, torch.Tensor]:
    """HR@K / NDCG@K over the users that have both a positive and a negative row ({} as tensors; empty dict if none)."""
    user_ids = user_ids.reshape(-1).to(torch.int64)
    labels = labels.reshape(-1)
    scores = scores.reshape(-1)
    n = scores.numel()
    if n == 0:
        return {}
    # rows ordered by (user, score descending, row index): two stable sorts, minor key first
    o1 = torch.argsort(scores, descending=True, stable=True)
    o2 = torch.argsort(user_ids[o1], stable=True)
    order = o1[o2]
    u = user_ids[order]
    lab = labels[order] > 0
    start = torch.ones(n, dtype=torch.bool, device=u.device)
    start[1:] = u[1:] != u[:-1]
    seg = torch.cumsum(start.to(torch.int64), 0) - 1
    n_users = int(seg[-1].item()) + 1
    idx = torch.arange(n, device=u.device)
    seg_start = torch.zeros(n_users, dtype=torch.int64, device=u.device).scatter_reduce_(0, seg, idx, "amin", include_self=False)
    pos_in_user = idx - seg_start[seg]                      # 0-based rank inside the user's list
    big = torch.full((n_users,), n, dtype=torch.int64, device=u.device)
    first_pos = big.scatter_reduce(0, seg[lab], pos_in_user[lab], "amin", include_self=True)
    n_rows = torch.zeros(n_users, dtype=torch.int64, device=u.device).index_add_(0, seg, torch.ones(n, dtype=torch.int64, device=u.device))
    n_pos = torch.zeros(n_users, dtype=torch.int64, device=u.device).index_add_(0, seg, lab.to(torch.int64))
    keep = (n_pos > 0) & (n_pos < n_rows)
    n_eval = keep.sum()
    if int(n_eval.item()) == 0:
        return {}
    rank = (first_pos[keep] + 1).to(torch.float64)          # 1-based rank of the first positive
    out: Dict[str, torch.Tensor] = {}
    for k in ks:
        hit = rank <= k
        out[f"HR@{k}"] = hit.to(torch.float64).sum() / n_eval
        out[f"NDCG@{k}"] = torch.where(hit, 1.0 / torch.log2(rank + 1.0), torch.zeros_like(rank)).sum() / n_eval
    return out


@torch.no_grad()
def evaluate_on_device(model, batches: Iterable[Tuple[Dict[str, torch.Tensor], torch.Tensor]],
                       user_field: Optional[str] = "user_id", ks: Sequence[int] = (5, 10, 20)) -> Dict[str, float]:
    """The reference's ``Trainer.evaluate`` (trainer.py:244-294) with the scores kept on the device: ``model.predict``
    per batch, then auc / logloss / HR@K / NDCG@K in one go and a single host read."""
    was_training = model.training
    model.eval()
    scores: List[torch.Tensor] = []
    labels: List[torch.Tensor] = []
    users: List[torch.Tensor] = []
    for feats, y in batches:
        scores.append(model.predict(feats).squeeze(1))
        labels.append(y.to(scores[-1].device))
        if user_field is not None and user_field in feats:
            users.append(feats[user_field])
    model.train(was_training)
    s, y = torch.cat(scores), torch.cat(labels)
    res: Dict[str, torch.Tensor] = {"auc": auc(y, s), "logloss": logloss(y, s)}
    if users:
        res.update(ranking_metrics(torch.cat(users), y, s, ks))
    keys = list(res)
    vals = torch.stack([res[k].to(torch.float64) for k in keys]).cpu().tolist()     # the one host read
    out = dict(zip(keys, vals))
    if math.isnan(out["auc"]):
        out["auc"] = 0.0                                                            # trainer.py:277-280
    return out
