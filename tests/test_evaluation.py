"""On-device evaluation metrics (deepfm_b200/evaluation.py) against the reference's own formulas
(deepfm/training/metrics.py:9-111, trainer.py:296-332) restated with numpy / sklearn.  Runs on CPU tensors; the
same torch ops run on the GPU in evaluate_on_device."""

import numpy as np
import pytest
import torch

from deepfm_b200 import evaluation as E


def _ref_ranking(user_ids, labels, scores, ks):
    """trainer.py:307-332 + metrics.py:84-111 (stable argsort: ties in row order)."""
    groups = {}
    for i, u in enumerate(user_ids):
        groups.setdefault(int(u), []).append(i)
    ranks = []
    for u, idx in groups.items():
        ul, us = labels[idx], scores[idx]
        if ul.sum() > 0 and ul.sum() < len(ul):
            r = ul[np.argsort(-us, kind="stable")]
            ranks.append(r)
    out = {}
    for k in ks:
        out[f"HR@{k}"] = sum(1 for r in ranks if 1 in r[:k]) / len(ranks)
        out[f"NDCG@{k}"] = sum(1.0 / np.log2(np.where(r[:k] == 1)[0][0] + 2) for r in ranks if 1 in r[:k]) / len(ranks)
    return out


@pytest.mark.parametrize("ties", [False, True])
def test_auc_and_logloss_match_sklearn(ties):
    from sklearn.metrics import log_loss, roc_auc_score
    rng = np.random.default_rng(3)
    y = (rng.random(5000) < 0.2).astype(np.float32)
    p = rng.random(5000).astype(np.float32)
    if ties:
        p = np.round(p, 2)                 # many tied scores
    got = E.auc(torch.from_numpy(y), torch.from_numpy(p)).item()
    assert abs(got - roc_auc_score(y, p)) < 1e-12
    got = E.logloss(torch.from_numpy(y), torch.from_numpy(p)).item()
    assert abs(got - log_loss(y, np.clip(p.astype(np.float64), 1e-7, 1 - 1e-7))) < 1e-9


def test_ranking_metrics_match_reference_loops():
    rng = np.random.default_rng(5)
    n_users, per = 200, 50
    users = np.repeat(np.arange(1, n_users + 1), per)
    labels = np.zeros(n_users * per, dtype=np.float32)
    labels[np.arange(n_users) * per + rng.integers(0, per, n_users)] = 1      # one positive per user
    labels[:per] = 0                                                          # a user without positives: skipped
    labels[per:2 * per] = 1                                                   # a user without negatives: skipped
    labels[2 * per: 2 * per + 3] = 1                                          # several positives: the first one counts
    scores = np.round(rng.random(n_users * per), 1).astype(np.float32)        # ties
    perm = rng.permutation(len(users))                                        # rows arrive in any order
    users, labels, scores = users[perm], labels[perm], scores[perm]
    ks = (5, 10, 20)
    got = E.ranking_metrics(torch.from_numpy(users), torch.from_numpy(labels), torch.from_numpy(scores), ks)
    want = _ref_ranking(users, labels, scores, ks)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k].item() - want[k]) < 1e-12, k


def test_ranking_metrics_empty_when_no_user_qualifies():
    u = torch.tensor([1, 1, 2, 2])
    assert E.ranking_metrics(u, torch.zeros(4), torch.rand(4)) == {}
