"""Shared helpers of the GPU parity tests."""

import numpy as np
import torch

from tests.conftest import load_golden, split_prefixed  # noqa: F401


def to_dev(arrays, dev="cuda"):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in arrays.items()}


def load_params(module, arrays, prefix=""):
    sd = {k[len(prefix):]: torch.from_numpy(np.asarray(v)) for k, v in arrays.items() if k.startswith(prefix)}
    module.load_state_dict(sd)
    return module


def grads_of(module):
    return {k: (None if p.grad is None else p.grad.detach().cpu().numpy()) for k, p in module.named_parameters()}


def assert_close_rel(a, b, tol, what="", floor=1e-7):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    err = np.abs(a - b).max()
    ref = np.abs(b).max()
    assert err <= tol * ref + floor, f"{what}: max err {err:.3e} vs ref max {ref:.3e} (tol {tol})"
