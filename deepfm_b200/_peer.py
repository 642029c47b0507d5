"""The ONE place that touches ``torch.distributed._symmetric_memory`` (a private torch API).

Symmetric memory = CUDA VMM allocations that every rank of a process group maps into its own address space over
NVLink / NVSwitch; the sharded embedding's exchange kernels store straight into the peers' buffers through it
(``deepfm_b200/sharded.py: PeerExchange``).  Everything the repo needs from the API is behind these two functions, so a
torch upgrade that moves or renames it is a one-file change (``tests/test_sharded.py`` checks the shim's contract
against whatever torch is installed)."""

from __future__ import annotations

import torch


def available() -> bool:
    try:
        import torch.distributed._symmetric_memory as symm     # noqa: F401
        return hasattr(symm, "empty") and hasattr(symm, "rendezvous")
    except Exception:
        return False


def symmetric_empty(numel: int, device, group):
    """(float32 tensor of ``numel`` elements in symmetric memory, handle).  The handle exposes ``buffer_ptrs`` (the
    device address of every rank's copy) and ``barrier(channel=...)`` (a device-side all-ranks barrier)."""
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty((int(numel),), dtype=torch.float32, device=device)
    hdl = symm.rendezvous(buf, group.group_name)
    for attr in ("buffer_ptrs", "barrier"):
        if not hasattr(hdl, attr):
            raise RuntimeError(f"torch symmetric-memory handle has no {attr!r}: update deepfm_b200/_peer.py for this torch")
    return buf, hdl
