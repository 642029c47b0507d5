"""Lazy out-of-range-id reporting shared by ``FeatureEmbedding`` and ``ShardedFeatureEmbedding``.

The reference raises ``IndexError`` for an id outside ``[0, vocabulary_size)`` (ATen CPU embedding).  The kernels
clamp such an id to the padding row and set a device status word; the word travels to pinned host memory without
blocking after every forward and is looked at later, so the hot path has no host sync and a bad id still raises
(at the latest two forwards later, in ``backward`` if it has arrived by then, or in ``raise_if_bad_index()``).
"""

from __future__ import annotations

import torch


class IndexStatusMixin:
    def _init_status(self) -> None:
        # True (default): lazy check; "sync": check inside forward (one host sync per call); False: never check.
        self.check_indices = True
        self._status = None             # device int32 word written by the kernels
        self._status_pending = []       # [(pinned host int32, event)] of forwards not looked at yet
        self._status_free = []

    def _status_word(self, device):
        if not self.check_indices:
            return None
        if self._status is None or self._status.device != device:
            self._status = torch.zeros(1, dtype=torch.int32, device=device)
        return self._status

    def _post_status(self) -> None:
        """After the kernels of a forward: copy the status word to pinned host memory without blocking."""
        if self._status is None:
            return
        host = self._status_free.pop() if self._status_free else torch.zeros(1, dtype=torch.int32).pin_memory()
        host.copy_(self._status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._status_pending.append((host, ev))
        if self.check_indices == "sync":
            self.raise_if_bad_index()

    def raise_if_bad_index(self, block: bool = True, keep: int = 0) -> None:
        """Raise IndexError if a forward since the last check saw an id outside [0, vocabulary_size).  block=False
        only looks at status words that have already arrived; keep=n leaves the n most recent forwards unchecked
        unless they have arrived (so a training loop never waits on the step it has just enqueued)."""
        bad = False
        while self._status_pending:
            host, ev = self._status_pending[0]
            if not ev.query():
                if not block or len(self._status_pending) <= keep:
                    break
                ev.synchronize()
            self._status_pending.pop(0)
            bad = bad or int(host[0]) != 0
            self._status_free.append(host)
        if bad:
            self._status.zero_()
            raise IndexError(f"index out of range in {type(self).__name__} (an id is outside [0, vocabulary_size))")
