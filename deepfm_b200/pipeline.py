"""Device-side input pipeline (reference: deepfm/data/dataset.py:10-38 +
``DataLoader(train_ds, batch_size, shuffle, num_workers=0)`` in deepfm/training/trainer.py:202-217).

The reference builds every sample in Python (``TabularDataset.__getitem__``: one ``torch.tensor`` per feature per row)
and collates 4096 of them per batch -- about 5 k samples/s, three orders of magnitude below the kernels.  Here the
dataset stays COLUMNAR:

  ColumnarDataset   the same ``(features: dict[str, ndarray], labels)`` constructor as ``TabularDataset``; every column
                    becomes one tensor with the reference's dtypes (integer -> int64, float -> float32, ``(N, L)``
                    sequence columns kept 2-D), either resident on the device or in pinned host memory.
  DeviceLoader      yields ``(batch_features, batch_labels)`` exactly like the reference's DataLoader (same dict, same
                    dtypes / shapes, same ``len()``, ``drop_last=False``), but a batch is a slice (or, shuffled, one
                    ``index_select``) of each column: on-device shuffle for a device-resident dataset, pinned
                    host -> device copies issued ``prefetch`` batches ahead on a copy stream otherwise.  With an
                    ``embedding`` module it also runs that batch's ahead-of-step key sort (``FeatureEmbedding.prepare``:
                    the sort the embedding backward needs depends on the ids only) or, for the row-sharded module, its
                    routing (``ShardedFeatureEmbedding.prefetch``).
  to_csr            ``(B, L)`` zero-padded ids -> (values, offsets): the non-pad ids of every bag in bag order and the
                    exclusive prefix sum of the non-pad counts -- bit-exact against ``oracle.csr_flatten``.
"""

from __future__ import annotations

from typing import Dict, Iterator, Optional, Tuple

import numpy as np
import torch


def to_csr(ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(B, L) int64 zero-padded ids -> (values (nnz,), offsets (B + 1,)): pads skipped wherever they occur."""
    if ids.dim() != 2:
        raise ValueError(f"to_csr expects (B, L) ids, got {tuple(ids.shape)}")
    mask = ids != 0
    counts = mask.sum(dim=1)
    offsets = torch.zeros(ids.shape[0] + 1, dtype=torch.int64, device=ids.device)
    torch.cumsum(counts, 0, out=offsets[1:])
    return ids[mask], offsets


class ColumnarDataset:
    """Columnar stand-in for the reference's ``TabularDataset`` (dataset.py:10-38)."""

    def __init__(self, features: Dict[str, np.ndarray], labels: np.ndarray, device: Optional[torch.device] = None,
                 pin: bool = True) -> None:
        self._length = len(labels)
        self.features: Dict[str, torch.Tensor] = {}
        for name, values in features.items():
            t = torch.as_tensor(values)
            if len(t) != self._length:
                raise ValueError(f"feature {name!r} has {len(t)} rows, labels have {self._length}")
            # dataset.py:32-35: integer columns -> long, everything else -> float32
            t = t.to(torch.long) if not t.dtype.is_floating_point and t.dtype != torch.bool else t.to(torch.float32)
            self.features[name] = t.contiguous()
        self.labels = torch.as_tensor(labels).to(torch.float32).contiguous()     # dataset.py:37
        if device is not None and torch.device(device).type == "cuda":
            self.to(device)
        elif pin and torch.cuda.is_available():
            self.features = {k: v.pin_memory() for k, v in self.features.items()}
            self.labels = self.labels.pin_memory()

    def to(self, device) -> "ColumnarDataset":
        self.features = {k: v.to(device) for k, v in self.features.items()}
        self.labels = self.labels.to(device)
        return self

    @property
    def device(self) -> torch.device:
        return self.labels.device

    def __len__(self) -> int:
        return self._length

    def __getitem__(self, idx: int):
        """Row access with the reference's return convention (0-d / 1-d tensors); the loader never uses it."""
        return {k: v[idx] for k, v in self.features.items()}, self.labels[idx]


class DeviceLoader:
    def __init__(self, dataset: ColumnarDataset, batch_size: int, shuffle: bool = False, device=None,
                 drop_last: bool = False, prefetch: int = 2, embedding=None, seed: Optional[int] = None) -> None:
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        self.ds, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), bool(shuffle), bool(drop_last)
        self.device = torch.device(device) if device is not None else dataset.device
        self.prefetch = max(int(prefetch), 1)
        self.embedding = embedding
        self.gen = None
        if seed is not None:
            self.gen = torch.Generator(device=dataset.device if dataset.device.type == "cuda" else "cpu")
            self.gen.manual_seed(seed)
        self._copy_stream = None

    def __len__(self) -> int:
        n = len(self.ds)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    # -- one batch: slices (sequential) or one gather per column (shuffled), on whatever device the dataset lives
    def _take(self, perm: Optional[torch.Tensor], lo: int, hi: int):
        if perm is None:
            feats = {k: v[lo:hi] for k, v in self.ds.features.items()}
            return feats, self.ds.labels[lo:hi]
        idx = perm[lo:hi]
        feats = {k: v.index_select(0, idx) for k, v in self.ds.features.items()}
        return feats, self.ds.labels.index_select(0, idx)

    def _stage(self, feats, labels):
        """Host-resident dataset: pinned staging (gathers produce pageable tensors) + async copies on the copy stream."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            pin = lambda t: t if t.is_pinned() else t.pin_memory()
            d_feats = {k: pin(v).to(self.device, non_blocking=True) for k, v in feats.items()}
            d_labels = pin(labels).to(self.device, non_blocking=True)
            if self.embedding is not None and hasattr(self.embedding, "prepare"):
                self.embedding.prepare(d_feats, stream=self._copy_stream)     # the sort rides behind the copy
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        for t in list(d_feats.values()) + [d_labels]:
            t.record_stream(main)
        return d_feats, d_labels, ev

    def __iter__(self) -> Iterator[Tuple[Dict[str, torch.Tensor], torch.Tensor]]:
        n = len(self.ds)
        on_device = self.ds.device.type == "cuda"
        perm = None
        if self.shuffle:
            perm = torch.randperm(n, device=self.ds.device, generator=self.gen)
        bounds = [(lo, min(lo + self.batch_size, n)) for lo in range(0, n, self.batch_size)]
        if self.drop_last and bounds and bounds[-1][1] - bounds[-1][0] < self.batch_size:
            bounds.pop()
        if on_device or self.device.type != "cuda":
            pending = None
            for lo, hi in bounds:
                feats, labels = self._take(perm, lo, hi)
                if self.device != self.ds.device:
                    feats = {k: v.to(self.device) for k, v in feats.items()}
                    labels = labels.to(self.device)
                if self.embedding is not None and feats[next(iter(feats))].is_cuda:
                    # one batch of look-ahead: batch i + 1 is prepared (sorted / routed) before batch i is consumed
                    if hasattr(self.embedding, "prepare"):
                        self.embedding.prepare(feats)
                if pending is not None:
                    yield pending
                pending = (feats, labels)
            if pending is not None:
                yield pending
            return
        # host-resident dataset -> CUDA device: `prefetch` batches of copies in flight
        queue = []
        it = iter(bounds)
        main = torch.cuda.current_stream(self.device)

        def issue():
            try:
                lo, hi = next(it)
            except StopIteration:
                return False
            queue.append(self._stage(*self._take(perm, lo, hi)))
            return True

        for _ in range(self.prefetch):
            issue()
        while queue:
            feats, labels, ev = queue.pop(0)
            main.wait_event(ev)
            issue()
            yield feats, labels


class PackedBatchLayout:
    """Byte layout of one batch (features + labels) inside ONE contiguous buffer, every column 256-byte aligned.

    The reference's DataLoader hands the step ~40 separate tensors (dataset.py:28-38), i.e. ~40 host -> device copies
    of a few hundred KB each; packed, a batch crosses PCIe / NVLink-C2C as a single ``cudaMemcpyAsync`` and the
    columns are views of the landed buffer (same dtypes / shapes as the reference's batch dict)."""

    def __init__(self, features: Dict[str, torch.Tensor], labels: torch.Tensor) -> None:
        self.columns = []        # (name, dtype, shape, byte offset, bytes); name None = labels
        off = 0
        for name, t in list(features.items()) + [(None, labels)]:
            nbytes = t.numel() * t.element_size()
            self.columns.append((name, t.dtype, tuple(t.shape), off, nbytes))
            off = (off + nbytes + 255) // 256 * 256
        self.nbytes = max(off, 256)
        self.payload_bytes = sum(c[4] for c in self.columns)

    def views(self, buf: torch.Tensor):
        """(features dict, labels) as views of a uint8 buffer of ``nbytes`` bytes."""
        feats, labels = {}, None
        for name, dtype, shape, off, nbytes in self.columns:
            v = buf[off: off + nbytes].view(dtype).view(shape)
            if name is None:
                labels = v
            else:
                feats[name] = v
        return feats, labels

    def pack(self, features: Dict[str, torch.Tensor], labels: torch.Tensor, pin: bool = True) -> torch.Tensor:
        """A (pinned) host buffer holding this batch."""
        buf = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
        feats, lab = self.views(buf)
        for k, v in features.items():
            feats[k].copy_(v)
        lab.copy_(labels)
        return buf


class DeviceStagingRing:
    """``depth`` device buffers of one layout with their column views built once: staging batch ``i`` is one
    non-blocking copy into buffer ``i % depth`` on the copy stream (no per-column allocation, no per-column launch).
    A buffer is reused ``depth`` batches later; the caller keeps at most ``depth - 1`` batches in flight."""

    def __init__(self, layout: PackedBatchLayout, device, depth: int = 3) -> None:
        self.layout, self.depth = layout, int(depth)
        self.bufs = [torch.empty(layout.nbytes, dtype=torch.uint8, device=device) for _ in range(self.depth)]
        self.batches = [layout.views(b) for b in self.bufs]
        self.stream = torch.cuda.Stream(device=device)
        self._released = [None] * self.depth

    def release(self, i: int) -> None:
        """Call after the step that consumed batch ``i`` was enqueued: its slot is refilled only after that point."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.bufs[0].device))
        self._released[i % self.depth] = ev

    def stage(self, i: int, host_buf: torch.Tensor):
        """Issue the copy of ``host_buf`` into slot ``i % depth``; returns (features, labels, event)."""
        k = i % self.depth
        with torch.cuda.stream(self.stream):
            if self._released[k] is not None:
                self.stream.wait_event(self._released[k])
                self._released[k] = None
            self.bufs[k].copy_(host_buf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        feats, labels = self.batches[k]
        return feats, labels, ev
