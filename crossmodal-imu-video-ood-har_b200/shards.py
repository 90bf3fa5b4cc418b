"""Packed IMU window shards.

The reference stores every window as its own ``.npy`` of shape (T, C) fp32 (written at
src/data/preprocessing.py:352-358, opened one file per sample at src/data/datasets.py:291-321 and transposed to
(C, T) at :139-141).  The inference path needs only channel 0, samples [0, 16*(S-1)) of each window (SURVEY.md
F4): 960 of 6 000 bytes.  A shard stores exactly those bytes contiguously, so an evaluator batch is one slice
of a memory map (no file open per window, 6.25x fewer bytes from disk and over PCIe) -- and keeps the dead
bytes in a second, optional section so the original windows can be rebuilt bit for bit.

Layout (little endian):
    [  0, 64)   header: magic "CMW1", version, n, channels, window length L, live samples, has_rest
    live        float32 (n, live)            channel 0, samples [0, live)          <- what the kernels read
    labels      int64   (n,)
    rest        float32 (n, C*L - live)      [channel 0 tail | channel 1 | ... ]   (only when has_rest)

``WindowShard.batches`` yields ``{'imu': (B, live) float32, 'label': (B,) int64}`` dicts, which
``Evaluator.predict`` accepts as they are (a 2-D ``imu`` is taken as already compacted).
"""
from __future__ import annotations

import struct
from typing import Iterable, Iterator, Optional, Sequence

import numpy as np
import torch

__all__ = ["write_shard", "pack_npy_windows", "WindowShard", "live_samples"]

MAGIC = b"CMW1"
HEADER = struct.Struct("<4sIqIIII32x")        # magic, version, n, channels, L, live, has_rest  -> 64 bytes
assert HEADER.size == 64


def live_samples(window: int, patch: int = 16, stride: int = 16, max_tokens: int = 16) -> int:
    """Samples of channel 0 the reference's encoder actually consumes: tokens kept = min(1 + C*N, N + 1)
    (src/models/models.py:122-123) -> stride*(S-2) + patch samples; 240 for L = 250, 96 for L = 100."""
    n = (window - patch) // stride + 1
    s = min(n + 1, max_tokens)
    return stride * (s - 2) + patch


def write_shard(path: str, windows: np.ndarray, labels: Optional[Sequence[int]] = None, keep_rest: bool = True,
                live: Optional[int] = None) -> None:
    """windows: (n, C, L) float32 (the reference's dataset layout)."""
    w = np.ascontiguousarray(windows, dtype=np.float32)
    if w.ndim != 3:
        raise ValueError("windows must be (n, C, L)")
    n, C, L = w.shape
    live = live_samples(L) if live is None else int(live)
    lab = np.zeros(n, np.int64) if labels is None else np.asarray(labels, np.int64).reshape(n)
    with open(path, "wb") as f:
        f.write(HEADER.pack(MAGIC, 1, n, C, L, live, int(keep_rest)))
        f.write(np.ascontiguousarray(w[:, 0, :live]).tobytes())
        f.write(lab.tobytes())
        if keep_rest:
            rest = np.concatenate([w[:, 0, live:], w[:, 1:, :].reshape(n, -1)], axis=1)
            f.write(np.ascontiguousarray(rest).tobytes())


def pack_npy_windows(npy_paths: Iterable[str], labels: Sequence[int], out_path: str, keep_rest: bool = True) -> int:
    """Packs the reference's per-window ``.npy`` files ((T, C) each) into one shard; returns the window count."""
    wins = [np.load(p).astype(np.float32).T for p in npy_paths]          # (C, T), as datasets.py:139-141
    write_shard(out_path, np.stack(wins), labels, keep_rest)
    return len(wins)


class WindowShard:
    """Memory-mapped reader."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            magic, version, n, C, L, live, has_rest = HEADER.unpack(f.read(HEADER.size))
        if magic != MAGIC or version != 1:
            raise ValueError(f"{path}: not a CMW1 window shard")
        self.n, self.channels, self.window, self.live_len, self.has_rest = n, C, L, live, bool(has_rest)
        off = HEADER.size
        self.live = np.memmap(path, np.float32, "r", off, (n, live))
        off += n * live * 4
        self.labels = np.memmap(path, np.int64, "r", off, (n,))
        off += n * 8
        self.rest = np.memmap(path, np.float32, "r", off, (n, C * L - live)) if self.has_rest else None

    def __len__(self) -> int:
        return self.n

    def window_full(self, i: int) -> np.ndarray:
        """Rebuilds window i as (C, L), bit for bit (needs the rest section)."""
        if self.rest is None:
            raise ValueError("shard was written without the dead-input section")
        C, L, live = self.channels, self.window, self.live_len
        out = np.empty((C, L), np.float32)
        out[0, :live] = self.live[i]
        out[0, live:] = self.rest[i, :L - live]
        out[1:] = self.rest[i, L - live:].reshape(C - 1, L)
        return out

    def batches(self, batch_size: int, lo: int = 0, hi: Optional[int] = None, pin: bool = False) -> Iterator[dict]:
        """Evaluator-ready batches of windows [lo, hi) (a rank's slice: ``evaluator.shard_bounds``)."""
        hi = self.n if hi is None else hi
        for s in range(lo, hi, batch_size):
            e = min(s + batch_size, hi)
            imu = torch.from_numpy(np.array(self.live[s:e]))              # copy out of the read-only map
            yield {"imu": imu.pin_memory() if pin else imu, "label": torch.from_numpy(np.array(self.labels[s:e]))}
