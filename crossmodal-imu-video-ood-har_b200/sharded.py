"""Multi-GPU exchange steps of the path: one process per GPU, windows sharded by rank.

``ShardedSimilarity`` -- the contrastive similarity matrix / sigmoid loss of the FULL batch (the reference computes it on the
outputs ``nn.DataParallel`` gathered on GPU 0: main.py:89-95, src/train/trainer.py:135-136, src/models/losses.py:37-52) with
the IMU rows sharded by rank and the video embeddings of every rank needed by every rank:

* transport ``"peer"`` (the product path): every rank's projection-head kernel writes its video-embedding operand image into a
  ``PeerBuffer`` the other ranks have mapped (CUDA IPC over NVLink / NVSwitch).  After one barrier kernel, each rank's tensor-core
  similarity kernel streams the B tiles of ALL ranks straight out of the peers' HBM through its cp.async.bulk ring -- the
  all-gather happens inside the GEMM, tile by tile, and no gathered copy exists.  The kernel's last CTA stores the rank's partial
  loss into slot[rank] of every rank; a second barrier kernel adds the slots in rank order (identical bits on every rank).
* transport ``"nccl"`` (the baseline it is measured against): ``all_gather_into_tensor`` of the 2 MiB operand image, the same
  similarity kernel on the gathered copy, ``all_reduce`` of the scalar.

``PeerBuffer`` -- a ``cudaMalloc``'ed, zero-filled buffer of this rank mapped into every rank of the process group.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _native as N
from .losses import similarity_img_native, similarity_img_work

__all__ = ["PeerBuffer", "ShardedSimilarity"]

CHUNK = 16384


class _RawCuda:
    """``__cuda_array_interface__`` holder: lets torch view memory this library allocated (no copy, no ownership)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerBuffer:
    """``nbytes`` of device memory on this rank, readable and writable by every rank of ``group`` through ``ptrs[r]``.
    Collective: every rank must construct it at the same point.  World size 1 (or no process group) degenerates to a local
    buffer."""

    def __init__(self, nbytes: int, device, group=None):
        self.device = torch.device(device)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 8:
            raise ValueError("PeerBuffer spans the GPUs of one NVSwitch box (<= 8 ranks)")
        self.nbytes = int(nbytes)
        lib = N.lib()
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(lib.cmhar_peer_alloc(self.nbytes, C.byref(p)))
            self.local = int(p.value)
            handle = (C.c_ubyte * 64)()
            N.check(lib.cmhar_peer_export(self.local, handle))
            handles: List[Optional[bytes]] = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, bytes(handle), group=group)
            self.ptrs: List[int] = []
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs.append(self.local)
                else:
                    q = C.c_void_p()
                    N.check(lib.cmhar_peer_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(q)))
                    self.ptrs.append(int(q.value))
        self._group = group
        self._closed = False

    def tensor(self, offset: int = 0, nbytes: Optional[int] = None) -> torch.Tensor:
        """uint8 view of the LOCAL buffer (torch does not own it: keep this object alive)."""
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        return torch.as_tensor(_RawCuda(self.local + offset, nbytes), device=self.device)

    def close(self) -> None:
        """Collective: unmaps the peers' buffers, then frees the local one once every rank has unmapped it."""
        if self._closed:
            return
        self._closed = True
        lib = N.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for r, q in enumerate(self.ptrs):
                if r != self.rank:
                    lib.cmhar_peer_close(q)
            if self.world > 1:
                dist.barrier(group=self._group)
            lib.cmhar_peer_free(self.local)


class ShardedSimilarity:
    """Full-batch sigmoid contrastive loss with ``rows`` windows per rank (``rows`` % 128 == 0 when world > 1).

    Buffer layout per rank: [flags: 8 x u64][slots: 8 x f64][epoch u64 ...] in the first 1 KiB, then the rank's video
    embedding operand image ([rows/128][dim/64] chunks of 16 KiB).  ``video_image_ptr`` is where this rank's projection-head
    kernel must write its image (``ProjectionHead.forward_fused(..., img_out=...)``)."""

    def __init__(self, rows: int, dim: int, device, sigmoid=(10.0, -10.0), group=None, transport: str = "peer"):
        if transport not in ("peer", "nccl"):
            raise ValueError(f"unknown transport {transport!r}")
        self.rows, self.dim, self.sigmoid, self.transport, self.group = int(rows), int(dim), sigmoid, transport, group
        self.device = torch.device(device)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1 and self.rows % 128:
            raise ValueError("sharded similarity needs rows per rank % 128 == 0 (whole operand-image tiles per rank)")
        self.img_bytes = N.lib().cmhar_operand_image_bytes(self.rows, self.dim)
        self.buf = PeerBuffer(1024 + self.img_bytes, self.device, group)
        self.video_image_ptr = self.buf.local + 1024
        self.total = self.rows * self.world
        self.work = similarity_img_work(self.rows, self.total, self.device)
        self.loss = torch.zeros((), dtype=torch.float64, device=self.device)
        if transport == "nccl":
            self.gathered = torch.empty(self.world * self.img_bytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self.gathered.data_ptr()) % 1024
            self.gathered = self.gathered[off:off + self.world * self.img_bytes]
            self.partial = torch.zeros((), dtype=torch.float64, device=self.device)

    def video_image(self) -> torch.Tensor:
        return self.buf.tensor(1024, self.img_bytes)

    def _barrier(self, with_sum: bool) -> None:
        lib = N.lib()
        b = self.buf
        with torch.cuda.device(self.device):
            N.check(lib.cmhar_peer_barrier(N.ptr_array(b.ptrs), self.rank, self.world, b.local + 128,
                                           (b.local + 64) if with_sum else None, self.world if with_sum else 0, 1.0,
                                           self.loss.data_ptr() if with_sum else None, N.stream_ptr(self.device)))

    @torch.no_grad()
    def __call__(self, imu_img: torch.Tensor) -> torch.Tensor:
        """``imu_img``: this rank's IMU embedding operand image (``rows`` x ``dim``); this rank's video image must already be in
        ``video_image_ptr`` (same stream).  Returns the full-batch mean loss (0-dim float64, identical on every rank)."""
        scale = 1.0 / (float(self.total) * float(self.total))
        if self.transport == "peer":
            self._barrier(False)                                   # every rank's image is complete
            similarity_img_native(imu_img, self.rows, [p + 1024 for p in self.buf.ptrs], self.total, self.dim, sigmoid=self.sigmoid,
                                  rows_per_part=self.rows, out_scale=scale, dst_ptrs=[p + 64 + 8 * self.rank for p in self.buf.ptrs],
                                  work=self.work)
            self._barrier(True)                                    # every rank's slot has landed; sum in rank order
            return self.loss
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.video_image(), group=self.group)
            src = self.gathered
        else:
            src = self.video_image()
        similarity_img_native(imu_img, self.rows, src, self.total, self.dim, sigmoid=self.sigmoid, out_scale=scale,
                              work=self.work, loss=self.partial)
        self.loss.copy_(self.partial)
        if self.world > 1:
            dist.all_reduce(self.loss, group=self.group)
        return self.loss

    def close(self) -> None:
        self.buf.close()
