"""Drop-in surface of the reference's ``src/models/losses.py`` contrastive losses.

Inference / validation (gradients disabled): one fused kernel computes the similarity matrix tile
by tile and reduces it in registers -- the B x B matrix never reaches HBM.  Training (gradients
enabled): the same algebra in differentiable torch ops (the reference's trainers back-propagate
through the loss, ``src/train/trainer.py:136-139``).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from .models import _prec_code

__all__ = ["SigmoidContrastiveLoss", "InfoNCELoss", "similarity_native", "similarity_img_native"]


def similarity_native(a: torch.Tensor, b: torch.Tensor, *, materialize: bool = False,
                      sigmoid: Optional[tuple] = None, lse_scale: Optional[float] = None,
                      diag_offset: int = 0, precision: Optional[str] = None) -> dict:
    """``cmhar_similarity`` wrapper.  Returns a dict with any of: 'sim' (na,nb), 'sigmoid_sum'
    (0-dim float64: sum softplus(-(s*scale+bias))), 'row_lse' (na), 'col_lse' (nb), 'diag'."""
    N.require_cuda(a, "similarity")
    a, b = N.f32c(a), N.f32c(b)
    na, nb, dim = a.shape[0], b.shape[0], a.shape[1]
    dev = a.device
    lib = N.lib()
    out = {}
    sim = torch.empty((na, nb), dtype=torch.float32, device=dev) if materialize else None
    ssum = torch.zeros((), dtype=torch.float64, device=dev) if sigmoid is not None else None
    row = col = diag = work = None
    if lse_scale is not None:
        row = torch.empty(na, dtype=torch.float32, device=dev)
        col = torch.empty(nb, dtype=torch.float32, device=dev)
        diag = torch.zeros(na, dtype=torch.float32, device=dev)
    if lse_scale is not None or _prec_code(precision) == N.BF16:     # the bf16 tensor-core path stages operand images there
        work = torch.empty(lib.cmhar_similarity_work_bytes(na, nb, dim), dtype=torch.uint8, device=dev)
    sc, sb = sigmoid if sigmoid is not None else (1.0, 0.0)
    with torch.cuda.device(dev):
        N.check(lib.cmhar_similarity(a.data_ptr(), b.data_ptr(), na, nb, dim, diag_offset, N.ptr(sim),
                                     float(sc), float(sb), N.ptr(ssum), float(lse_scale or 1.0), N.ptr(row),
                                     N.ptr(col), N.ptr(diag), N.ptr(work), _prec_code(precision), N.stream_ptr(dev)))
    for k, v in (("sim", sim), ("sigmoid_sum", ssum), ("row_lse", row), ("col_lse", col), ("diag", diag)):
        if v is not None:
            out[k] = v
    return out


def similarity_img_native(a_img: torch.Tensor, na: int, b_imgs, nb: int, dim: int, *, sigmoid=(10.0, -10.0),
                          rows_per_part: int = 0, out_scale: Optional[float] = None, dst_ptrs=None,
                          work: Optional[torch.Tensor] = None, loss: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``cmhar_similarity_img``: the fused sigmoid contrastive loss from operands that already are bf16 operand images.
    ``b_imgs``: one image (tensor) or a list of device pointers / tensors, one per shard of ``rows_per_part`` rows (other
    ranks' shards are peer-mapped pointers read over NVLink inside the GEMM).  The result, ``out_scale`` (default 1/(na nb))
    times the softplus sum of this (na x nb) block, is stored to ``loss`` (0-dim float64, returned) or to ``dst_ptrs``.
    ``work`` must have been zeroed once (``similarity_img_work``); it is re-armed by the kernel."""
    dev = a_img.device
    parts = b_imgs if isinstance(b_imgs, (list, tuple)) else [b_imgs]
    ptrs = [q.data_ptr() if isinstance(q, torch.Tensor) else int(q) for q in parts]
    if work is None:
        work = similarity_img_work(na, nb, dev)
    if dst_ptrs is None:
        if loss is None:
            loss = torch.empty((), dtype=torch.float64, device=dev)
        dst_ptrs = [loss.data_ptr()]
    sc, sb = sigmoid
    scale = float(out_scale) if out_scale is not None else 1.0 / (float(na) * float(nb))
    with torch.cuda.device(dev):
        N.check(N.lib().cmhar_similarity_img(a_img.data_ptr(), na, N.ptr_array(ptrs), len(ptrs), int(rows_per_part), nb, dim,
                                             float(sc), float(sb), scale, N.ptr_array(dst_ptrs), len(dst_ptrs), work.data_ptr(),
                                             N.stream_ptr(dev)))
    return loss


def similarity_img_work(na: int, nb: int, device) -> torch.Tensor:
    """Zeroed workspace (ticket + per-CTA partials) of ``cmhar_similarity_img``; one per concurrently running launch."""
    return torch.zeros(N.lib().cmhar_similarity_img_work_bytes(na, nb), dtype=torch.uint8, device=device)


class SigmoidContrastiveLoss(nn.Module):
    """reference src/models/losses.py:9-54.  NOTE: as written in the reference the
    +-1 labels multiply the logits AND {0,1} targets go to BCE, which is algebraically
    ``mean_ij softplus(-(t*s_ij + b))`` over all pairs; parity requires reproducing exactly that."""

    def __init__(self, init_temperature=10.0, init_bias=-10.0, learnable=True):
        super().__init__()
        if learnable:
            self.temperature = nn.Parameter(torch.tensor(init_temperature).log())
            self.bias = nn.Parameter(torch.tensor(init_bias))
        else:
            self.register_buffer("temperature", torch.tensor(init_temperature).log())
            self.register_buffer("bias", torch.tensor(init_bias))
        self._scalars = None          # cached host copies of (exp(temperature), bias)

    def _host_scalars(self):
        key = (self.temperature._version, self.bias._version, self.temperature.data_ptr())
        if self._scalars is None or self._scalars[0] != key:
            self._scalars = (key, float(self.temperature.detach().exp()), float(self.bias.detach()))
        return self._scalars[1], self._scalars[2]

    def forward(self, imu_embeds, video_embeds):
        if not torch.is_grad_enabled() and imu_embeds.is_cuda:
            t, b = self._host_scalars()
            n = imu_embeds.shape[0] * video_embeds.shape[0]
            res = similarity_native(imu_embeds, video_embeds, sigmoid=(t, b))
            return (res["sigmoid_sum"] / n).to(torch.float32)
        if not torch.is_grad_enabled():
            N.require_cuda(imu_embeds, "SigmoidContrastiveLoss")
        z = imu_embeds @ video_embeds.T * self.temperature.exp() + self.bias
        return F.softplus(-z).mean()


class InfoNCELoss(nn.Module):
    """reference src/models/losses.py:57-87: symmetric cross-entropy with diagonal targets."""

    def __init__(self, temperature=0.07):
        super().__init__()
        self.temperature = temperature

    def forward(self, imu_embeds, video_embeds):
        if not torch.is_grad_enabled():
            N.require_cuda(imu_embeds, "InfoNCELoss")
            res = similarity_native(imu_embeds, video_embeds, lse_scale=1.0 / self.temperature)
            return ((res["row_lse"] - res["diag"]).mean() + (res["col_lse"] - res["diag"]).mean()) / 2
        logits = imu_embeds @ video_embeds.T / self.temperature
        labels = torch.arange(imu_embeds.shape[0], device=logits.device)
        return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.T, labels)) / 2
