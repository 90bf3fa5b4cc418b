"""1-D temporal conv / BatchNorm / ReLU IMU encoder (north-star item 1).

SPEC-DEFINED -- NOT IN THE REFERENCE.  The reference's IMU encoder is the PatchTST transformer
(src/models/models.py:53-132): there is no conv stack, so nothing here can be "parity".  The
block is defined below and restated in plain PyTorch in ``oracle/fusion_spec.py`` (``conv_encoder``); results
are self-consistent with that spec only.

    Conv1d(6 -> 32, k=5, s=1, p=2) BN ReLU -> Conv1d(32 -> 64, k=5, s=2, p=2) BN ReLU
    -> Conv1d(64 -> 128, k=5, s=2, p=2) BN ReLU -> mean over time -> (B, 128)

The 128-d output has the width of the reference encoder's CLS feature, so the reference's classifier head
layout, the OOD scorers and ``MahalanobisOOD`` apply unchanged (``ConvIMUClassifier``).
Inference (eval + no_grad, CUDA): ``cmhar_conv_encoder_forward`` (one CTA per window, shared-memory halo tiles)
followed by the head / score kernel.  Training: ordinary differentiable torch ops on the same parameters.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _native as N
from .fusion import _head, head_scores_native
from .models import _PackedMixin, _native_mode, _prec_code, pack_head_blob

__all__ = ["ConvIMUEncoder", "ConvIMUClassifier"]


class ConvIMUEncoder(_PackedMixin, nn.Module):
    def __init__(self, config=None, in_channels: Optional[int] = None):
        super().__init__()
        cin = in_channels if in_channels is not None else (config.data.imu_channels if config is not None else 6)
        if cin != 6:
            raise NotImplementedError("the native conv encoder is specialised to 6 IMU channels")
        self.out_dim = 128
        self.features = nn.Sequential(
            nn.Conv1d(cin, 32, 5, stride=1, padding=2), nn.BatchNorm1d(32), nn.ReLU(inplace=True),
            nn.Conv1d(32, 64, 5, stride=2, padding=2), nn.BatchNorm1d(64), nn.ReLU(inplace=True),
            nn.Conv1d(64, 128, 5, stride=2, padding=2), nn.BatchNorm1d(128), nn.ReLU(inplace=True))
        self._init_packed()

    def packed_blob(self, device) -> torch.Tensor:
        key = ("conv", str(device))
        if key not in self._packed:
            lib = N.lib()
            p = N.ConvEncoderParams()
            keep = []

            def dp(t):
                t = N.f32c(t.detach())
                keep.append(t)
                return t.data_ptr()
            for l in range(3):
                conv, bn = self.features[3 * l], self.features[3 * l + 1]
                q = p.layer[l]
                q.weight, q.bias = dp(conv.weight), (dp(conv.bias) if conv.bias is not None else None)
                q.bn_weight, q.bn_bias, q.bn_mean, q.bn_var = dp(bn.weight), dp(bn.bias), dp(bn.running_mean), dp(bn.running_var)
            blob = N.alloc_blob(lib.cmhar_conv_encoder_blob_bytes(), device)
            with torch.cuda.device(device):
                N.check(lib.cmhar_conv_encoder_pack(C.byref(p), blob.data_ptr(), N.stream_ptr(device)))
                torch.cuda.current_stream(device).synchronize()
            del keep
            self._packed[key] = blob
        return self._packed[key]

    def forward_native(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, precision: Optional[str] = None) -> torch.Tensor:
        """precision 'bf16' = the tensor-core implicit-GEMM kernel, 'fp32' = the CUDA-core kernel (default: the package default)."""
        N.require_cuda(x, "ConvIMUEncoder")
        x = N.f32c(x)
        B, Cc, L = x.shape
        if Cc != 6:
            raise ValueError("expected (B, 6, L) windows")
        feat = out if out is not None else torch.empty((B, self.out_dim), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.check(N.lib().cmhar_conv_encoder_forward_ex(self.packed_blob(x.device).data_ptr(), x.data_ptr(), B, L, Cc * L,
                                                          feat.data_ptr(), _prec_code(precision), N.stream_ptr(x.device)))
        return feat

    def forward(self, x):
        if _native_mode(self):
            return self.forward_native(x)
        return self.features(x).mean(dim=2)


class ConvIMUClassifier(_PackedMixin, nn.Module):
    """Conv encoder + the reference's classifier-head layout (src/models/models.py:312-326) + OOD scores."""

    def __init__(self, config, encoder: Optional[ConvIMUEncoder] = None):
        super().__init__()
        self.config = config
        self.encoder = encoder if encoder is not None else ConvIMUEncoder(config)
        self.classifier = _head(self.encoder.out_dim, config)
        self.num_classes = config.model.num_classes
        self._maha_state = None
        self._init_packed()

    def set_mahalanobis(self, maha) -> None:
        self._maha_state = maha

    def _head_blob(self, device) -> torch.Tensor:
        key = ("head", str(device))
        if key not in self._packed:
            self._packed[key] = pack_head_blob(self.classifier, device)
        return self._packed[key]

    @torch.no_grad()
    def forward_scores(self, imu, *, precision: Optional[str] = None, out: Optional[Dict[str, torch.Tensor]] = None
                       ) -> Dict[str, torch.Tensor]:
        if self.training:
            raise RuntimeError("forward_scores is an inference entry point: call .eval() first")
        feat = self.encoder.forward_native(imu, precision=precision)
        maha_blob = self._maha_state.blob(feat.device) if self._maha_state is not None else None
        res = head_scores_native(self._head_blob(feat.device), maha_blob, feat, self.num_classes, out, precision)
        res["cls"] = feat
        return res

    def forward(self, imu):
        if _native_mode(self):
            return self.forward_scores(imu)["logits"]
        return self.classifier(self.encoder(imu))
