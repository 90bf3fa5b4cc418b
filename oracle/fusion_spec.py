"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch (CPU) definition of the SPEC-DEFINED fusion classifiers and the
conv/BN/ReLU IMU encoder.

PARITY UNPINNED: /root/reference contains no late-fusion classifier, no cross-attention block and no conv
encoder, so there is nothing to pin these against.  This file IS the definition the CUDA
path (``crossmodal-imu-video-ood-har_b200/fusion.py``, ``csrc/fusion.cu``, ``csrc/dense.cu``) is checked
against; results are "self-consistent with the in-repo spec", never reference parity.  The pieces that do
exist in the reference (IMU encoder, video tail, classifier-head layout) are taken from ``oracle/oracle.py``,
which is pinned.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import oracle
from .oracle import _bn_eval, _layer_norm, _t
from .weights import Dims, _batchnorm, _layernorm, _linear, cross_modal_state


def fusion_state(seed: int, dims: Dims = Dims()) -> Dict[str, np.ndarray]:
    """Deterministic parameters of both fusion classifiers: the cross-modal encoders' parameters
    (``weights.cross_modal_state``) + fusion layers + a classifier head of the reference layout."""
    sd = cross_modal_state(seed, dims)
    sd = {k: v for k, v in sd.items() if k.startswith(("imu_encoder.", "video_encoder."))}
    rs = np.random.RandomState(seed + 500011)
    d = dims.d_model
    _linear(rs, d, d + dims.video_d_model, "fusion.0", sd)          # late fusion: Linear(896 -> 128), BN
    _batchnorm(rs, d, "fusion.1", sd)
    _linear(rs, d, d, "q_proj", sd)                                   # cross attention
    _linear(rs, 2 * d, dims.video_d_model, "kv_proj", sd)
    sd["q_proj.weight"] *= 2.0
    sd["kv_proj.weight"][:d] *= 4.0
    _linear(rs, d, d, "out_proj", sd)
    _layernorm(rs, d, "norm", sd)
    in_dim, idx = d, 0
    for h in dims.head_hidden:
        _linear(rs, h, in_dim, f"classifier.{idx}", sd)
        sd[f"classifier.{idx}.weight"] *= 3.0
        _batchnorm(rs, h, f"classifier.{idx + 1}", sd)
        in_dim, idx = h, idx + 4
    _linear(rs, dims.num_classes, in_dim, f"classifier.{idx}", sd)
    sd[f"classifier.{idx}.weight"] *= 4.0
    return sd


def late_fusion(imu, fmap, sd, frames: int, dims: Dims = Dims(), dtype=torch.float32):
    """logits, fused = head(ReLU(BN(Linear([imu_cls | video_feat]))))."""
    cls, _ = oracle.imu_encoder(imu, sd, dims, "imu_encoder.", dtype)
    vfeat = oracle.video_tail(fmap, sd, frames, dtype)
    x = torch.cat([cls, vfeat], dim=1)
    f = x @ _t(sd, "fusion.0.weight", dtype).T + _t(sd, "fusion.0.bias", dtype)
    f = torch.relu(_bn_eval(f, sd, "fusion.1", dtype))
    return oracle.classifier_head(f, sd, dims, dtype), f


def frame_features(fmap, sd, frames: int, dtype=torch.float32):
    """Per-frame video tokens: spatial mean + projection, no temporal mean -> (B, T, video_d_model)."""
    fmap = torch.as_tensor(fmap).to(dtype)
    BT, F = fmap.shape[0], fmap.shape[1]
    pooled = fmap.reshape(BT, F, -1).mean(-1)
    feats = pooled @ _t(sd, "video_encoder.projection.weight", dtype).T + _t(sd, "video_encoder.projection.bias", dtype)
    return feats.reshape(BT // frames, frames, -1)


def cross_attention_fusion(imu, fmap, sd, frames: int, dims: Dims = Dims(), dtype=torch.float32):
    """logits, fused = head(mean_s LayerNorm(tokens + out_proj(MHA(q = tokens, k = v = frame tokens))))."""
    _, tokens = oracle.imu_encoder(imu, sd, dims, "imu_encoder.", dtype)
    ft = frame_features(fmap, sd, frames, dtype)
    B, S, d = tokens.shape
    H, hd = dims.nhead, d // dims.nhead
    q = tokens @ _t(sd, "q_proj.weight", dtype).T + _t(sd, "q_proj.bias", dtype)
    kv = ft @ _t(sd, "kv_proj.weight", dtype).T + _t(sd, "kv_proj.bias", dtype)
    k, v = kv[..., :d], kv[..., d:]
    q = q.reshape(B, S, H, hd).transpose(1, 2)
    k = k.reshape(B, -1, H, hd).transpose(1, 2)
    v = v.reshape(B, -1, H, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / np.sqrt(hd), dim=-1)
    a = (att @ v).transpose(1, 2).reshape(B, S, d)
    o = a @ _t(sd, "out_proj.weight", dtype).T + _t(sd, "out_proj.bias", dtype)
    y = _layer_norm(tokens + o, _t(sd, "norm.weight", dtype), _t(sd, "norm.bias", dtype))
    f = y.mean(1)
    return oracle.classifier_head(f, sd, dims, dtype), f


# ------------------------------------------------------------------ conv / BN / ReLU IMU encoder (spec-defined)
def conv_encoder_state(seed: int, dims: Dims = Dims()) -> Dict[str, np.ndarray]:
    """Parameters of ``ConvIMUClassifier``: encoder.features.{0,3,6} convs + {1,4,7} BatchNorms + a head."""
    rs = np.random.RandomState(seed + 600013)
    sd: Dict[str, np.ndarray] = {}
    for idx, (cin, cout) in zip((0, 3, 6), ((6, 32), (32, 64), (64, 128))):
        bound = 1.0 / np.sqrt(cin * 5)
        sd[f"encoder.features.{idx}.weight"] = (2.0 * rs.uniform(-bound, bound, size=(cout, cin, 5))).astype(np.float32)
        sd[f"encoder.features.{idx}.bias"] = rs.uniform(-bound, bound, size=cout).astype(np.float32)
        _batchnorm(rs, cout, f"encoder.features.{idx + 1}", sd)
    in_dim, idx = dims.d_model, 0
    for h in dims.head_hidden:
        _linear(rs, h, in_dim, f"classifier.{idx}", sd)
        sd[f"classifier.{idx}.weight"] *= 3.0
        _batchnorm(rs, h, f"classifier.{idx + 1}", sd)
        in_dim, idx = h, idx + 4
    _linear(rs, dims.num_classes, in_dim, f"classifier.{idx}", sd)
    sd[f"classifier.{idx}.weight"] *= 4.0
    return sd


def conv_encoder(x, sd, dtype=torch.float32, prefix: str = "encoder."):
    """Conv1d(6->32,k5,s1,p2) BN ReLU, Conv1d(32->64,k5,s2,p2) BN ReLU, Conv1d(64->128,k5,s2,p2) BN ReLU, time mean."""
    h = torch.as_tensor(x).to(dtype)
    for idx, stride in ((0, 1), (3, 2), (6, 2)):
        w, b = _t(sd, f"{prefix}features.{idx}.weight", dtype), _t(sd, f"{prefix}features.{idx}.bias", dtype)
        h = torch.nn.functional.conv1d(h, w, b, stride=stride, padding=2)
        h = _bn_eval(h.transpose(1, 2), sd, f"{prefix}features.{idx + 1}", dtype).transpose(1, 2)
        h = torch.relu(h)
    return h.mean(dim=2)


def conv_classifier(x, sd, dims: Dims = Dims(), dtype=torch.float32):
    f = conv_encoder(x, sd, dtype)
    return oracle.classifier_head(f, sd, dims, dtype), f
