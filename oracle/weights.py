"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic parameters and inputs.

The reference pins no weights and there is no network for checkpoints, so every parity test,
golden fixture and bench run uses parameters drawn from ``numpy.random.RandomState`` (the
frozen legacy MT19937 stream, identical on every numpy version and every box).  The SAME
dictionary is loaded (``load_state_dict(strict=True)``) into the unmodified reference modules
when the golden vectors are generated (``oracle/make_golden.py``) and into this repo's drop-in
modules on the GPU box, so the goldens never have to carry megabytes of weights.

Key names and shapes follow the reference ``state_dict`` (reference
``src/models/models.py:24-27,78-98,226-231,254-268,312-326``).
BatchNorm running statistics and affine terms are randomised so that BN folding is exercised
(SURVEY.md section 8d).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np


@dataclass
class Dims:
    """Shape parameters of the path; defaults = reference ``configs/config.py:53-95``."""

    imu_channels: int = 6          # configs/config.py:56
    imu_window: int = 250          # configs/config.py:53
    patch: int = 16                # configs/config.py:77
    stride: int = 16               # configs/config.py:78
    d_model: int = 128             # configs/config.py:79
    nhead: int = 8                 # configs/config.py:80
    layers: int = 4                # configs/config.py:81
    video_feature_dim: int = 512   # resnet18 trunk, src/models/models.py:167
    video_d_model: int = 768       # configs/config.py:87
    proj_hidden: int = 512         # configs/config.py:91
    proj_dim: int = 256            # configs/config.py:90
    num_classes: int = 32          # configs/config.py:94
    head_hidden: List[int] = field(default_factory=lambda: [256, 128])  # configs/config.py:95

    @property
    def ffn(self) -> int:          # src/models/models.py:88 (dim_feedforward = 4*d_model)
        return 4 * self.d_model

    @property
    def num_patches(self) -> int:  # src/models/models.py:81
        return (self.imu_window - self.patch) // self.stride + 1

    @property
    def seq(self) -> int:
        """Effective token count after the positional-encoding truncation quirk
        (src/models/models.py:122-123): min(1 + C*N, N + 1) = N + 1."""
        return min(1 + self.imu_channels * self.num_patches, self.num_patches + 1)


def _uniform(rs: np.random.RandomState, shape, bound: float) -> np.ndarray:
    return rs.uniform(-bound, bound, size=shape).astype(np.float32)


def _linear(rs, out_f: int, in_f: int, prefix: str, sd: Dict[str, np.ndarray]) -> None:
    bound = 1.0 / np.sqrt(in_f)
    sd[prefix + ".weight"] = _uniform(rs, (out_f, in_f), bound)
    sd[prefix + ".bias"] = _uniform(rs, (out_f,), bound)


def _layernorm(rs, n: int, prefix: str, sd) -> None:
    sd[prefix + ".weight"] = (1.0 + 0.1 * rs.standard_normal(n)).astype(np.float32)
    sd[prefix + ".bias"] = (0.1 * rs.standard_normal(n)).astype(np.float32)


def _batchnorm(rs, n: int, prefix: str, sd) -> None:
    sd[prefix + ".weight"] = rs.uniform(0.5, 1.5, size=n).astype(np.float32)
    sd[prefix + ".bias"] = (0.2 * rs.standard_normal(n)).astype(np.float32)
    sd[prefix + ".running_mean"] = (0.3 * rs.standard_normal(n)).astype(np.float32)
    sd[prefix + ".running_var"] = rs.uniform(0.5, 2.0, size=n).astype(np.float32)
    sd[prefix + ".num_batches_tracked"] = np.array(7, dtype=np.int64)


def imu_encoder_state(seed: int, dims: Dims = Dims(), prefix: str = "") -> Dict[str, np.ndarray]:
    """Parameters of the reference ``IMUEncoder`` (src/models/models.py:59-98)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    d = dims.d_model
    # Scales are chosen so that the synthetic model is input-sensitive (a trained model is):
    # with torch's default init the CLS/positional terms swamp the patch embeddings and every
    # window gets the same arg-max, which would make the label-parity tests vacuous.
    sd[prefix + "cls_token"] = (0.5 * rs.standard_normal((1, 1, d))).astype(np.float32)
    sd[prefix + "pos_encoding"] = (0.5 * rs.standard_normal((1, dims.num_patches + 1, d))).astype(np.float32)
    for c in range(dims.imu_channels):
        _linear(rs, d, dims.patch, f"{prefix}patch_embed.projections.{c}", sd)
        sd[f"{prefix}patch_embed.projections.{c}.weight"] *= 4.0
    for l in range(dims.layers):
        p = f"{prefix}transformer.layers.{l}."
        # 1.5 x xavier: attention is input-dependent without being chaotic.  (At 2.5 x every single
        # bf16 rounding point moves the logits by ~2e-2 -- tools/bf16_error_budget.py -- so no bf16
        # implementation, torch autocast included, can meet the 2e-2 contract on such a model.)
        xav = 1.5 * np.sqrt(6.0 / (d + 3 * d))
        sd[p + "self_attn.in_proj_weight"] = _uniform(rs, (3 * d, d), xav)
        sd[p + "self_attn.in_proj_bias"] = (0.05 * rs.standard_normal(3 * d)).astype(np.float32)
        _linear(rs, d, d, p + "self_attn.out_proj", sd)
        _linear(rs, dims.ffn, d, p + "linear1", sd)
        _linear(rs, d, dims.ffn, p + "linear2", sd)
        _layernorm(rs, d, p + "norm1", sd)
        _layernorm(rs, d, p + "norm2", sd)
    _layernorm(rs, d, prefix + "norm", sd)
    return sd


def classifier_state(seed: int, dims: Dims = Dims()) -> Dict[str, np.ndarray]:
    """Parameters of the reference ``IMUClassifier`` (src/models/models.py:301-326)."""
    sd = imu_encoder_state(seed, dims, prefix="imu_encoder.")
    rs = np.random.RandomState(seed + 100003)
    in_dim, idx = dims.d_model, 0
    for h in dims.head_hidden:
        _linear(rs, h, in_dim, f"classifier.{idx}", sd)
        sd[f"classifier.{idx}.weight"] *= 3.0          # window-dependent part dominates the biases
        _batchnorm(rs, h, f"classifier.{idx + 1}", sd)
        in_dim, idx = h, idx + 4
    _linear(rs, dims.num_classes, in_dim, f"classifier.{idx}", sd)
    # spread the logits so arg-max margins are not degenerate
    sd[f"classifier.{idx}.weight"] *= 4.0
    return sd


def projection_head_state(seed: int, in_dim: int, hidden: int, out_dim: int,
                          prefix: str) -> Dict[str, np.ndarray]:
    """Parameters of the reference ``ProjectionHead`` (src/models/models.py:224-231)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    _linear(rs, hidden, in_dim, prefix + "net.0", sd)
    _batchnorm(rs, hidden, prefix + "net.1", sd)
    _linear(rs, out_dim, hidden, prefix + "net.3", sd)
    return sd


def cross_modal_state(seed: int, dims: Dims = Dims()) -> Dict[str, np.ndarray]:
    """Parameters of the reference ``CrossModalModel`` WITHOUT ``video_encoder.backbone.*``
    (third-party trunk, out of scope; src/models/models.py:244-268)."""
    sd = imu_encoder_state(seed, dims, prefix="imu_encoder.")
    rs = np.random.RandomState(seed + 200003)
    _linear(rs, dims.video_d_model, dims.video_feature_dim, "video_encoder.projection", sd)
    sd.update(projection_head_state(seed + 300007, dims.d_model, dims.proj_hidden,
                                    dims.proj_dim, "imu_proj."))
    sd.update(projection_head_state(seed + 400009, dims.video_d_model, dims.proj_hidden,
                                    dims.proj_dim, "video_proj."))
    sd["temperature"] = np.array(np.log(10.0), dtype=np.float32)
    sd["bias"] = np.array(-10.0, dtype=np.float32)
    return sd


# ----------------------------------------------------------------------------- inputs

def imu_windows(seed: int, n: int, dims: Dims = Dims()) -> np.ndarray:
    """(n, C, L) fp32 ~ N(0,1): z-scored IMU windows (reference src/data/preprocessing.py:216-219,
    layout (C,T) per src/data/datasets.py:139-141)."""
    rs = np.random.RandomState(seed)
    return rs.standard_normal((n, dims.imu_channels, dims.imu_window)).astype(np.float32)


def video_feature_maps(seed: int, n: int, frames: int = 16, dims: Dims = Dims(),
                       hw: int = 4) -> np.ndarray:
    """(n*frames, F, hw, hw) fp32 >= 0: post-ReLU trunk feature maps (resnet18 @112 gives 4x4,
    SURVEY.md section 8a row a4)."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n * frames, dims.video_feature_dim, hw, hw)).astype(np.float32)
    return np.maximum(x, 0.0)


def class_features(seed: int, n: int, num_classes: int = 32, dim: int = 128,
                   ood_fraction: float = 0.0, mean_seed: int = 777):
    """Synthetic class-conditional features for the Mahalanobis rows (SURVEY.md section 8d cfg 4):
    f ~ N(mu_c, I), mu_c ~ N(0, 4I); an ``ood_fraction`` of rows is drawn around held-out means
    and labelled -1."""
    rm = np.random.RandomState(mean_seed)            # class means are shared across sample seeds
    mu = (2.0 * rm.standard_normal((num_classes, dim))).astype(np.float32)
    mu_ood = (2.0 * rm.standard_normal((max(num_classes // 4, 1), dim))).astype(np.float32)
    rs = np.random.RandomState(seed)
    labels = rs.randint(0, num_classes, size=n).astype(np.int64)
    feats = mu[labels] + rs.standard_normal((n, dim)).astype(np.float32)
    n_ood = int(round(n * ood_fraction))
    if n_ood:
        which = rs.randint(0, mu_ood.shape[0], size=n_ood)
        feats[n - n_ood:] = mu_ood[which] + rs.standard_normal((n_ood, dim)).astype(np.float32)
        labels[n - n_ood:] = -1
    return feats.astype(np.float32), labels
