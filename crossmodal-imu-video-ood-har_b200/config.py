"""Minimal configuration tree with the attribute names the hot-path modules read from the
reference's global ``CONFIG`` (reference configs/config.py:49-95).  The reference's own ``CONFIG``
object can be passed to every class in this package instead; this stand-in exists because the
reference tree is not importable on the GPU box and creates ``./outputs`` as an import side
effect (configs/config.py:33-46)."""
from __future__ import annotations

from types import SimpleNamespace


def default_config(imu_window_size: int = 250, video_backbone="identity", video_feature_dim: int = 512,
                   num_classes: int = 32) -> SimpleNamespace:
    data = SimpleNamespace(imu_window_size=imu_window_size, imu_stride=125, imu_sampling_rate=50,
                           imu_channels=6, video_fps=25, video_frames_per_window=16,
                           video_resize=(224, 224))
    model = SimpleNamespace(imu_patch_size=16, imu_stride=16, imu_d_model=128, imu_nhead=8,
                            imu_num_layers=4, imu_dropout=0.1,
                            video_backbone=video_backbone, video_pretrained=False, video_d_model=768,
                            video_feature_dim=video_feature_dim,
                            projection_dim=256, projection_hidden_dim=512,
                            num_classes=num_classes, classifier_hidden_dims=[256, 128],
                            classifier_dropout=0.3)
    training = SimpleNamespace(seed=42, device="cuda", num_workers=2, train_batch_size=64,
                               pretrain_batch_size=16, temperature=0.07, use_sigmoid_loss=True)
    return SimpleNamespace(data=data, model=model, training=training)
