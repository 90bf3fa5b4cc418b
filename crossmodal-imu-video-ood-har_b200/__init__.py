"""crossmodal-imu-video-ood-har_b200 -- B200-native (sm_100a) cross-modal IMU/video inference and
OOD-scoring hot path, behind the reference's ``nn.Module`` / evaluator surface.

Import name: ``crossmodal_imu_video_ood_har_b200`` (the hyphenated directory is not a valid
Python identifier; ``crossmodal_imu_video_ood_har_b200/__init__.py`` at the repo root aliases it).
"""
from . import _native
from .config import default_config
from .models import (PatchEmbedding, IMUEncoder, VideoEncoder, ProjectionHead, CrossModalModel,
                     IMUClassifier, set_default_precision, get_default_precision)
from .fusion import LateFusionClassifier, CrossAttentionFusionClassifier, head_scores_native
from .conv_encoder import ConvIMUEncoder, ConvIMUClassifier
from .losses import SigmoidContrastiveLoss, InfoNCELoss, similarity_native
from .ood import (logit_scores, MahalanobisOOD, ScoreHistogram, auroc_fpr95, roc_from_histograms,
                  finalize_mahalanobis)
from .evaluator import Evaluator, classification_metrics, shard_bounds
from .pipeline import CrossModalOODPipeline
from .shards import WindowShard, write_shard, pack_npy_windows, live_samples
from .tables import generate_ood_table, ood_rows, save_tables
from .sweep import OODSweep, held_out_activity_split
from .video_trunk import DeviceVideoTrunk, fold_conv_bn

__version__ = "0.1.0"
