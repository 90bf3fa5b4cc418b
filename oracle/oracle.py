"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path (rows a1-a9).

Every function restates, as plain tensor algebra on CPU, what the reference's PyTorch modules
compute in ``eval()`` mode, and cites the reference lines it follows.  Nothing here is shared
with the product path: the product (``crossmodal-imu-video-ood-har_b200/``) never imports this
module; ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

Pinning: the reference has no tests and no golden vectors.  This
restatement is pinned against the UNMODIFIED reference modules imported from /root/reference
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py``).

The arithmetic that the reference delegates to PyTorch (``nn.TransformerEncoderLayer``,
``nn.Linear``, ``nn.BatchNorm1d``, ``nn.LayerNorm``, ``F.normalize``, BCE-with-logits) is
restated from the installed torch 2.11.0 sources; the reference pins no torch version
(``requirements.txt`` is empty).  Citations starting with ``torch/`` are relative to
site-packages.

``dtype=torch.float64`` evaluates the same algebra in double precision (the "true" value used
to put error bars on both the reference and the CUDA path).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .weights import Dims

LN_EPS = 1e-5   # torch/nn/modules/normalization.py LayerNorm default, used at models.py:85-98
BN_EPS = 1e-5   # torch/nn/modules/batchnorm.py default, used at models.py:228,319


def _t(sd: Dict[str, np.ndarray], key: str, dtype) -> torch.Tensor:
    v = sd[key]
    if isinstance(v, torch.Tensor):
        return v.detach().to("cpu", dtype)
    return torch.from_numpy(np.asarray(v)).to(dtype)


def _layer_norm(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)          # biased variance
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def _bn_eval(x, sd, prefix, dtype):
    """nn.BatchNorm1d in eval mode = per-feature affine (models.py:228,319)."""
    rm, rv = _t(sd, prefix + ".running_mean", dtype), _t(sd, prefix + ".running_var", dtype)
    w, b = _t(sd, prefix + ".weight", dtype), _t(sd, prefix + ".bias", dtype)
    return (x - rm) / torch.sqrt(rv + BN_EPS) * w + b


# ------------------------------------------------------------------ a1: PatchEmbedding.forward
def patch_embed(x: torch.Tensor, sd, dims: Dims, prefix: str, dtype) -> torch.Tensor:
    """reference src/models/models.py:30-50: unfold(2, patch, stride) then one Linear per
    channel, stacked -> (B, C, N, d)."""
    B, C, L = x.shape
    n = (L - dims.patch) // dims.stride + 1
    idx = (torch.arange(n)[:, None] * dims.stride + torch.arange(dims.patch)[None, :])  # (N, P)
    patches = x[:, :, idx]                                                            # (B,C,N,P)
    out = []
    for c in range(C):
        w = _t(sd, f"{prefix}patch_embed.projections.{c}.weight", dtype)
        b = _t(sd, f"{prefix}patch_embed.projections.{c}.bias", dtype)
        out.append(patches[:, c] @ w.T + b)
    return torch.stack(out, dim=1)


# ------------------------------------------------------------------ a2: IMUEncoder.forward
def imu_encoder(x, sd, dims: Dims = Dims(), prefix: str = "", dtype=torch.float32
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference src/models/models.py:100-132 (+ torch/nn/modules/transformer.py:946-990,
    post-norm branch; MultiheadAttention with batch_first, no mask, dropout inactive).

    Reproduces the positional-encoding truncation (models.py:122-123): tokens are CLS followed
    by channel-major patches (models.py:114-119) and only the first ``pos_encoding.shape[1]``
    survive.  Returns (cls (B,d), tokens (B,S,d))."""
    x = torch.as_tensor(x).to(dtype)
    B = x.shape[0]
    d, H = dims.d_model, dims.nhead
    emb = patch_embed(x, sd, dims, prefix, dtype)                       # (B,C,N,d)  models.py:111
    _, C, N, _ = emb.shape
    tok = emb.reshape(B, C * N, d)                                      # models.py:114-115
    cls = _t(sd, prefix + "cls_token", dtype).expand(B, -1, -1)         # models.py:118
    tok = torch.cat([cls, tok], dim=1)                                  # models.py:119
    pos = _t(sd, prefix + "pos_encoding", dtype)
    S = min(tok.shape[1], pos.shape[1])                                 # models.py:122
    h = tok[:, :S] + pos[:, :S]                                         # models.py:123
    hd = d // H
    for l in range(dims.layers):                                        # models.py:126
        p = f"{prefix}transformer.layers.{l}."
        w_in, b_in = _t(sd, p + "self_attn.in_proj_weight", dtype), _t(sd, p + "self_attn.in_proj_bias", dtype)
        qkv = h @ w_in.T + b_in                                         # (B,S,3d)
        q, k, v = qkv.split(d, dim=-1)
        q = q.reshape(B, S, H, hd).transpose(1, 2)                      # (B,H,S,hd)
        k = k.reshape(B, S, H, hd).transpose(1, 2)
        v = v.reshape(B, S, H, hd).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) / np.sqrt(hd), dim=-1)
        a = (att @ v).transpose(1, 2).reshape(B, S, d)
        a = a @ _t(sd, p + "self_attn.out_proj.weight", dtype).T + _t(sd, p + "self_attn.out_proj.bias", dtype)
        h = _layer_norm(h + a, _t(sd, p + "norm1.weight", dtype), _t(sd, p + "norm1.bias", dtype))
        f = torch.relu(h @ _t(sd, p + "linear1.weight", dtype).T + _t(sd, p + "linear1.bias", dtype))
        f = f @ _t(sd, p + "linear2.weight", dtype).T + _t(sd, p + "linear2.bias", dtype)
        h = _layer_norm(h + f, _t(sd, p + "norm2.weight", dtype), _t(sd, p + "norm2.bias", dtype))
    tokens = _layer_norm(h, _t(sd, prefix + "norm.weight", dtype), _t(sd, prefix + "norm.bias", dtype))  # models.py:127
    return tokens[:, 0], tokens                                         # models.py:130-132


# ------------------------------------------------------------------ a3: IMUClassifier.forward
def classifier_head(feat, sd, dims: Dims = Dims(), dtype=torch.float32) -> torch.Tensor:
    """reference src/models/models.py:312-326,338: [Linear, BN(eval), ReLU, Dropout(off)] per
    hidden dim, then Linear(-> num_classes)."""
    t = torch.as_tensor(feat).to(dtype)
    idx = 0
    for _ in dims.head_hidden:
        t = t @ _t(sd, f"classifier.{idx}.weight", dtype).T + _t(sd, f"classifier.{idx}.bias", dtype)
        t = torch.relu(_bn_eval(t, sd, f"classifier.{idx + 1}", dtype))
        idx += 4
    return t @ _t(sd, f"classifier.{idx}.weight", dtype).T + _t(sd, f"classifier.{idx}.bias", dtype)


def imu_classifier(x, sd, dims: Dims = Dims(), dtype=torch.float32):
    """reference src/models/models.py:328-339. Returns (logits, cls_feature)."""
    feat, _ = imu_encoder(x, sd, dims, "imu_encoder.", dtype)
    return classifier_head(feat, sd, dims, dtype), feat


# ------------------------------------------------------------------ a4: VideoEncoder tail
def video_tail(fmap, sd, frames: int, dtype=torch.float32, prefix: str = "video_encoder."):
    """reference src/models/models.py:210-216 (CNN branch, AFTER the third-party trunk):
    adaptive_avg_pool2d(1,1) per frame -> Linear(F -> video_d_model) per frame -> mean over T."""
    fmap = torch.as_tensor(fmap).to(dtype)
    BT, F = fmap.shape[0], fmap.shape[1]
    B = BT // frames
    pooled = fmap.reshape(BT, F, -1).mean(-1)                           # models.py:210
    feats = pooled.reshape(B, frames, F)                                # models.py:211
    feats = feats @ _t(sd, prefix + "projection.weight", dtype).T + _t(sd, prefix + "projection.bias", dtype)
    return feats.mean(1)                                                # models.py:214-215


def video_cls_projection(cls_feat, sd, dtype=torch.float32, prefix: str = "video_encoder."):
    """reference src/models/models.py:201-205 (VideoMAE branch after the HF trunk): Linear on the
    CLS token."""
    x = torch.as_tensor(cls_feat).to(dtype)
    return x @ _t(sd, prefix + "projection.weight", dtype).T + _t(sd, prefix + "projection.bias", dtype)


# ------------------------------------------------------------------ a5: ProjectionHead.forward
def projection_head(x, sd, prefix: str, dtype=torch.float32):
    """reference src/models/models.py:226-234: Linear -> BN1d(eval) -> ReLU -> Linear."""
    x = torch.as_tensor(x).to(dtype)
    t = x @ _t(sd, prefix + "net.0.weight", dtype).T + _t(sd, prefix + "net.0.bias", dtype)
    t = torch.relu(_bn_eval(t, sd, prefix + "net.1", dtype))
    return t @ _t(sd, prefix + "net.3.weight", dtype).T + _t(sd, prefix + "net.3.bias", dtype)


def l2_normalize(x, eps: float = 1e-12):
    """F.normalize(dim=1): x / max(||x||_2, eps) (models.py:288-289)."""
    return x / x.norm(dim=1, keepdim=True).clamp_min(eps)


# ------------------------------------------------------------------ a6: CrossModalModel.forward
def cross_modal(imu, fmap, sd, frames: int, dims: Dims = Dims(), dtype=torch.float32):
    """reference src/models/models.py:270-291 with the trunk's feature map as the video input.
    Returns (imu_proj, video_proj), both unit-norm rows."""
    imu_feat, _ = imu_encoder(imu, sd, dims, "imu_encoder.", dtype)     # models.py:280
    vid_feat = video_tail(fmap, sd, frames, dtype)                      # models.py:281
    ip = projection_head(imu_feat, sd, "imu_proj.", dtype)              # models.py:284
    vp = projection_head(vid_feat, sd, "video_proj.", dtype)            # models.py:285
    return l2_normalize(ip), l2_normalize(vp)


# ------------------------------------------------------------------ a7: contrastive losses
def similarity_matrix(imu_embeds, video_embeds, dtype=torch.float32):
    """reference src/models/losses.py:37: logits = imu @ video.T."""
    return torch.as_tensor(imu_embeds).to(dtype) @ torch.as_tensor(video_embeds).to(dtype).T


def sigmoid_contrastive_loss(imu_embeds, video_embeds, log_temperature=float(np.log(10.0)),
                             bias=-10.0, dtype=torch.float32):
    """reference src/models/losses.py:25-54, restated literally: z = S*exp(log_t) + bias;
    labels = 2*eye-1; BCE_with_logits(z*labels, (labels+1)/2), mean.
    (Algebraically mean softplus(-z) over ALL pairs)"""
    S = similarity_matrix(imu_embeds, video_embeds, dtype)
    z = S * float(np.exp(log_temperature)) + bias                      # losses.py:40-41
    n = S.shape[0]
    labels = 2 * torch.eye(n, dtype=dtype) - 1                         # losses.py:44
    u, tgt = z * labels, (labels + 1) / 2                              # losses.py:48-50
    # BCE with logits: max(u,0) - u*t + log1p(exp(-|u|))
    loss = torch.clamp_min(u, 0) - u * tgt + torch.log1p(torch.exp(-u.abs()))
    return loss.mean()                                                 # losses.py:51


def info_nce_loss(imu_embeds, video_embeds, temperature=0.07, dtype=torch.float32):
    """reference src/models/losses.py:67-87: symmetric cross-entropy with diagonal targets."""
    S = similarity_matrix(imu_embeds, video_embeds, dtype) / temperature  # losses.py:76
    diag = torch.diagonal(S)
    l_i2v = (torch.logsumexp(S, dim=1) - diag).mean()                  # losses.py:82
    l_v2i = (torch.logsumexp(S, dim=0) - diag).mean()                  # losses.py:83
    return (l_i2v + l_v2i) / 2                                         # losses.py:85


# ------------------------------------------------------------------ a8: Evaluator.predict
def predict(logits) -> np.ndarray:
    """reference src/eval/evaluator.py:45: ``_, preds = logits.max(1)`` (first max index)."""
    return torch.as_tensor(logits).max(1)[1].numpy().astype(np.int64)


# ------------------------------------------------------------------ a9: Evaluator.compute_metrics
def compute_metrics(y_true, y_pred) -> Dict[str, float]:
    """reference src/eval/evaluator.py:55-65, restated without sklearn (confusion-matrix algebra):
    accuracy, balanced accuracy (mean recall over classes present in y_true), macro/weighted F1,
    macro precision/recall with zero_division=0, all x100.  Label set = union(y_true, y_pred) as
    sklearn does."""
    y_true = np.asarray(y_true).astype(np.int64)
    y_pred = np.asarray(y_pred).astype(np.int64)
    labels = np.union1d(y_true, y_pred)
    k = len(labels)
    remap = {int(v): i for i, v in enumerate(labels)}
    t = np.array([remap[int(v)] for v in y_true])
    p = np.array([remap[int(v)] for v in y_pred])
    cm = np.zeros((k, k), dtype=np.int64)
    np.add.at(cm, (t, p), 1)
    tp = np.diag(cm).astype(np.float64)
    support = cm.sum(1).astype(np.float64)
    predicted = cm.sum(0).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = np.where(support > 0, tp / support, 0.0)
        precision = np.where(predicted > 0, tp / predicted, 0.0)
        f1 = np.where(precision + recall > 0, 2 * precision * recall / (precision + recall), 0.0)
    present = support > 0
    return {
        "accuracy": float(tp.sum() / max(len(y_true), 1) * 100),
        "balanced_accuracy": float(recall[present].mean() * 100),
        "f1_macro": float(f1.mean() * 100),
        "f1_weighted": float((f1 * support).sum() / support.sum() * 100),
        "precision_macro": float(precision.mean() * 100),
        "recall_macro": float(recall.mean() * 100),
    }
