"""Cross-modal fusion classifiers (north-star item 3: "cross-attention or
concat-MLP").

SPEC-DEFINED -- NOT IN THE REFERENCE.  /root/reference has no fusion classifier: its only
cross-modal coupling is two projection heads and the similarity matrix inside the loss.  BASELINE.json's
configs[1] ("IMU+video late-fusion classifier") and configs[2] ("cross-attention fusion") name these blocks
anyway, so they are defined here explicitly, built from the reference's own pieces (its ``IMUEncoder``,
``VideoEncoder`` tail and classifier-head layout), with a plain-PyTorch restatement in
``oracle/fusion_spec.py``.  Every result is "self-consistent with the in-repo spec", never reference parity.

``LateFusionClassifier``     f = ReLU(BN(Linear([imu_cls (128) | video_feat (768)] -> 128)))  (concat-MLP)
                              logits = classifier_head(f)      head = the reference's layout
                              (src/models/models.py:312-326: [Linear, BN, ReLU, Dropout] x 2, Linear)
``CrossAttentionFusionClassifier``
                              q = Linear(imu_tokens), [k|v] = Linear(frame_feats); 8-head attention of the S IMU
                              tokens over the T frame tokens; y = LayerNorm(imu_tokens + Linear(attn));
                              f = mean_s y; logits = classifier_head(f)

Inference (eval + no_grad, CUDA tensors) runs only hand-written kernels: the fused IMU encoder, the video
pooling kernel, ``cmhar_concat_linear_forward`` / ``cmhar_linear_forward``, ``cmhar_cross_attention``,
``cmhar_residual_ln_pool`` and the head + OOD-score kernel.  Training routes through differentiable torch ops
on the same parameters, exactly like the other modules of this package.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from .models import (IMUEncoder, VideoEncoder, _PackedLinear, _PackedMixin, _native_mode, _prec_code,
                     imu_forward_native, pack_head_blob)

__all__ = ["LateFusionClassifier", "CrossAttentionFusionClassifier", "head_scores_native"]


def _head(in_dim: int, config) -> nn.Sequential:
    m = config.model
    layers = []
    for hidden in m.classifier_hidden_dims:
        layers += [nn.Linear(in_dim, hidden), nn.BatchNorm1d(hidden), nn.ReLU(inplace=True), nn.Dropout(m.classifier_dropout)]
        in_dim = hidden
    layers.append(nn.Linear(in_dim, m.num_classes))
    return nn.Sequential(*layers)


def head_scores_native(head_blob: torch.Tensor, maha_blob: Optional[torch.Tensor], feat: torch.Tensor, classes: int,
                       out: Optional[Dict[str, torch.Tensor]] = None, precision: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Classifier head + arg-max / MSP / energy (+ Mahalanobis) on stored (n,128) features: one launch."""
    feat = N.f32c(feat)
    n, dev = feat.shape[0], feat.device
    out = {} if out is None else out

    def buf(name, shape, dtype=torch.float32):
        t = out.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t
    logits, pred = buf("logits", (n, classes)), buf("pred", (n,), torch.int64)
    msp, energy = buf("msp", (n,)), buf("energy", (n,))
    maha = buf("maha", (n,)) if maha_blob is not None else None
    with torch.cuda.device(dev):
        N.check(N.lib().cmhar_head_forward(head_blob.data_ptr(), N.ptr(maha_blob), feat.data_ptr(), n, logits.data_ptr(),
                                           pred.data_ptr(), msp.data_ptr(), energy.data_ptr(), N.ptr(maha), _prec_code(precision),
                                           N.stream_ptr(dev)))
    return out


class _FusionBase(_PackedMixin, nn.Module):
    def __init__(self, imu_encoder: IMUEncoder, video_encoder: VideoEncoder, config):
        super().__init__()
        self.imu_encoder, self.video_encoder, self.config = imu_encoder, video_encoder, config
        self.d_model = config.model.imu_d_model
        self.num_classes = config.model.num_classes
        self.classifier = _head(self.d_model, config)
        self._maha_state = None

    def set_mahalanobis(self, maha) -> None:
        """Attach a fitted ``ood.MahalanobisOOD`` (on the fused 128-d feature)."""
        from .models import _PACK_GENERATION
        self._maha_state = maha
        _PACK_GENERATION[0] += 1          # recorded graphs of this model now compute a different set of outputs

    def _head_blob(self, device) -> torch.Tensor:
        key = ("head", str(device))
        if key not in self._packed:
            self._packed[key] = pack_head_blob(self.classifier, device)
        return self._packed[key]

    def _check(self):
        if self.training:
            raise RuntimeError("forward_scores is an inference entry point: call .eval() first")
        if self.d_model != 128:
            raise NotImplementedError("native fusion kernels are specialised to a 128-d fused feature")

    def _scores(self, fused: torch.Tensor, out: Optional[Dict[str, torch.Tensor]], precision: Optional[str] = None
                ) -> Dict[str, torch.Tensor]:
        dev = fused.device
        maha_blob = self._maha_state.blob(dev) if self._maha_state is not None else None
        res = head_scores_native(self._head_blob(dev), maha_blob, fused, self.num_classes, out, precision)
        res["fused"] = fused
        return res


class LateFusionClassifier(_FusionBase):
    """concat-MLP late fusion (spec-defined, see the module docstring).

    ``forward(imu (B,6,L), video (B,T,3,H,W)) -> logits (B, num_classes)``;
    ``forward_scores(imu, fmap (B*T,F,h,w), frames)`` is the fused inference entry used by the pipeline."""

    def __init__(self, imu_encoder: IMUEncoder, video_encoder: VideoEncoder, config):
        super().__init__(imu_encoder, video_encoder, config)
        m = config.model
        self.fusion = nn.Sequential(nn.Linear(m.imu_d_model + m.video_d_model, self.d_model),
                                    nn.BatchNorm1d(self.d_model), nn.ReLU(inplace=True), nn.Dropout(m.classifier_dropout))
        self._init_packed()

    def _fusion_packed(self, device) -> _PackedLinear:
        key = ("fusion", str(device))
        if key not in self._packed:
            self._packed[key] = _PackedLinear(self.fusion[0], self.fusion[1], device)
        return self._packed[key]

    def fuse_native(self, imu_cls: torch.Tensor, video_feat: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
        """ReLU(BN(Linear([imu_cls | video_feat]))) without materialising the concatenation."""
        pl = self._fusion_packed(imu_cls.device)
        a, b = N.f32c(imu_cls), N.f32c(video_feat)
        n = a.shape[0]
        y = torch.empty((n, pl.out_dim), dtype=torch.float32, device=a.device)
        lib = N.lib()
        work, wbytes = None, 0
        if 0 < n <= 2048 and not (_prec_code(precision) == N.BF16 and pl.in_dim % 64 == 0):
            wbytes = lib.cmhar_linear_work_bytes(n, pl.out_dim)
            work = torch.empty(wbytes, dtype=torch.uint8, device=a.device)
        with torch.cuda.device(a.device):
            N.check(lib.cmhar_concat_linear_forward(pl.blob.data_ptr(), a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], n,
                                                    pl.out_dim, 1, y.data_ptr(), N.ptr(work), wbytes, _prec_code(precision),
                                                    N.stream_ptr(a.device)))
        return y

    @torch.no_grad()
    def forward_scores_img(self, cls_img: torch.Tensor, vfeat_img: torch.Tensor, n: int,
                           out: Optional[Dict[str, torch.Tensor]] = None) -> Optional[Dict[str, torch.Tensor]]:
        """Fusion layer + head + arg-max / MSP / energy (+ Mahalanobis) as ONE launch (``cmhar_fused_head_forward``) from
        the two inputs' bf16 operand images.  Returns None when the dimensions are not served by the fused kernel."""
        self._check()
        dev = cls_img.device
        pl = self._fusion_packed(dev)
        out = {} if out is None else out

        def buf(name, shape, dtype=torch.float32):
            t = out.get(name)
            if t is None:
                t = out[name] = torch.empty(shape, dtype=dtype, device=dev)
            return t
        maha_blob = self._maha_state.blob(dev) if self._maha_state is not None else None
        fused, logits = buf("fused", (n, pl.out_dim)), buf("logits", (n, self.num_classes))
        pred, msp, energy = buf("pred", (n,), torch.int64), buf("msp", (n,)), buf("energy", (n,))
        maha = buf("maha", (n,)) if maha_blob is not None else None
        k1 = self.imu_encoder.d_model
        with torch.cuda.device(dev):
            rc = N.lib().cmhar_fused_head_forward(pl.blob.data_ptr(), cls_img.data_ptr(), k1, vfeat_img.data_ptr(), pl.in_dim - k1, n,
                                                  self._head_blob(dev).data_ptr(), N.ptr(maha_blob), fused.data_ptr(), logits.data_ptr(),
                                                  pred.data_ptr(), msp.data_ptr(), energy.data_ptr(), N.ptr(maha), N.stream_ptr(dev))
        if rc == N.UNSUPPORTED:
            return None
        N.check(rc)
        return out

    @torch.no_grad()
    def forward_scores(self, imu, fmap, frames: int, *, precision: Optional[str] = None, window_stride: Optional[int] = None,
                       imu_cls: Optional[torch.Tensor] = None, video_feat: Optional[torch.Tensor] = None,
                       out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """logits, pred, msp, energy (maha) of the fused classifier.  ``imu_cls`` / ``video_feat`` may be passed
        when the caller already holds them (the pipeline shares them with the contrastive branch)."""
        self._check()
        if imu_cls is None:
            imu_cls = imu_forward_native(self.imu_encoder, None, None, imu, want_cls=True, precision=precision,
                                         window_stride=window_stride)["cls"]
        if video_feat is None:
            video_feat = self.video_encoder.forward_features(fmap, frames, precision=precision)
        return self._scores(self.fuse_native(imu_cls, video_feat, precision), out, precision)

    def forward(self, imu, video):
        if _native_mode(self):
            N.require_cuda(imu, "LateFusionClassifier")
            cls = imu_forward_native(self.imu_encoder, None, None, imu, want_cls=True)["cls"]
            vfeat = self.video_encoder(video)
            return self._scores(self.fuse_native(cls, vfeat), None)["logits"]
        cls, _ = self.imu_encoder(imu)
        vfeat = self.video_encoder(video)
        return self.classifier(self.fusion(torch.cat([cls, vfeat], dim=1)))


class CrossAttentionFusionClassifier(_FusionBase):
    """Cross-attention fusion (spec-defined): the S IMU tokens attend over the T per-frame video tokens."""

    def __init__(self, imu_encoder: IMUEncoder, video_encoder: VideoEncoder, config):
        super().__init__(imu_encoder, video_encoder, config)
        m = config.model
        d = self.d_model
        self.nhead = m.imu_nhead
        self.q_proj = nn.Linear(d, d)
        self.kv_proj = nn.Linear(m.video_d_model, 2 * d)
        self.out_proj = nn.Linear(d, d)
        self.norm = nn.LayerNorm(d)
        self._init_packed()

    def _packed_linears(self, device):
        key = ("xattn", str(device))
        if key not in self._packed:
            self._packed[key] = tuple(_PackedLinear(l, None, device) for l in (self.q_proj, self.kv_proj, self.out_proj))
        return self._packed[key]

    def _packed_kv_folded(self, device) -> _PackedLinear:
        """[k|v] = kv_proj(projection(pooled frame)) is two Linear layers with nothing in between (models.py:213 has no
        activation), i.e. ONE Linear F -> 2d with W = W_kv W_p, b = W_kv b_p + b_kv (folded in fp64 at pack time): the
        per-frame F -> video_d_model GEMM (2/3 of the block's flops) disappears."""
        from .models import pack_generation
        key = ("xattn_kv_folded", str(device))
        # the fold reads the VIDEO encoder's projection weights: re-fold whenever any module re-packed (load_state_dict / train /
        # to() on the encoder alone does not reach this module's cache)
        if key not in self._packed or self._packed[key][0] != pack_generation():
            proj = self.video_encoder.projection
            wp, bp = proj.weight.detach().double(), proj.bias.detach().double()
            wk, bk = self.kv_proj.weight.detach().double(), self.kv_proj.bias.detach().double()
            lin = nn.Linear(proj.in_features, self.kv_proj.out_features)
            lin.weight.data.copy_((wk @ wp).float())
            lin.bias.data.copy_((wk @ bp + bk).float())
            self._packed[key] = (pack_generation(), _PackedLinear(lin.to(device), None, device))
        return self._packed[key][1]

    def _xattn_blob(self, device) -> Optional[torch.Tensor]:
        """Packed weights of the fused cross-attention kernel (``cmhar_xattn_pack``): the video projection folded into the k / v
        projections, the value bias folded into the out-projection bias (fp64), keyed by the pack generation."""
        from .models import pack_generation
        key = ("xattn_fused", str(device))
        if key not in self._packed or self._packed[key][0] != pack_generation():
            lib = N.lib()
            proj = self.video_encoder.projection
            F_dim = proj.in_features
            nbytes = lib.cmhar_xattn_blob_bytes(F_dim)
            if nbytes == 0:
                self._packed[key] = (pack_generation(), None)
                return None
            dd = lambda t: t.detach().double()
            d = self.d_model
            wkv = dd(self.kv_proj.weight) @ dd(proj.weight)                                  # (2d, F)
            bkv = dd(self.kv_proj.weight) @ dd(proj.bias) + dd(self.kv_proj.bias)
            bo = dd(self.out_proj.bias) + dd(self.out_proj.weight) @ bkv[d:]
            ts = [self.q_proj.weight, self.q_proj.bias, wkv[:d], wkv[d:], self.out_proj.weight, bo, self.norm.weight, self.norm.bias]
            ts = [t.detach().to(device=device, dtype=torch.float32).contiguous() for t in ts]
            blob = N.alloc_blob(nbytes, device)
            with torch.cuda.device(device):
                N.check(lib.cmhar_xattn_pack(*[t.data_ptr() for t in ts], F_dim, blob.data_ptr(), N.stream_ptr(device)))
                torch.cuda.current_stream(device).synchronize()
            self._packed[key] = (pack_generation(), blob)
        return self._packed[key][1]

    def fuse_native_img(self, tokens: torch.Tensor, frame_img: torch.Tensor, frames: int, fused_kernel: bool = True) -> torch.Tensor:
        """bf16 route from the per-frame pooled operand image (``VideoEncoder.pool_features_frames``): folded kv GEMM on
        tensor cores straight from the image, then attention, out-projection, residual LayerNorm + token mean."""
        if self.nhead != 8 or self.d_model != 128:
            raise NotImplementedError("native cross-attention is specialised to 8 heads of 16")
        B, S, d = tokens.shape
        dev = tokens.device
        blob = self._xattn_blob(dev) if fused_kernel else None
        if blob is not None:            # the whole block as ONE tcgen05 launch (csrc/xattn_tc.cu)
            fused = torch.empty((B, d), dtype=torch.float32, device=dev)
            tok = N.f32c(tokens)
            with torch.cuda.device(dev):
                rc = N.lib().cmhar_xattn_forward(blob.data_ptr(), tok.data_ptr(), frame_img.data_ptr(), B, S, frames,
                                                 self.video_encoder.projection.in_features, float(self.norm.eps), fused.data_ptr(),
                                                 N.stream_ptr(dev))
            if rc != N.UNSUPPORTED:
                N.check(rc)
                return fused
        ql, _, ol = self._packed_linears(dev)
        kvl = self._packed_kv_folded(dev)
        tok2 = N.f32c(tokens).reshape(B * S, d)
        q = ql(tok2, relu=False, precision="bf16")
        kv, _ = kvl.forward_img(B * frames, False, x_img=frame_img, want_rows=True, want_img=False)
        attn = torch.empty((B * S, d), dtype=torch.float32, device=dev)
        fused = torch.empty((B, d), dtype=torch.float32, device=dev)
        lib = N.lib()
        with torch.cuda.device(dev):
            N.check(lib.cmhar_cross_attention(q.data_ptr(), kv.data_ptr(), B, S, frames, attn.data_ptr(), N.stream_ptr(dev)))
            o = ol(attn, relu=False, precision="bf16")
            g, b = N.f32c(self.norm.weight.detach()), N.f32c(self.norm.bias.detach())
            N.check(lib.cmhar_residual_ln_pool(tok2.data_ptr(), o.data_ptr(), g.data_ptr(), b.data_ptr(), B, S,
                                               float(self.norm.eps), fused.data_ptr(), N.stream_ptr(dev)))
        return fused

    def fuse_native(self, tokens: torch.Tensor, frame_feats: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
        """tokens (B,S,128), frame_feats (B,T,video_d_model) -> fused (B,128)."""
        if self.nhead != 8 or self.d_model != 128:
            raise NotImplementedError("native cross-attention is specialised to 8 heads of 16")
        B, S, d = tokens.shape
        T = frame_feats.shape[1]
        dev = tokens.device
        ql, kvl, ol = self._packed_linears(dev)
        tok2 = N.f32c(tokens).reshape(B * S, d)
        q = ql(tok2, relu=False, precision=precision)
        kv = kvl(N.f32c(frame_feats).reshape(B * T, -1), relu=False, precision=precision)
        attn = torch.empty((B * S, d), dtype=torch.float32, device=dev)
        fused = torch.empty((B, d), dtype=torch.float32, device=dev)
        lib = N.lib()
        with torch.cuda.device(dev):
            N.check(lib.cmhar_cross_attention(q.data_ptr(), kv.data_ptr(), B, S, T, attn.data_ptr(), N.stream_ptr(dev)))
            o = ol(attn, relu=False, precision=precision)
            g, b = N.f32c(self.norm.weight.detach()), N.f32c(self.norm.bias.detach())
            N.check(lib.cmhar_residual_ln_pool(tok2.data_ptr(), o.data_ptr(), g.data_ptr(), b.data_ptr(), B, S,
                                               float(self.norm.eps), fused.data_ptr(), N.stream_ptr(dev)))
        return fused

    @torch.no_grad()
    def forward_scores(self, imu, fmap, frames: int, *, precision: Optional[str] = None,
                       out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        self._check()
        tokens = imu_forward_native(self.imu_encoder, None, None, imu, want_tokens=True, precision=precision)["tokens"]
        B = tokens.shape[0]
        ve = self.video_encoder
        if _prec_code(precision) == N.BF16 and ve.feature_dim % 64 == 0 and B > 0 and not ve.is_videomae:
            # bf16: one pooling pass emits the frame tokens as an operand image; projection folded into the kv GEMM
            _, frame_img = ve.pool_features_frames(fmap, frames, want_clip_img=False)
            return self._scores(self.fuse_native_img(tokens, frame_img, frames), out, precision)
        frame_feats = ve.forward_frame_features(fmap, precision=precision)
        return self._scores(self.fuse_native(tokens, frame_feats.view(B, frames, -1), precision), out, precision)

    def _fuse_autograd(self, tokens, frame_feats):
        B, S, d = tokens.shape
        H = self.nhead
        q = self.q_proj(tokens).view(B, S, H, d // H).transpose(1, 2)
        k, v = self.kv_proj(frame_feats).chunk(2, dim=-1)
        k = k.reshape(B, -1, H, d // H).transpose(1, 2)
        v = v.reshape(B, -1, H, d // H).transpose(1, 2)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, d)
        return self.norm(tokens + self.out_proj(a)).mean(dim=1)

    def forward(self, imu, video):
        B, T = video.shape[0], video.shape[1]
        if _native_mode(self):
            N.require_cuda(imu, "CrossAttentionFusionClassifier")
            ve = self.video_encoder
            if ve.is_videomae:
                raise NotImplementedError("cross-attention fusion needs a per-frame (CNN) trunk")
            # the trunk (third-party torch module) produces the feature maps; everything behind it is forward_scores
            fmap = ve.backbone(video.reshape(B * T, *video.shape[2:]))
            return self.forward_scores(imu, fmap, T)["logits"]
        _, tokens = self.imu_encoder(imu)
        frame_feats = self.video_encoder.frame_features(video)
        return self.classifier(self._fuse_autograd(tokens, frame_feats))
