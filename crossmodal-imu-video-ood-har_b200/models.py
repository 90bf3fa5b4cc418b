"""Drop-in ``nn.Module`` surface of the reference's ``src/models/models.py``.

Same class names, constructor signatures, attribute names and ``state_dict`` keys/shapes as the
reference, so ``main.py`` can construct these classes and load
its checkpoints with ``strict=True``.

Two execution routes share the same ``nn.Parameter`` objects:

* inference (``not self.training`` and gradients disabled, i.e. what ``Evaluator.predict`` does
  under ``@torch.no_grad()``): hand-written sm_100a kernels behind the C ABI
  (``include/cmhar_b200.h``).  CUDA tensors only; anything else raises -- no CPU fallback.
* training / autograd (``self.training`` or grad enabled): ordinary differentiable torch ops on
  the same parameters, written out functionally here.  This is what the reference's trainers need
  (``src/train/trainer.py:139,303``); it is not the hot path and carries no performance claim.

Packed weight blobs (BN folded, transposed, bf16 images) are built lazily on the first native call
and dropped whenever parameters can have changed (``train()``, ``load_state_dict``, ``.to()``,
or an explicit ``invalidate_packed()``).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N

__all__ = ["PatchEmbedding", "IMUEncoder", "VideoEncoder", "ProjectionHead", "CrossModalModel",
           "IMUClassifier", "set_default_precision", "get_default_precision", "pack_head_blob"]

_DEFAULT_PRECISION = "fp32"


def set_default_precision(p: str) -> None:
    """'fp32' (CUDA-core fp32, 1e-3 contract) or 'bf16' (tcgen05 bf16 GEMMs, 2e-2 contract)."""
    global _DEFAULT_PRECISION
    if p not in ("fp32", "bf16"):
        raise ValueError(f"unknown precision {p!r}")
    _DEFAULT_PRECISION = p


def get_default_precision() -> str:
    return _DEFAULT_PRECISION


def _prec_code(p: Optional[str]) -> int:
    p = p or _DEFAULT_PRECISION
    return N.BF16 if p == "bf16" else N.FP32


def _native_mode(module: nn.Module) -> bool:
    return (not module.training) and (not torch.is_grad_enabled())


_PACK_GENERATION = [0]


def pack_generation() -> int:
    """Bumped whenever any module drops its packed blobs (parameters may have changed): recorded CUDA graphs hold raw
    pointers into those blobs and must be re-recorded when this number moves."""
    return _PACK_GENERATION[0]


class _PackedMixin:
    """Cache of packed device blobs, invalidated whenever parameters may have changed."""

    def _init_packed(self):
        object.__setattr__(self, "_packed", {})
        self.register_load_state_dict_post_hook(lambda m, _: m.invalidate_packed())

    def invalidate_packed(self):
        _PACK_GENERATION[0] += 1
        self._packed.clear()
        for child in self.children():
            if isinstance(child, _PackedMixin):
                child.invalidate_packed()

    def train(self, mode: bool = True):
        self.invalidate_packed()
        return super().train(mode)

    def _apply(self, fn, *a, **kw):
        self.invalidate_packed()
        return super()._apply(fn, *a, **kw)

    def __deepcopy__(self, memo):
        # blobs are device buffers tied to this instance's parameters: never share them
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            # (the device trunk holds CUDA graphs and a snapshot of the backbone weights: rebuilt on demand, never copied)
            object.__setattr__(new, k, {} if k == "_packed" else None if k == "_device_trunk" else copy.deepcopy(v, memo))
        return new


# =============================================================================== IMU encoder
class PatchEmbedding(nn.Module):
    """Per-channel ``Linear(patch_size -> d_model)`` over non-overlapping/strided patches
    (reference src/models/models.py:16-50).  Kept for state_dict/key parity and the autograd
    route; at inference only channel 0 is live and the fused kernel reads its
    projection directly."""

    def __init__(self, in_channels, patch_size, stride, d_model):
        super().__init__()
        self.patch_size, self.stride, self.d_model = patch_size, stride, d_model
        self.projections = nn.ModuleList(nn.Linear(patch_size, d_model) for _ in range(in_channels))

    def forward(self, x):
        # (B, C, L) -> (B, C, N, d_model)
        windows = x.unfold(2, self.patch_size, self.stride)
        return torch.stack([proj(windows[:, c]) for c, proj in enumerate(self.projections)], dim=1)


class _SelfAttentionParams(nn.Module):
    """Parameter holder with ``nn.MultiheadAttention``'s key names (in_proj_weight, in_proj_bias,
    out_proj.weight, out_proj.bias)."""

    def __init__(self, d_model: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class _EncoderLayerParams(nn.Module):
    """Parameter holder with ``nn.TransformerEncoderLayer``'s key names."""

    def __init__(self, d_model: int, ffn: int, dropout: float):
        super().__init__()
        self.self_attn = _SelfAttentionParams(d_model)
        self.linear1 = nn.Linear(d_model, ffn)
        self.linear2 = nn.Linear(ffn, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.p_drop = dropout


class _EncoderStack(nn.Module):
    def __init__(self, d_model: int, ffn: int, layers: int, dropout: float):
        super().__init__()
        self.layers = nn.ModuleList(_EncoderLayerParams(d_model, ffn, dropout) for _ in range(layers))


class IMUEncoder(_PackedMixin, nn.Module):
    """PatchTST-style IMU encoder (reference src/models/models.py:53-132).

    ``forward(x: (B, C, L)) -> (cls (B, d_model), tokens (B, S, d_model))`` with
    S = min(1 + C*N, N + 1): the reference truncates the channel-major token sequence to the
    length of ``pos_encoding`` (models.py:122-123), which this class reproduces on both routes."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        self.in_channels = config.data.imu_channels
        self.patch_size, self.stride, self.d_model = m.imu_patch_size, m.imu_stride, m.imu_d_model
        self.nhead, self.num_layers = m.imu_nhead, m.imu_num_layers
        self.patch_embed = PatchEmbedding(self.in_channels, self.patch_size, self.stride, self.d_model)
        self.cls_token = nn.Parameter(torch.randn(1, 1, self.d_model))
        max_patches = (config.data.imu_window_size - self.patch_size) // self.stride + 1
        self.pos_encoding = nn.Parameter(torch.randn(1, max_patches + 1, self.d_model))
        self.transformer = _EncoderStack(self.d_model, 4 * self.d_model, self.num_layers, m.imu_dropout)
        self.norm = nn.LayerNorm(self.d_model)
        self._init_packed()

    # ------------------------------------------------------------------ shapes
    def _seq_len(self, L: int) -> int:
        n = (L - self.patch_size) // self.stride + 1
        return min(1 + self.in_channels * n, self.pos_encoding.shape[1])

    def _check_native_dims(self, L: int) -> int:
        S = self._seq_len(L)
        n = (L - self.patch_size) // self.stride + 1
        ok = (self.d_model == 128 and self.nhead == 8 and self.patch_size == 16 and self.stride == 16
              and 1 <= self.num_layers <= N.MAX_LAYERS and 2 <= S <= N.MAX_SEQ and S - 1 <= n)
        if not ok:
            raise NotImplementedError(
                "cmhar_b200 kernels are specialised to the reference configuration (d_model=128, nhead=8, "
                f"patch=stride=16, <= {N.MAX_LAYERS} layers, <= {N.MAX_SEQ} tokens); got d_model={self.d_model}, "
                f"nhead={self.nhead}, patch={self.patch_size}, stride={self.stride}, layers={self.num_layers}, "
                f"tokens={S}")
        return S

    # ------------------------------------------------------------------ packing
    def packed_blob(self, device, S: int) -> torch.Tensor:
        key = ("enc", str(device), S)
        blob = self._packed.get(key)
        if blob is None:
            lib = N.lib()
            p = N.ImuEncoderParams()
            p.seq, p.layers = S, self.num_layers
            keep = []

            def dp(t):
                t = N.f32c(t.detach())
                keep.append(t)
                return t.data_ptr()
            p.cls_token, p.pos_encoding = dp(self.cls_token), dp(self.pos_encoding)
            p.patch_weight = dp(self.patch_embed.projections[0].weight)
            p.patch_bias = dp(self.patch_embed.projections[0].bias)
            p.norm_weight, p.norm_bias = dp(self.norm.weight), dp(self.norm.bias)
            for l, layer in enumerate(self.transformer.layers):
                q = p.layer[l]
                q.in_proj_weight, q.in_proj_bias = dp(layer.self_attn.in_proj_weight), dp(layer.self_attn.in_proj_bias)
                q.out_proj_weight, q.out_proj_bias = dp(layer.self_attn.out_proj.weight), dp(layer.self_attn.out_proj.bias)
                q.linear1_weight, q.linear1_bias = dp(layer.linear1.weight), dp(layer.linear1.bias)
                q.linear2_weight, q.linear2_bias = dp(layer.linear2.weight), dp(layer.linear2.bias)
                q.norm1_weight, q.norm1_bias = dp(layer.norm1.weight), dp(layer.norm1.bias)
                q.norm2_weight, q.norm2_bias = dp(layer.norm2.weight), dp(layer.norm2.bias)
            blob = N.alloc_blob(lib.cmhar_imu_encoder_blob_bytes(S, self.num_layers), device)
            with torch.cuda.device(device):
                N.check(lib.cmhar_imu_encoder_pack(C.byref(p), blob.data_ptr(), N.stream_ptr(device)))
            del keep
            self._packed[key] = blob
        return blob

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        if _native_mode(self):
            out = imu_forward_native(self, None, None, x, want_cls=True, want_tokens=True)
            return out["cls"], out["tokens"]
        return self._forward_autograd(x)

    def encode_cls(self, x, precision: Optional[str] = None):
        """Inference-only shortcut: CLS embedding without materialising the token tensor."""
        return imu_forward_native(self, None, None, x, want_cls=True, precision=precision)["cls"]

    def _forward_autograd(self, x):
        B = x.shape[0]
        d, H = self.d_model, self.nhead
        emb = self.patch_embed(x)                                     # (B, C, N, d)
        tok = torch.cat([self.cls_token.expand(B, -1, -1), emb.flatten(1, 2)], dim=1)
        S = min(tok.shape[1], self.pos_encoding.shape[1])
        h = tok[:, :S] + self.pos_encoding[:, :S]
        for layer in self.transformer.layers:
            p = layer.p_drop if self.training else 0.0
            qkv = F.linear(h, layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias)
            q, k, v = (t.reshape(B, S, H, d // H).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
            a = F.scaled_dot_product_attention(q, k, v, dropout_p=p)
            a = layer.self_attn.out_proj(a.transpose(1, 2).reshape(B, S, d))
            h = layer.norm1(h + F.dropout(a, p, self.training))
            f = layer.linear2(F.dropout(F.relu(layer.linear1(h)), p, self.training))
            h = layer.norm2(h + F.dropout(f, p, self.training))
        h = self.norm(h)
        return h[:, 0], h


def imu_forward_native(encoder: "IMUEncoder", head_blob: Optional[torch.Tensor],
                       maha_blob: Optional[torch.Tensor], x: torch.Tensor, *, want_cls=False,
                       want_tokens=False, want_logits=False, want_pred=False, want_msp=False,
                       want_energy=False, want_maha=False, classes: int = 0,
                       precision: Optional[str] = None, window_stride: Optional[int] = None,
                       out: Optional[Dict[str, torch.Tensor]] = None, want_cls_img: bool = False) -> Dict[str, torch.Tensor]:
    """One fused launch of ``cmhar_imu_forward_ex`` on the current stream.  ``want_cls_img`` (bf16 precision): the CLS
    features also as a bf16 operand image (``out['cls_img']``) for the fused projection-head / fusion kernels.

    ``x`` is either the reference layout (B, C, L) -- only channel 0 is read, through the row
    stride, never copied -- or an already compacted (B, L') channel-0 buffer with
    ``window_stride`` given.  ``out`` may supply preallocated result tensors (CUDA-graph use)."""
    N.require_cuda(x, "IMUEncoder")
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(-1) != 1:
        x = x.contiguous()
    B, L = x.shape[0], x.shape[-1]
    if window_stride is not None:
        stride = window_stride
    elif B > 1:
        stride = x.stride(0)
    else:
        stride = x.numel()
    # tokens kept = min(1 + C*N, len(pos_encoding)) = min(1 + N, len(pos_encoding))  (F4)
    S = encoder._check_native_dims(L)
    if B > 0 and stride < 16 * (S - 1):
        raise ValueError(f"window stride {stride} shorter than the {16 * (S - 1)} live samples")
    dev = x.device
    blob = encoder.packed_blob(dev, S)
    out = {} if out is None else out

    def buf(name, want, shape, dtype=torch.float32):
        if not want:
            return None
        t = out.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t
    # the head / score launch reads the CLS features the encoder launch wrote: always materialised then
    cls = buf("cls", want_cls or head_blob is not None or maha_blob is not None, (B, encoder.d_model))
    tokens = buf("tokens", want_tokens, (B, S, encoder.d_model))
    logits = buf("logits", want_logits, (B, classes))
    pred = buf("pred", want_pred, (B,), torch.int64)
    msp = buf("msp", want_msp, (B,))
    energy = buf("energy", want_energy, (B,))
    maha = buf("maha", want_maha, (B,))
    cls_img = None
    if want_cls_img and B > 0:
        cls_img = out.get("cls_img")
        if cls_img is None:
            # rows the encoder does not write (past its last 8-window tile) must read as zeros, not stale memory; a whole number
            # of 128-row image tiles is written completely
            cls_img = out["cls_img"] = operand_image(B, encoder.d_model, dev, zero=(B % 128 != 0))
    o = N.ImuOutputs(N.ptr(cls), N.ptr(cls_img), N.ptr(tokens), N.ptr(logits), N.ptr(pred), N.ptr(msp), N.ptr(energy), N.ptr(maha))
    with torch.cuda.device(dev):
        N.check(N.lib().cmhar_imu_forward_ex(blob.data_ptr(), N.ptr(head_blob), N.ptr(maha_blob), x.data_ptr(), B, stride,
                                             C.byref(o), _prec_code(precision), N.stream_ptr(dev)))
    return out


# =============================================================================== dense helpers
class _PackedLinear:
    """y = relu?(BN_eval(x W^T + b)) through ``cmhar_linear_forward``; BN folded at pack time."""

    def __init__(self, linear: nn.Linear, bn: Optional[nn.BatchNorm1d], device):
        lib = N.lib()
        self.in_dim, self.out_dim = linear.in_features, linear.out_features
        self.blob = N.alloc_blob(lib.cmhar_linear_blob_bytes(self.in_dim, self.out_dim), device)
        w = N.f32c(linear.weight.detach())
        b = N.f32c(linear.bias.detach()) if linear.bias is not None else None
        bn_t = [None] * 4
        if bn is not None:
            bn_t = [N.f32c(t.detach()) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var)]
            if abs(bn.eps - 1e-5) > 1e-12:
                raise NotImplementedError("BatchNorm eps other than 1e-5 is not supported by the packed path")
        with torch.cuda.device(device):
            N.check(lib.cmhar_linear_pack(w.data_ptr(), N.ptr(b), *[N.ptr(t) for t in bn_t], self.in_dim,
                                          self.out_dim, self.blob.data_ptr(), N.stream_ptr(device)))

    def __call__(self, x: torch.Tensor, relu: bool, precision: Optional[str] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = N.f32c(x)
        n = x.shape[0]
        y = out if out is not None else torch.empty((n, self.out_dim), dtype=torch.float32, device=x.device)
        lib = N.lib()
        work, wbytes = None, 0
        tc_path = _prec_code(precision) == N.BF16 and self.in_dim % 64 == 0      # tensor-core tiles: no k-split workspace
        if 0 < n <= 2048 and not tc_path:        # small batches: workspace for the deterministic k-split
            wbytes = lib.cmhar_linear_work_bytes(n, self.out_dim)
            work = torch.empty(wbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            N.check(lib.cmhar_linear_forward(self.blob.data_ptr(), x.data_ptr(), n, self.in_dim, self.out_dim,
                                             int(relu), y.data_ptr(), N.ptr(work), wbytes, _prec_code(precision),
                                             N.stream_ptr(x.device)))
        return y

    def forward_img(self, n: int, relu: bool, x: Optional[torch.Tensor] = None, x_img: Optional[torch.Tensor] = None,
                    want_rows: bool = True, want_img: bool = False):
        """Tensor-core path with operand images on either side (bf16 precision only): the input is fp32 rows ``x`` or the
        previous layer's image ``x_img``; returns (rows or None, image or None)."""
        dev = (x if x is not None else x_img).device
        y = torch.empty((n, self.out_dim), dtype=torch.float32, device=dev) if want_rows else None
        img = operand_image(n, self.out_dim, dev) if want_img else None
        if x is not None:
            x = N.f32c(x)
        with torch.cuda.device(dev):
            N.check(N.lib().cmhar_linear_forward_img(self.blob.data_ptr(), N.ptr(x), N.ptr(x_img), n, self.in_dim, self.out_dim,
                                                     int(relu), N.ptr(y), N.ptr(img), N.stream_ptr(dev)))
        return y, img

    def img_capable(self) -> bool:
        return self.in_dim % 64 == 0 and self.out_dim % 64 == 0


def _is_channels_last_map(fmap: torch.Tensor) -> bool:
    """A 4-d feature map stored channels-last (what ``DeviceVideoTrunk`` returns) that ``cmhar_video_pool_nhwc`` can read in place."""
    if fmap.dim() != 4 or fmap.is_contiguous() or not fmap.is_contiguous(memory_format=torch.channels_last):
        return False
    per16 = 8 if fmap.dtype == torch.bfloat16 else 4
    return fmap.shape[1] % per16 == 0 and fmap.data_ptr() % 16 == 0


def operand_image(n: int, dim: int, device, zero: bool = False) -> torch.Tensor:
    """bf16 operand image for (n, dim) activations: [ceil(n/128)][dim/64] SWIZZLE_128B chunks of 16 KiB (what
    ``cmhar_linear_forward_img`` / ``cmhar_mlp2_forward_img`` write for the next kernel).  ``zero``: rows no kernel writes
    (past the last window) read as zeros instead of stale memory."""
    img = N.alloc_blob(N.lib().cmhar_operand_image_bytes(n, dim), device)
    if zero:
        img.zero_()
    return img


def l2_normalize_native(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    x = N.f32c(x)
    y = out if out is not None else torch.empty_like(x)
    with torch.cuda.device(x.device):
        N.check(N.lib().cmhar_l2_normalize(x.data_ptr(), x.shape[0], x.shape[1], y.data_ptr(), N.stream_ptr(x.device)))
    return y


# =============================================================================== video encoder
class VideoEncoder(_PackedMixin, nn.Module):
    """Video encoder (reference src/models/models.py:137-216).

    The trunk (HF VideoMAE, torchvision resnet18 / mobilenet_v2) is third-party code and runs as
    the ordinary torch module it is (out of scope) -- eagerly in fp32 by default, or, after
    ``attach_device_trunk()``, in channels-last bf16 under a CUDA graph (video_trunk.py).  Everything after it --
    spatial average pool, ``projection`` and the temporal mean -- is the hot path: one HBM-bound
    pooling kernel over the trunk's feature map followed by one small GEMM (pooling and the
    Linear commute, models.py:210-215)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        vb = m.video_backbone
        self.is_videomae = False
        if isinstance(vb, str) and ("videomae" in vb.lower() or "/" in vb):
            from transformers import VideoMAEModel
            self.is_videomae = True
            self.backbone = VideoMAEModel.from_pretrained(vb)
            self.feature_dim = self.backbone.config.hidden_size
        elif vb == "resnet18":
            from torchvision import models as tvm
            trunk = tvm.resnet18(weights="DEFAULT" if m.video_pretrained else None)
            self.backbone = nn.Sequential(*list(trunk.children())[:-2])
            self.feature_dim = 512
        elif vb == "mobilenet_v2":
            from torchvision import models as tvm
            self.backbone = tvm.mobilenet_v2(weights="DEFAULT" if m.video_pretrained else None).features
            self.feature_dim = 1280
        elif vb in (None, "none", "identity"):
            # extension: the caller feeds trunk feature maps (B, T, F, h, w) directly
            self.backbone = nn.Identity()
            self.feature_dim = int(getattr(m, "video_feature_dim", 512))
        else:
            raise ValueError(f"Backbone inconnu: {vb}")
        self.projection = nn.Linear(self.feature_dim, m.video_d_model)
        if not self.is_videomae:
            self.temporal_pool = nn.AdaptiveAvgPool1d(1)
        self._device_trunk, self._device_trunk_kwargs, self._device_trunk_stale = None, None, False
        self._init_packed()

    def attach_device_trunk(self, trunk=True, **kwargs):
        """Route the eval-mode forward of a CNN trunk through ``video_trunk.DeviceVideoTrunk`` (channels-last bf16, folded
        BatchNorm, CUDA graph; bf16 contract).  ``trunk``: a DeviceVideoTrunk, True (build one from the current ``backbone``
        weights -- call again after loading new weights) or None / False (back to the eager fp32 trunk).  With it the module
        also accepts decoded uint8 frames ``(B, T, H, W, 3)``, normalised on the device."""
        self._device_trunk_kwargs = dict(kwargs) if trunk is True else None
        if trunk is True:
            from .video_trunk import DeviceVideoTrunk
            trunk = DeviceVideoTrunk(self, **kwargs)
        self._device_trunk = trunk or None
        self._device_trunk_stale = False
        return self._device_trunk

    def invalidate_packed(self):
        # parameters may have changed: a trunk built here from the backbone's weights is a snapshot and is rebuilt on next use
        if getattr(self, "_device_trunk_kwargs", None) is not None:
            self._device_trunk_stale = True
        super().invalidate_packed()

    def _trunk(self):
        if self._device_trunk is None and getattr(self, "_device_trunk_kwargs", None) is not None:
            self._device_trunk_stale = True                      # dropped by a deep copy
        if getattr(self, "_device_trunk_stale", False):
            self.attach_device_trunk(True, **self._device_trunk_kwargs)
        return self._device_trunk

    def _packed_projection(self, device) -> _PackedLinear:
        key = ("proj", str(device))
        if key not in self._packed:
            self._packed[key] = _PackedLinear(self.projection, None, device)
        return self._packed[key]

    def pool_features(self, fmap: torch.Tensor, frames: int, out: Optional[torch.Tensor] = None,
                      coresident: bool = False, want_img: bool = False, want_rows: bool = True):
        """HBM-bound half of the native tail: fmap (B*T, F, h, w) bf16/fp32 -> spatio-temporal mean (B, F) fp32
        (reference models.py:210-211,215 commuted in front of the linear projection).  ``coresident`` selects the
        one-small-CTA-per-SM ring kernel that can run next to the encoder's CTAs (see include/cmhar_b200.h).
        ``want_img``: returns ``(rows or None, image)`` with the result (also) as a bf16 operand image for
        ``project_pooled(..., x_img=image)`` -- the projection then reads no fp32 rows at all."""
        N.require_cuda(fmap, "VideoEncoder")
        if fmap.dtype not in (torch.float32, torch.bfloat16):
            fmap = fmap.float()
        nhwc = _is_channels_last_map(fmap)
        if not nhwc:
            fmap = fmap.contiguous()
        BT, Fd = fmap.shape[0], fmap.shape[1]
        hw = fmap[0, 0].numel()
        if BT % frames:
            raise ValueError(f"{BT} frames do not split into clips of {frames}")
        B = BT // frames
        want_img = want_img and Fd % 64 == 0 and B > 0 and not coresident
        pooled = out
        if pooled is None and (want_rows or not want_img):
            pooled = torch.empty((B, Fd), dtype=torch.float32, device=fmap.device)
        with torch.cuda.device(fmap.device):
            if nhwc:            # the device trunk's output as it lies (video_trunk.py): physical (B*T, h*w, F), no permuted copy
                img = operand_image(B, Fd, fmap.device) if want_img else None
                N.check(N.lib().cmhar_video_pool_nhwc(fmap.data_ptr(), int(fmap.dtype == torch.bfloat16), B, frames, Fd, hw,
                                                      N.ptr(pooled), N.ptr(img), None, N.stream_ptr(fmap.device)))
                return (pooled, img) if want_img else pooled
            if want_img:
                img = operand_image(B, Fd, fmap.device)
                N.check(N.lib().cmhar_video_pool_img(fmap.data_ptr(), int(fmap.dtype == torch.bfloat16), B, frames, Fd, hw,
                                                     N.ptr(pooled), img.data_ptr(), N.stream_ptr(fmap.device)))
                return pooled, img
            fn = N.lib().cmhar_video_pool_coresident if coresident else N.lib().cmhar_video_pool
            N.check(fn(fmap.data_ptr(), int(fmap.dtype == torch.bfloat16), B, frames, Fd, hw,
                       pooled.data_ptr(), N.stream_ptr(fmap.device)))
        return pooled

    def pool_features_frames(self, fmap: torch.Tensor, frames: int, want_clip_img: bool = True):
        """ONE pass over fmap (B*T, F, h, w) -> (clip-mean operand image (B, F) or None, per-frame operand image (B*T, F)):
        the temporal-mean feature of the contrastive branch and the frame tokens of the cross-attention block from the same
        read of the feature maps (``cmhar_video_pool_frames_img``; bf16 operand images, F % 64 == 0)."""
        N.require_cuda(fmap, "VideoEncoder")
        if fmap.dtype not in (torch.float32, torch.bfloat16):
            fmap = fmap.float()
        nhwc = _is_channels_last_map(fmap)
        if not nhwc:
            fmap = fmap.contiguous()
        BT, Fd = fmap.shape[0], fmap.shape[1]
        hw = fmap[0, 0].numel()
        if BT % frames or Fd % 64:
            raise ValueError(f"{BT} frames of {Fd} channels: need whole clips of {frames} frames and channels % 64 == 0")
        B = BT // frames
        clip_img = operand_image(B, Fd, fmap.device) if want_clip_img else None
        frame_img = operand_image(BT, Fd, fmap.device)
        with torch.cuda.device(fmap.device):
            if nhwc:
                N.check(N.lib().cmhar_video_pool_nhwc(fmap.data_ptr(), int(fmap.dtype == torch.bfloat16), B, frames, Fd, hw, None,
                                                      N.ptr(clip_img), frame_img.data_ptr(), N.stream_ptr(fmap.device)))
                return clip_img, frame_img
            N.check(N.lib().cmhar_video_pool_frames_img(fmap.data_ptr(), int(fmap.dtype == torch.bfloat16), B, frames, Fd, hw, None,
                                                        N.ptr(clip_img), frame_img.data_ptr(), N.stream_ptr(fmap.device)))
        return clip_img, frame_img

    def project_pooled(self, pooled: Optional[torch.Tensor], precision: Optional[str] = None, want_img: bool = False,
                       x_img: Optional[torch.Tensor] = None, n: Optional[int] = None):
        """(B, F) pooled features -> (B, video_d_model): the reference's ``projection`` (models.py:213).  ``want_img``
        (bf16 precision): also returns the feature as a bf16 operand image for the projection head that follows;
        ``x_img`` (+ ``n`` rows): the pooled features as an operand image from ``pool_features(want_img=True)``."""
        dev = (pooled if pooled is not None else x_img).device
        lin = self._packed_projection(dev)
        rows = pooled.shape[0] if pooled is not None else int(n)
        if (want_img or x_img is not None) and _prec_code(precision) == N.BF16 and lin.img_capable() and rows > 0:
            y, img = lin.forward_img(rows, False, x=pooled if x_img is None else None, x_img=x_img, want_rows=True, want_img=want_img)
            return (y, img) if want_img else y
        if pooled is None:
            raise ValueError("project_pooled: the operand-image input needs the bf16 tensor-core path")
        y = lin(pooled, relu=False, precision=precision)
        return (y, None) if want_img else y

    def forward_features(self, fmap: torch.Tensor, frames: int, precision: Optional[str] = None) -> torch.Tensor:
        """Native tail: fmap (B*T, F, h, w) bf16/fp32 -> (B, video_d_model) fp32."""
        return self.project_pooled(self.pool_features(fmap, frames), precision)

    def forward_frame_features(self, fmap: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
        """Per-frame tail (cross-attention fusion needs frame tokens): fmap (B*T, F, h, w) -> (B*T, video_d_model),
        i.e. the reference's spatial average pool + ``projection`` (models.py:210-213) WITHOUT the temporal mean."""
        N.require_cuda(fmap, "VideoEncoder")
        if fmap.dtype not in (torch.float32, torch.bfloat16):
            fmap = fmap.float()
        nhwc = _is_channels_last_map(fmap)
        if not nhwc:
            fmap = fmap.contiguous()
        BT, Fd = fmap.shape[0], fmap.shape[1]
        pooled = torch.empty((BT, Fd), dtype=torch.float32, device=fmap.device)
        with torch.cuda.device(fmap.device):
            if nhwc:
                N.check(N.lib().cmhar_video_pool_nhwc(fmap.data_ptr(), int(fmap.dtype == torch.bfloat16), BT, 1, Fd, fmap[0, 0].numel(),
                                                      pooled.data_ptr(), None, None, N.stream_ptr(fmap.device)))
            else:
                N.check(N.lib().cmhar_video_pool(fmap.data_ptr(), int(fmap.dtype == torch.bfloat16), BT, 1, Fd, fmap[0, 0].numel(),
                                                 pooled.data_ptr(), N.stream_ptr(fmap.device)))
        return self._packed_projection(fmap.device)(pooled, relu=False, precision=precision)

    def frame_features(self, x):
        """(B, T, 3, H, W) -> per-frame features (B, T, video_d_model); CNN trunks only."""
        if self.is_videomae:
            raise NotImplementedError("frame_features needs a per-frame (CNN) trunk")
        B, T = x.shape[0], x.shape[1]
        if _native_mode(self) and self._trunk() is not None:
            return self.forward_frame_features(self._device_trunk(x)).view(B, T, -1)
        fmap = self.backbone(x.reshape(B * T, *x.shape[2:]))
        if _native_mode(self):
            return self.forward_frame_features(fmap).view(B, T, -1)
        return self.projection(fmap.mean(dim=(2, 3))).view(B, T, -1)

    def forward(self, x):
        B, T = x.shape[0], x.shape[1]
        native = _native_mode(self)
        if self.is_videomae:
            feat = self.backbone(pixel_values=x).last_hidden_state[:, 0]
            if native:
                N.require_cuda(feat, "VideoEncoder")
                return self._packed_projection(feat.device)(feat, relu=False)
            return self.projection(feat)
        if native and self._trunk() is not None:
            return self.forward_features(self._device_trunk(x), T)      # channels-last bf16 trunk under a CUDA graph (video_trunk.py)
        fmap = self.backbone(x.reshape(B * T, *x.shape[2:]))
        if native:
            return self.forward_features(fmap, T)
        feats = self.projection(fmap.mean(dim=(2, 3)).view(B, T, self.feature_dim))
        return self.temporal_pool(feats.transpose(1, 2)).squeeze(-1)


# =============================================================================== projection heads
class ProjectionHead(_PackedMixin, nn.Module):
    """Linear -> BatchNorm1d -> ReLU -> Linear (reference src/models/models.py:221-234)."""

    def __init__(self, in_dim, hidden_dim, out_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.BatchNorm1d(hidden_dim),
                                 nn.ReLU(inplace=True), nn.Linear(hidden_dim, out_dim))
        self._init_packed()

    def _packed_layers(self, device):
        key = ("mlp", str(device))
        if key not in self._packed:
            self._packed[key] = (_PackedLinear(self.net[0], self.net[1], device),
                                 _PackedLinear(self.net[3], None, device))
        return self._packed[key]

    def forward_native(self, x, precision: Optional[str] = None, x_img: Optional[torch.Tensor] = None):
        """Inference route with an explicit precision ('bf16' = tcgen05 tiles, 'fp32' = CUDA-core tiles).  ``x_img``:
        the input already as a bf16 operand image (``VideoEncoder.project_pooled(..., want_img=True)``)."""
        N.require_cuda(x, "ProjectionHead")
        l0, l1 = self._packed_layers(x.device)
        if _prec_code(precision) == N.BF16 and l0.img_capable() and l1.in_dim % 64 == 0 and x.dim() == 2 and x.shape[0] > 0:
            # the hidden activation goes from layer to layer as a bf16 operand image (never as fp32 rows): the second
            # layer's A operand is a plain bulk copy.  Same bits as the fp32-row hand-off (both round to bf16 once).
            n = x.shape[0]
            _, h_img = l0.forward_img(n, True, x=x if x_img is None else None, x_img=x_img, want_rows=False, want_img=True)
            return l1.forward_img(n, False, x_img=h_img, want_rows=True)[0]
        return l1(l0(x, relu=True, precision=precision), relu=False, precision=precision)

    def forward_fused(self, x_img: torch.Tensor, n: int, *, normalize: bool = True, want_rows: bool = True,
                      want_img: bool = True, img_out: Optional[int] = None):
        """The whole head (+ ``F.normalize``) as ONE tensor-core launch (``cmhar_mlp2_forward_img``): input and output as
        bf16 operand images, the hidden activation never leaves tensor memory.  Returns ``(rows (n, out) or None, image or
        None)``, or ``None`` when the dimensions are not the reference's (.. -> 512 -> 256): the caller chains
        ``forward_native`` + ``l2_normalize_native`` then.  ``img_out``: raw device pointer the output image is written to
        instead of a fresh allocation (a rank's slice of a peer-mapped buffer, ``peer.PeerBuffer``); the image slot of the
        result is then that pointer."""
        dev = x_img.device
        l0, l1 = self._packed_layers(dev)
        y = torch.empty((n, l1.out_dim), dtype=torch.float32, device=dev) if want_rows else None
        img = None
        if img_out is None and want_img and l1.out_dim % 64 == 0:
            img = operand_image(n, l1.out_dim, dev)
        img_ptr = img_out if img_out is not None else N.ptr(img)
        with torch.cuda.device(dev):
            rc = N.lib().cmhar_mlp2_forward_img(l0.blob.data_ptr(), l1.blob.data_ptr(), x_img.data_ptr(), n, l0.in_dim, l0.out_dim,
                                                l1.out_dim, int(normalize), N.ptr(y), img_ptr, N.stream_ptr(dev))
        if img_out is not None:
            img = img_out
        if rc == N.UNSUPPORTED:
            return None
        N.check(rc)
        return y, img

    def forward(self, x):
        if _native_mode(self):
            return self.forward_native(x)
        return self.net(x)


# =============================================================================== cross-modal model
class CrossModalModel(_PackedMixin, nn.Module):
    """IMU + video encoders, projection heads and L2 normalisation
    (reference src/models/models.py:239-291)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        self.imu_encoder = IMUEncoder(config)
        self.video_encoder = VideoEncoder(config)
        self.imu_proj = ProjectionHead(m.imu_d_model, m.projection_hidden_dim, m.projection_dim)
        self.video_proj = ProjectionHead(m.video_d_model, m.projection_hidden_dim, m.projection_dim)
        self.temperature = nn.Parameter(torch.ones([]) * math.log(10))
        self.bias = nn.Parameter(torch.ones([]) * -10)
        self._init_packed()

    def forward(self, imu, video):
        if _native_mode(self):
            imu_feat = self.imu_encoder.encode_cls(imu)
            video_feat = self.video_encoder(video)
            return (l2_normalize_native(self.imu_proj(imu_feat)),
                    l2_normalize_native(self.video_proj(video_feat)))
        imu_feat, _ = self.imu_encoder(imu)
        video_feat = self.video_encoder(video)
        return (F.normalize(self.imu_proj(imu_feat), dim=1), F.normalize(self.video_proj(video_feat), dim=1))

    @torch.no_grad()
    def embed_from_features(self, imu, fmap, frames: int, precision: Optional[str] = None):
        """Inference entry for pipelines that already hold the trunk's feature maps
        (fmap: (B*frames, F, h, w) bf16/fp32).  Returns unit-norm (imu_proj, video_proj)."""
        imu_feat = self.imu_encoder.encode_cls(imu, precision=precision)
        video_feat = self.video_encoder.forward_features(fmap, frames, precision=precision)
        return (l2_normalize_native(self.imu_proj.forward_native(imu_feat, precision)),
                l2_normalize_native(self.video_proj.forward_native(video_feat, precision)))


# =============================================================================== classifier
def pack_head_blob(classifier: nn.Sequential, device) -> torch.Tensor:
    """Packs a head of the reference layout ([Linear, BN, ReLU, Dropout] x 2, Linear on a 128-d feature,
    src/models/models.py:312-326) into the device blob ``cmhar_head_forward`` / ``cmhar_imu_forward`` take."""
    mods = list(classifier)
    if len(mods) != 9 or mods[0].in_features != 128:
        raise NotImplementedError("native head supports the reference layout: 2 hidden blocks on a 128-d feature")
    lib = N.lib()
    p = N.HeadParams()
    p.hidden1, p.hidden2, p.classes = mods[0].out_features, mods[4].out_features, mods[8].out_features
    keep = []

    def dp(t):
        t = N.f32c(t.detach())
        keep.append(t)
        return t.data_ptr()
    for i, (lin, bn) in enumerate(((mods[0], mods[1]), (mods[4], mods[5]))):
        setattr(p, f"w{i}", dp(lin.weight)); setattr(p, f"b{i}", dp(lin.bias))
        setattr(p, f"bn{i}_weight", dp(bn.weight)); setattr(p, f"bn{i}_bias", dp(bn.bias))
        setattr(p, f"bn{i}_mean", dp(bn.running_mean)); setattr(p, f"bn{i}_var", dp(bn.running_var))
    p.w2, p.b2 = dp(mods[8].weight), dp(mods[8].bias)
    blob = N.alloc_blob(lib.cmhar_head_blob_bytes(p.hidden1, p.hidden2, p.classes), device)
    with torch.cuda.device(device):
        N.check(lib.cmhar_head_pack(C.byref(p), blob.data_ptr(), N.stream_ptr(device)))
        torch.cuda.current_stream(device).synchronize()       # the fp32 staging copies in `keep` may be freed on return
    del keep
    return blob


class IMUClassifier(_PackedMixin, nn.Module):
    """IMU encoder + MLP head (reference src/models/models.py:296-348).

    Inference fuses encoder, head, arg-max and the logit-based OOD scores into one launch
    (``forward_scores``); ``forward`` returns only the logits, as the reference does."""

    def __init__(self, imu_encoder, config, freeze_encoder=False):
        super().__init__()
        self.imu_encoder = imu_encoder
        self.config = config
        m = config.model
        if freeze_encoder:
            for p in self.imu_encoder.parameters():
                p.requires_grad = False
        layers, in_dim = [], m.imu_d_model
        for hidden in m.classifier_hidden_dims:
            layers += [nn.Linear(in_dim, hidden), nn.BatchNorm1d(hidden), nn.ReLU(inplace=True),
                       nn.Dropout(m.classifier_dropout)]
            in_dim = hidden
        layers.append(nn.Linear(in_dim, m.num_classes))
        self.classifier = nn.Sequential(*layers)
        self.num_classes = m.num_classes
        self._maha_state = None
        self._init_packed()

    @property
    def freeze_encoder(self):
        return not next(self.imu_encoder.parameters()).requires_grad

    def unfreeze_encoder(self):
        for p in self.imu_encoder.parameters():
            p.requires_grad = True

    # ------------------------------------------------------------------ packing
    def _head_blob(self, device) -> torch.Tensor:
        key = ("head", str(device))
        if key not in self._packed:
            self._packed[key] = pack_head_blob(self.classifier, device)
        return self._packed[key]

    def set_mahalanobis(self, maha) -> None:
        """Attach a fitted ``ood.MahalanobisOOD`` so ``forward_scores`` also emits its score."""
        self._maha_state = maha
        _PACK_GENERATION[0] += 1          # recorded graphs of this model now compute a different set of outputs

    # ------------------------------------------------------------------ forward
    def forward(self, imu):
        if _native_mode(self):
            N.require_cuda(imu, "IMUClassifier")
            return imu_forward_native(self.imu_encoder, self._head_blob(imu.device), None, imu,
                                      want_logits=True, classes=self.num_classes)["logits"]
        with torch.set_grad_enabled(self.training or not self.freeze_encoder):
            feat, _ = self.imu_encoder(imu)
        return self.classifier(feat)

    @torch.no_grad()
    def forward_scores(self, imu, *, precision: Optional[str] = None, want_cls: bool = False,
                       want_logits: bool = True, window_stride: Optional[int] = None,
                       out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Fused inference: logits, pred (int64 arg-max), msp, energy (and maha when a fitted
        Mahalanobis state is attached; cls when asked) from ONE kernel launch."""
        if self.training:
            raise RuntimeError("forward_scores is an inference entry point: call .eval() first")
        N.require_cuda(imu, "IMUClassifier")
        if (precision or _DEFAULT_PRECISION) == "bf16_refined":
            return self._forward_scores_refined(imu, want_cls, window_stride, out)
        maha_blob = self._maha_state.blob(imu.device) if self._maha_state is not None else None
        return imu_forward_native(self.imu_encoder, self._head_blob(imu.device), maha_blob, imu,
                                  want_cls=want_cls, want_logits=want_logits, want_pred=True, want_msp=True,
                                  want_energy=True, want_maha=maha_blob is not None, classes=self.num_classes,
                                  precision=precision, window_stride=window_stride, out=out)

    #: "bf16_refined": rows whose top-2 logit margin is below REFINE_REL_TAU * max|logit| are re-run on the fp32 path.
    #: The bf16 path's measured logit error is 5e-3 * max|logit| (tests/test_gpu_bf16_contract.py); a row can only change its
    #: arg-max if its margin is below twice the error, so 4e-2 leaves a 4x safety factor.
    REFINE_REL_TAU = 4e-2

    @torch.no_grad()
    def _forward_scores_refined(self, imu, want_cls, window_stride, out):
        """bf16 tensor-core pass over every window, then the fp32 kernel over the near-tie rows only (``cmhar_near_tie_rows``):
        the predicted labels are those of the fp32 path -- the reference's, bit for bit (goldens) -- at close to bf16
        throughput.  Scores of the rows that were not re-run keep bf16 accuracy (AUROC within 2e-3 of the reference; the
        3-decimal AUROC contract is the fp32 path's).  Synchronises once (the number of selected rows is read back)."""
        res = self.forward_scores(imu, precision="bf16", want_cls=want_cls, want_logits=True, window_stride=window_stride, out=out)
        n, dev = res["logits"].shape[0], imu.device
        if n == 0:
            return res
        work = torch.empty(2, dtype=torch.int32, device=dev)
        idx = torch.empty(n, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib().cmhar_near_tie_rows(res["logits"].data_ptr(), n, self.num_classes, float(self.REFINE_REL_TAU),
                                                work.data_ptr(), idx.data_ptr(), N.stream_ptr(dev)))
        k = int(work[1].item())
        res["refined_rows"] = k
        if k == 0:
            return res
        sel = idx[:k]
        sub = imu.index_select(0, sel)                       # data movement only: the selected windows, contiguous
        exact = self.forward_scores(sub, precision="fp32", want_cls=want_cls, want_logits=True, window_stride=window_stride)
        for key, val in exact.items():
            if key in res and isinstance(res[key], torch.Tensor) and res[key].shape[:1] == (n,):
                res[key].index_copy_(0, sel, val)
        return res
