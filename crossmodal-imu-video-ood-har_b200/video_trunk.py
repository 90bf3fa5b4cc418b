"""The video trunk on the device: channels-last, bf16, BatchNorm folded, one CUDA graph per input shape.

The reference runs its CNN trunk eagerly in fp32 NCHW (``self.backbone(x)``, src/models/models.py:163-173,209) on frames the
DataLoader normalised on the CPU (``ToTensor`` + ``Normalize``, src/data/datasets.py:52-58).  The trunk is third-party library code
(torchvision modules, cuDNN kernels) and stays that here -- none of it is claimed as a hand-written kernel.  What this module adds
around it is the part of the path this repo owns:

* decoded ``uint8`` HWC frames are normalised ON the device by ``cmhar_frames_normalize`` (a clip crosses PCIe as 602 KB of bytes
  instead of 2.4 MB of fp32), straight into the trunk's static channels-last input buffer (input channels zero-padded to 4: 8-byte
  pixels; measured as fast as 3 and 22 % faster than 8 through cuDNN's first convolution);
* Conv2d + BatchNorm2d pairs are folded (eval-mode algebra, fp32) before the cast to bf16; for resnet trunks bias, ReLU and the
  residual add ride in the convolutions' epilogues (cuDNN's fused conv-bias-activation launches: two launches per BasicBlock instead
  of five kernels); the whole trunk is captured once per input shape in a CUDA graph (no Python or launch overhead per layer);
* the trunk's channels-last output is consumed as it lies by ``cmhar_video_pool_nhwc`` (``VideoEncoder.pool_features``): the
  feature map is never permuted, never copied and never leaves the device.

``DeviceVideoTrunk(video_encoder)`` wraps the trunk of a ``VideoEncoder``; ``VideoEncoder.attach_device_trunk`` routes the
module's own eval-mode ``forward`` through it, ``CrossModalOODPipeline.attach_trunk`` the pipeline's frame entry points.
Precision: bf16 activations and weights, fp32 accumulation (cuDNN) -- the bf16 contract (2e-2 normwise), not the fp32 one; the
eager fp32 trunk remains the default and the fp32-contract path.
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _native as N

__all__ = ["DeviceVideoTrunk", "fold_conv_bn", "IMAGENET_MEAN", "IMAGENET_STD"]

IMAGENET_MEAN = (0.485, 0.456, 0.406)          # reference src/data/datasets.py:56-57
IMAGENET_STD = (0.229, 0.224, 0.225)


def fold_conv_bn(module: nn.Module) -> int:
    """Fold every BatchNorm2d that directly follows a Conv2d among the children of one container (torchvision's BasicBlock,
    Conv2dNormActivation, downsample Sequential, the trunk's own Sequential) into that convolution, in place; eval-mode
    algebra in the parameters' own precision.  Returns the number of folded pairs."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    folded = 0
    for parent in list(module.modules()):
        names = list(parent._modules.keys())
        for a, b in zip(names, names[1:]):
            conv, bn = parent._modules[a], parent._modules[b]
            if isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and bn.track_running_stats and conv.out_channels == bn.num_features:
                parent._modules[a] = fuse_conv_bn_eval(conv.eval(), bn.eval())
                parent._modules[b] = nn.Identity()
                folded += 1
    return folded


class _ConvReLU(nn.Module):
    """Conv2d (BatchNorm folded) + ReLU as ONE cuDNN launch (fused bias + activation epilogue)."""

    def __init__(self, conv: nn.Conv2d):
        super().__init__()
        self.conv = conv

    def forward(self, x):
        c = self.conv
        if x.is_cuda and c.padding_mode == "zeros":
            return torch.cudnn_convolution_relu(x, c.weight, c.bias, c.stride, c.padding, c.dilation, c.groups)
        return torch.relu(c(x))


class _FusedBasicBlock(nn.Module):
    """torchvision ``BasicBlock`` with folded BatchNorm as two cuDNN launches instead of five kernels:
    conv + bias + ReLU, then conv + bias + identity + ReLU (the residual add and both activations ride in the
    convolutions' epilogues: the early, activation-bandwidth-bound stages of the trunk read and write ~half the bytes)."""

    def __init__(self, blk: nn.Module):
        super().__init__()
        self.conv1, self.conv2 = blk.conv1, blk.conv2
        self.down = blk.downsample[0] if blk.downsample is not None else None

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        a, b = self.conv1, self.conv2
        if x.is_cuda:
            y = torch.cudnn_convolution_relu(x, a.weight, a.bias, a.stride, a.padding, a.dilation, a.groups)
            return torch.cudnn_convolution_add_relu(y, b.weight, idt, 1.0, b.bias, b.stride, b.padding, b.dilation, b.groups)
        return torch.relu(b(torch.relu(a(x))) + idt)


def fuse_epilogues(net: nn.Module) -> int:
    """After ``fold_conv_bn``: rewrite torchvision BasicBlocks and a [Conv2d, Identity, ReLU] stem into the fused modules above,
    in place.  Returns the number of rewritten modules (0 for trunks of other shapes, e.g. mobilenet_v2: left as they are)."""
    try:
        from torchvision.models.resnet import BasicBlock
    except Exception:                                          # torchvision absent: nothing to rewrite
        return 0
    n = 0
    for parent in list(net.modules()):
        for name, child in list(parent._modules.items()):
            if (isinstance(child, BasicBlock) and isinstance(child.conv1, nn.Conv2d) and child.conv1.bias is not None
                    and isinstance(child.bn1, nn.Identity) and isinstance(child.bn2, nn.Identity) and child.conv2.bias is not None
                    and (child.downsample is None or (isinstance(child.downsample[0], nn.Conv2d) and isinstance(child.downsample[1], nn.Identity)))):
                parent._modules[name] = _FusedBasicBlock(child)
                n += 1
    if isinstance(net, nn.Sequential) and len(net) >= 3:
        m = list(net._modules.items())
        (k0, c), (_, i), (k2, r) = m[0], m[1], m[2]
        if isinstance(c, nn.Conv2d) and c.bias is not None and isinstance(i, nn.Identity) and isinstance(r, nn.ReLU):
            net._modules[k0] = _ConvReLU(c)
            net._modules[k2] = nn.Identity()
            n += 1
    return n


def _first_conv(module: nn.Module) -> Tuple[nn.Module, str]:
    for parent in module.modules():
        for name, child in parent._modules.items():
            if isinstance(child, nn.Conv2d):
                return parent, name
    raise ValueError("DeviceVideoTrunk: the trunk has no Conv2d")


class DeviceVideoTrunk:
    """``trunk(frames)`` -> channels-last bf16 feature map (B*T, F, h, w) on the device.

    ``frames``: uint8 ``(B, T, H, W, 3)`` / ``(N, H, W, 3)`` decoded frames (normalised here), or the reference's float layout
    ``(B, T, 3, H, W)`` / ``(N, 3, H, W)``, already normalised (converted to channels-last bf16 by one torch copy).
    The returned map is the graph's static output buffer of ``slot``: it is overwritten by the next call with the same
    shape and slot (a copy ring of depth d uses slots 0..d-1)."""

    def __init__(self, video_encoder_or_backbone: nn.Module, *, pad_in_channels: int = 4, fold_bn: bool = True,
                 fused_epilogues: bool = True, graphs: bool = True, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD):
        src = getattr(video_encoder_or_backbone, "backbone", video_encoder_or_backbone)
        if getattr(video_encoder_or_backbone, "is_videomae", False):
            raise NotImplementedError("DeviceVideoTrunk serves the per-frame CNN trunks (resnet18 / mobilenet_v2)")
        if pad_in_channels not in (3, 4, 8):
            raise ValueError("pad_in_channels must be 3, 4 or 8")
        self.cpad = int(pad_in_channels)
        self._src, self._fold_bn = src, bool(fold_bn)
        self.fused_error: Optional[str] = None
        self._build(bool(fused_epilogues))
        self.graphs = bool(graphs)
        self._mean = (C.c_float * 3)(*[float(v) for v in mean])
        self._std = (C.c_float * 3)(*[float(v) for v in std])
        self._mean_t, self._std_t = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self._device: Optional[torch.device] = None
        self._slots: Dict[tuple, dict] = {}

    def _build(self, fused_epilogues: bool) -> None:
        net = copy.deepcopy(self._src).float().eval()
        for p in net.parameters():
            p.requires_grad_(False)
        self.folded = fold_conv_bn(net) if self._fold_bn else 0
        parent, name = _first_conv(net)
        conv = parent._modules[name]
        if conv.in_channels != 3 or conv.groups != 1:
            raise ValueError(f"DeviceVideoTrunk: the first convolution takes {conv.in_channels} channels, not RGB")
        if self.cpad != 3:                                    # zero input channels 3..cpad-1: same sums, tensor-core shaped K
            wide = nn.Conv2d(self.cpad, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, 1,
                             conv.bias is not None, conv.padding_mode)
            with torch.no_grad():
                wide.weight.zero_()
                wide.weight[:, :3].copy_(conv.weight)
                if conv.bias is not None:
                    wide.bias.copy_(conv.bias)
            parent._modules[name] = wide.eval()
        # bias / ReLU / residual add in the convolutions' epilogues (cuDNN fused launches); needs the folded BatchNorms
        self.fused = fuse_epilogues(net) if (self._fold_bn and fused_epilogues) else 0
        self.net = net
        self._device = None

    # ------------------------------------------------------------------ placement
    def to(self, device) -> "DeviceVideoTrunk":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DeviceVideoTrunk runs only on a CUDA device; the eager module is the CPU / training path")
        if self._device != device:
            self.net = self.net.to(device=device, dtype=torch.bfloat16, memory_format=torch.channels_last)
            if self.fused:
                # cuDNN's fused conv-bias-activation launches are a library feature: if this build / device refuses them for
                # channels-last bf16, the trunk is rebuilt from plain convolutions (recorded in ``fused_error``) -- a choice
                # between two LIBRARY routes of third-party code, not a fallback of the hand-written path
                try:
                    with torch.no_grad():
                        self.net(torch.zeros((1, 32, 32, self.cpad), dtype=torch.bfloat16, device=device).permute(0, 3, 1, 2))
                    torch.cuda.synchronize(device)
                except Exception as e:                          # pragma: no cover  (depends on the cuDNN build)
                    self.fused_error = f"{type(e).__name__}: {e}"
                    self._build(False)
                    self.net = self.net.to(device=device, dtype=torch.bfloat16, memory_format=torch.channels_last)
            self._device, self._slots = device, {}
        return self

    # ------------------------------------------------------------------ buffers
    def _slot(self, n: int, h: int, w: int, slot: int) -> dict:
        key = (n, h, w, slot)
        sl = self._slots.get(key)
        if sl is None:
            dev = self._device
            # physical (n, h, w, cpad) == logical (n, cpad, h, w) in channels_last
            x = torch.zeros((n, h, w, self.cpad), dtype=torch.bfloat16, device=dev).permute(0, 3, 1, 2)
            sl = {"x": x, "u8": None, "graph": None, "y": None}
            self._slots[key] = sl
        return sl

    def frame_buffer(self, n: int, h: int, w: int, slot: int = 0) -> torch.Tensor:
        """The slot's static uint8 (n, h, w, 3) device buffer: an H2D copy can land in it directly (no extra device copy)."""
        sl = self._slot(n, h, w, slot)
        if sl["u8"] is None:
            sl["u8"] = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self._device)
        return sl["u8"]

    def normalize_into(self, frames_u8: torch.Tensor, x: torch.Tensor) -> None:
        """uint8 (n, h, w, 3) on the device -> the channels-last bf16 input ``x`` (n, cpad, h, w): cmhar_frames_normalize."""
        if frames_u8.dtype != torch.uint8 or frames_u8.shape[-1] != 3 or not x.is_contiguous(memory_format=torch.channels_last) and x.shape[-1] != self.cpad:
            raise ValueError("normalize_into: uint8 (.., H, W, 3) frames into a channels-last (n, cpad, H, W) bf16 buffer")
        frames_u8 = frames_u8.contiguous()
        with torch.cuda.device(frames_u8.device):
            N.check(N.lib().cmhar_frames_normalize(frames_u8.data_ptr(), frames_u8.numel() // 3, self._mean, self._std, self.cpad,
                                                   x.data_ptr(), N.stream_ptr(frames_u8.device)))

    # ------------------------------------------------------------------ call
    @torch.no_grad()
    def __call__(self, frames: torch.Tensor, slot: int = 0) -> torch.Tensor:
        N.require_cuda(frames, "DeviceVideoTrunk")
        if self._device != frames.device:
            self.to(frames.device)
        is_u8 = frames.dtype == torch.uint8
        if frames.dim() == 5:
            frames = frames.reshape(frames.shape[0] * frames.shape[1], *frames.shape[2:])
        if frames.dim() != 4 or (is_u8 and frames.shape[3] != 3) or (not is_u8 and frames.shape[1] != 3):
            raise ValueError("DeviceVideoTrunk: frames must be uint8 (.., H, W, 3) or float (.., 3, H, W); got "
                             f"{tuple(frames.shape)} {frames.dtype}")
        n = frames.shape[0]
        h, w = (frames.shape[1], frames.shape[2]) if is_u8 else (frames.shape[2], frames.shape[3])
        sl = self._slot(n, h, w, slot)
        dev = frames.device
        if is_u8:
            if sl["u8"] is None or frames.data_ptr() != sl["u8"].data_ptr():
                self.frame_buffer(n, h, w, slot).copy_(frames, non_blocking=True)
        else:
            sl["x"][:, :3].copy_(frames)                        # NCHW float -> channels-last bf16 (channels >= 3 stay zero)
        if not self.graphs:
            if is_u8:
                self.normalize_into(sl["u8"], sl["x"])
            return self.net(sl["x"])
        mode = "u8" if is_u8 else "f"
        if sl["graph"] is None or sl.get("mode") != mode:
            main = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(main)
            old = torch.backends.cudnn.benchmark
            torch.backends.cudnn.benchmark = True              # algorithms are chosen during the warm-up, outside the capture
            try:
                with torch.cuda.stream(side):
                    for _ in range(2):
                        if is_u8:
                            self.normalize_into(sl["u8"], sl["x"])
                        self.net(sl["x"])
                side.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    if is_u8:
                        self.normalize_into(sl["u8"], sl["x"])
                    y = self.net(sl["x"])
            finally:
                torch.backends.cudnn.benchmark = old
            main.wait_stream(side)
            if not y.is_contiguous(memory_format=torch.channels_last):
                raise RuntimeError("DeviceVideoTrunk: the trunk's output is not channels-last")
            sl.update(graph=g, y=y, mode=mode)
        sl["graph"].replay()
        return sl["y"]

    def reference_normalize(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """torch restatement of the transform (tests): uint8 (n, h, w, 3) -> fp32 (n, 3, h, w)."""
        x = frames_u8.permute(0, 3, 1, 2).float() / 255.0
        mean = torch.tensor(self._mean_t, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(self._std_t, device=x.device).view(1, 3, 1, 1)
        return (x - mean) / std
