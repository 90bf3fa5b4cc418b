"""Fused encode -> fuse -> OOD-score pass over one batch (BASELINE.json's metric).

``CrossModalOODPipeline.run`` is the call a user makes for a batch of IMU windows plus the video
trunk's feature maps:

    IMU windows --(encoder launch)--> CLS feature
    feature maps --(pool + GEMM)--> video feature
    both --> late-fusion concat-MLP --> head --> logits, arg-max, MSP, energy, Mahalanobis   (configs[1])
    both --> projection heads --> L2 normalise --> B x B similarity with fused sigmoid loss
(without a fusion classifier the IMU-only head of ``IMUClassifier`` is scored instead; without feature maps
the pass is the reference's ``Evaluator.predict`` body).

Every stage is a kernel of ``libcmhar_b200.so``; nothing is computed by torch ops.  ``capture``
records the whole pass into a CUDA graph (one graph launch with two parallel branches), which is what the
throughput benchmark replays.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from . import _native as N
from .losses import similarity_img_native, similarity_img_work, similarity_native
from .fusion import LateFusionClassifier
from .models import CrossModalModel, IMUClassifier, imu_forward_native, l2_normalize_native, _prec_code, pack_generation
from .ood import MahalanobisOOD

__all__ = ["CrossModalOODPipeline"]


class CrossModalOODPipeline:
    def __init__(self, classifier: IMUClassifier, cross_modal: CrossModalModel,
                 mahalanobis: Optional[MahalanobisOOD] = None, frames: int = 16,
                 precision: Optional[str] = None, sigmoid_scale: float = 10.0, sigmoid_bias: float = -10.0,
                 fusion: Optional[LateFusionClassifier] = None):
        """``mahalanobis`` is attached to the classifier that gets scored: the fusion classifier when one is
        given (it must have been fitted on FUSED features), else the IMU-only classifier."""
        self.clf, self.xm, self.frames, self.precision = classifier.eval(), cross_modal.eval(), frames, precision
        self.fusion = fusion.eval() if fusion is not None else None
        if mahalanobis is not None:
            (self.fusion if self.fusion is not None else self.clf).set_mahalanobis(mahalanobis)
        self.sig = (float(sigmoid_scale), float(sigmoid_bias))
        self._host = None
        self.trunk = None           # video_trunk.DeviceVideoTrunk (attach_trunk): frames in, feature maps never leave the device
        self._side = None           # side stream: the video branch runs concurrently with the IMU kernel
        self._fused_ok = True       # bf16: the 7-launch route (fused projection heads / fusion head / similarity on operand images)

    # ------------------------------------------------------------------ frames in, trunk on the device
    def attach_trunk(self, trunk=True, **kwargs):
        """Give the pipeline the video trunk (``video_trunk.DeviceVideoTrunk``: channels-last bf16, CUDA graph; True = build it from
        ``cross_modal.video_encoder.backbone``).  ``run_frames`` / ``run_host`` / ``stream_host`` then accept decoded uint8 frames
        ``(B, T, H, W, 3)`` where they take feature maps: the frames are normalised on the device, the trunk's channels-last map
        is pooled in place, and no feature map crosses PCIe."""
        self.trunk = self.xm.video_encoder.attach_device_trunk(trunk, **kwargs)
        return self.trunk

    def _frames_to_map(self, frames: torch.Tensor, slot: int = 0) -> torch.Tensor:
        trunk = getattr(self, "trunk", None)
        if trunk is None:
            raise RuntimeError("uint8 frames need the device trunk: call pipeline.attach_trunk() first")
        if frames.dim() == 5 and frames.shape[1] != self.frames:
            raise ValueError(f"clips of {frames.shape[1]} frames, pipeline built for {self.frames}")
        return trunk(frames, slot=slot)

    def _frame_buffer(self, shape, dev, slot: int) -> torch.Tensor:
        """Device landing buffer of a uint8 frame batch = the trunk graph's static input of ``slot`` (no second device copy)."""
        if self.trunk is None:
            raise RuntimeError("uint8 frames need the device trunk: call pipeline.attach_trunk() first")
        if len(shape) < 4 or int(shape[-1]) != 3:
            raise ValueError(f"frames must be uint8 (.., H, W, 3); got shape {tuple(shape)}")
        n = 1
        for d in shape[:-3]:
            n *= int(d)
        return self.trunk.to(dev).frame_buffer(n, int(shape[-3]), int(shape[-2]), slot).view(shape)

    @torch.no_grad()
    def run_frames(self, imu: torch.Tensor, frames: torch.Tensor, slot: int = 0, **kwargs) -> Dict[str, torch.Tensor]:
        """``run`` from frames on the device: uint8 ``(B, T, H, W, 3)`` (normalised here) or the reference's float ``(B, T, 3, H, W)``."""
        return self.run(imu, self._frames_to_map(frames, slot), **kwargs)

    def _fused_route(self, fmap: Optional[torch.Tensor], pooled) -> bool:
        ve = self.xm.video_encoder
        return (self._fused_ok and fmap is not None and pooled is None and _prec_code(self.precision) == N.BF16
                and ve.feature_dim % 64 == 0 and ve.projection.out_features % 64 == 0
                and self.clf.imu_encoder.d_model % 64 == 0 and fmap.shape[0] > 0)

    def _run_fused(self, imu, fmap, window_stride, sim_work) -> Optional[Dict[str, torch.Tensor]]:
        """bf16 route, 7 launches: [encoder -> IMU projection head] || [pooling -> video projection -> video projection head],
        then [fusion layer + classifier head + scores] and [similarity + sigmoid loss].  Every hand-over between kernels is a
        bf16 operand image (plain bulk copies on the consumer side); L2 normalisation, the fusion layer, the loss's
        zero-initialisation and its 1/n live inside those kernels.  Returns None if a fused kernel does not serve the dimensions."""
        dev = imu.device
        B = imu.shape[0]
        main = torch.cuda.current_stream(dev)
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        ve = self.xm.video_encoder
        if fmap.shape[0] != B * self.frames:
            raise ValueError(f"{fmap.shape[0]} frames for {B} windows of {self.frames} frames")
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, pimg = ve.pool_features(fmap, self.frames, want_img=True, want_rows=False)
            _, vimg = ve._packed_projection(dev).forward_img(B, False, x_img=pimg, want_rows=False, want_img=True)
            r = self.xm.video_proj.forward_fused(vimg, B)
        if r is None:
            main.wait_stream(side)
            return None
        vp, vp_img = r
        if self.fusion is None:
            scored = self.clf
            maha_blob = scored._maha_state.blob(dev) if scored._maha_state is not None else None
            out = imu_forward_native(self.clf.imu_encoder, scored._head_blob(dev), maha_blob, imu, want_cls=True, want_logits=True,
                                     want_pred=True, want_msp=True, want_energy=True, want_maha=maha_blob is not None,
                                     classes=scored.num_classes, precision=self.precision, window_stride=window_stride,
                                     want_cls_img=True)
        else:
            out = imu_forward_native(self.clf.imu_encoder, None, None, imu, want_cls=True, precision=self.precision,
                                     window_stride=window_stride, want_cls_img=True)
        r = self.xm.imu_proj.forward_fused(out["cls_img"], B)
        main.wait_stream(side)
        if not torch.cuda.is_current_stream_capturing():
            for t in (vimg, vp, vp_img):
                t.record_stream(main)
        if r is None:
            return None
        ip, ip_img = r
        if self.fusion is not None:
            if self.fusion.forward_scores_img(out["cls_img"], vimg, B, out) is None:
                return None
        if sim_work is None:
            sim_work = similarity_img_work(B, B, dev)
        loss = similarity_img_native(ip_img, B, vp_img, B, ip.shape[1], sigmoid=self.sig, work=sim_work)
        out.update(imu_proj=ip, video_proj=vp, loss=loss, _video_img=vimg, _imu_proj_img=ip_img, _video_proj_img=vp_img, _sim_work=sim_work)
        return out

    @torch.no_grad()
    def run(self, imu: torch.Tensor, fmap: Optional[torch.Tensor], window_stride: Optional[int] = None,
            pooled: Optional[torch.Tensor] = None, sim_work: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """imu (B,6,L) fp32 [or compact (B,live) with window_stride]; fmap (B*frames,F,h,w) bf16/fp32
        or None for the IMU-only path; ``pooled`` (B,F) = the feature maps already reduced by
        ``video_encoder.pool_features`` (the HBM-bound stage can then be scheduled separately from the
        tensor-bound ones).  Returns device tensors: pred, msp, energy, (maha,) logits,
        cls and, with fmap, imu_proj, video_proj, loss (mean sigmoid contrastive loss, fp64 0-dim)."""
        if fmap is None and pooled is None:
            if self.fusion is not None:
                raise ValueError("a pipeline with a fusion classifier needs the video feature maps")
            return self.clf.forward_scores(imu, precision=self.precision, want_cls=True, window_stride=window_stride)
        if self._fused_route(fmap, pooled):
            out = self._run_fused(imu, fmap, window_stride, sim_work)
            if out is not None:
                return out
            self._fused_ok = False          # dimensions outside the fused kernels: the chained route below, from now on
        # fork: the HBM-bound video tail + its projection head on a side stream, the tensor-bound IMU
        # kernel + its projection head on the current stream; join before the similarity kernel.
        # (Captured into a CUDA graph this becomes two parallel branches.)
        main = torch.cuda.current_stream(imu.device)
        if self._side is None or self._side.device != imu.device:
            self._side = torch.cuda.Stream(device=imu.device)
        side = self._side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ve, pimg, nclips = self.xm.video_encoder, None, None
            if pooled is None:
                if _prec_code(self.precision) == N.BF16:      # pooled features go straight into a bf16 operand image
                    nclips = fmap.shape[0] // self.frames
                    pooled, pimg = ve.pool_features(fmap, self.frames, want_img=True, want_rows=False)
                else:
                    pooled = ve.pool_features(fmap, self.frames)
            vfeat, vimg = ve.project_pooled(pooled, precision=self.precision, want_img=True, x_img=pimg, n=nclips)
            vp = l2_normalize_native(self.xm.video_proj.forward_native(vfeat, self.precision, x_img=vimg))
        if self.fusion is None:
            out = self.clf.forward_scores(imu, precision=self.precision, want_cls=True, window_stride=window_stride)
        else:
            out = imu_forward_native(self.clf.imu_encoder, None, None, imu, want_cls=True, precision=self.precision,
                                     window_stride=window_stride)
        ip = l2_normalize_native(self.xm.imu_proj.forward_native(out["cls"], self.precision))
        main.wait_stream(side)
        if not torch.cuda.is_current_stream_capturing():
            for t in (vfeat, vp):
                t.record_stream(main)
        if self.fusion is not None:
            out.update(self.fusion.forward_scores(None, None, self.frames, precision=self.precision, imu_cls=out["cls"],
                                                  video_feat=vfeat))
        res = similarity_native(ip, vp, sigmoid=self.sig, precision=self.precision)
        out.update(imu_proj=ip, video_proj=vp, loss=res["sigmoid_sum"] / float(ip.shape[0] * vp.shape[0]))
        return out

    def capture(self, imu: torch.Tensor, fmap: Optional[torch.Tensor]):
        """Record ``run`` on static inputs into a CUDA graph; returns (graph, outputs)."""
        side = torch.cuda.Stream(device=imu.device)
        side.wait_stream(torch.cuda.current_stream(imu.device))
        # the similarity kernel's ticket workspace is zeroed ONCE here, outside the graph (the kernel re-arms it itself)
        work = similarity_img_work(imu.shape[0], imu.shape[0], imu.device) if self._fused_route(fmap, None) else None
        with torch.cuda.stream(side):
            self.run(imu, fmap, sim_work=work)       # warm-up: packs weights, sizes allocations
        torch.cuda.current_stream(imu.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.run(imu, fmap, sim_work=work)
        return graph, out

    # ------------------------------------------------------------------ host-buffer entry (e2e)
    @torch.no_grad()
    def run_host(self, imu_host: torch.Tensor, fmap_host: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Same pass from HOST tensors: stages the live IMU samples (channel 0, 16*(S-1) samples) and
        the feature maps through pinned buffers, runs the fused pass, and copies the per-window
        results (pred, msp, energy, maha) and the loss back to pinned host memory.  Synchronises."""
        dev = next(self.clf.parameters()).device
        B, L = imu_host.shape[0], imu_host.shape[-1]
        live = 16 * (self.clf.imu_encoder._check_native_dims(L) - 1)
        h = self._host
        if h is None or h["B"] != B or h["live"] != live or (fmap_host is not None and h.get("fmap_shape") != (tuple(fmap_host.shape), fmap_host.dtype)):
            h = {"B": B, "live": live,
                 "imu_pin": torch.empty((B, live), dtype=torch.float32).pin_memory(),
                 "imu_dev": torch.empty((B, live), dtype=torch.float32, device=dev),
                 "res_pin": torch.empty((4, B), dtype=torch.float32).pin_memory(),
                 "pred_pin": torch.empty((B,), dtype=torch.int64).pin_memory(),
                 "loss_pin": torch.empty((), dtype=torch.float64).pin_memory()}
            if fmap_host is not None:
                h["fmap_shape"] = (tuple(fmap_host.shape), fmap_host.dtype)
                h["fmap_pin"] = fmap_host if fmap_host.is_pinned() else torch.empty_like(fmap_host).pin_memory()
                h["fmap_dev"] = (self._frame_buffer(tuple(fmap_host.shape), dev, 0) if fmap_host.dtype == torch.uint8
                                 else torch.empty(fmap_host.shape, dtype=fmap_host.dtype, device=dev))
            self._host = h
        h["imu_pin"].copy_(imu_host[:, 0, :live] if imu_host.dim() == 3 else imu_host[:, :live])
        h["imu_dev"].copy_(h["imu_pin"], non_blocking=True)
        fdev = None
        if fmap_host is not None:
            src = fmap_host
            if not fmap_host.is_pinned():
                h["fmap_pin"].copy_(fmap_host)
                src = h["fmap_pin"]
            h["fmap_dev"].copy_(src, non_blocking=True)
            fdev = h["fmap_dev"]
            if fdev.dtype == torch.uint8:                    # decoded frames: normalise + trunk on the device
                fdev = self._frames_to_map(fdev, 0)
        out = self.run(h["imu_dev"], fdev, window_stride=live)
        h["pred_pin"].copy_(out["pred"], non_blocking=True)
        h["res_pin"][0].copy_(out["msp"], non_blocking=True)
        h["res_pin"][1].copy_(out["energy"], non_blocking=True)
        if "maha" in out:
            h["res_pin"][2].copy_(out["maha"], non_blocking=True)
        if "loss" in out:
            h["loss_pin"].copy_(out["loss"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        res = {"pred": h["pred_pin"], "msp": h["res_pin"][0], "energy": h["res_pin"][1]}
        if "maha" in out:
            res["maha"] = h["res_pin"][2]
        if "loss" in out:
            res["loss"] = h["loss_pin"]
        return res

    @torch.no_grad()
    def stream_host(self, batches: Iterable, depth: int = 2, graphs: bool = True):
        """Streaming form of ``run_host`` for a sequence of HOST batches ``(imu_host, fmap_host)``: the
        host->device copy of batch i+1 (copy engine, its own stream) overlaps the kernels of batch i, the
        per-window results of every batch are copied back to pinned memory, and the generator yields them in
        order -- what an evaluator loop over a DataLoader does (reference src/eval/evaluator.py:40-49, with
        the per-batch ``.cpu()`` sync replaced by a ``depth``-deep ring).  Every batch still pays its own
        H2D and D2H; only their latency is hidden.

        IMU-only batches (``fmap_host is None``, the reference's ``Evaluator.predict`` workload) are small -- 246 KB in,
        5 KB out, two kernel launches -- and host-bound when every copy and launch is issued from Python (~100 us per
        batch).  With ``graphs`` each ring slot records its [H2D, encoder + head + scores, D2H] sequence once as a CUDA
        graph on its own stream and replays it per batch: one host call instead of eight."""
        dev = next(self.clf.parameters()).device
        main = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(device=dev)
        # ring slots (pinned / device buffers, events, recorded graphs) persist across calls: recording a slot's graph costs
        # milliseconds, a short evaluation must not pay it again
        if not hasattr(self, "_stream_slots"):
            self._stream_slots = {}
        slots: List[dict] = self._stream_slots.setdefault((str(dev), depth), [])
        pending: List[dict] = []

        def finish(sl):
            # the slot's pinned buffers are overwritten when the slot is re-enqueued `depth` batches later: hand out OWNED
            # copies (5 KB per batch), so `list(pipe.stream_host(...))` and consumers that hold results stay correct
            sl["done"].synchronize()
            return {k: v.clone() for k, v in sl["res"].items()}

        try:
            yield from self._stream_host_loop(batches, depth, graphs, dev, main, copy, slots, pending, finish)
        finally:
            # generator closed early (or raised): work of the handed-in batches may still be in flight on the copy / lane
            # streams -- wait for it before the caller can re-key the slots or free what they point to
            for sl in pending:
                sl["done"].synchronize()

    def _stream_host_loop(self, batches, depth, graphs, dev, main, copy, slots, pending, finish):
        for i, (imu_host, fmap_host) in enumerate(batches):
            B, L = imu_host.shape[0], imu_host.shape[-1]
            live = 16 * (self.clf.imu_encoder._check_native_dims(L) - 1)
            if len(slots) < depth:
                slots.append({"key": None})
            sl = slots[i % depth]
            if len(pending) == depth:                       # the slot about to be reused must have been handed out
                yield finish(pending.pop(0))
            key = (B, live, None if fmap_host is None else (tuple(fmap_host.shape), fmap_host.dtype))
            if sl["key"] != key:
                if sl.get("done") is not None and sl.get("used_any"):
                    sl["done"].synchronize()               # the buffers about to be dropped may still be read by earlier work
                sl.update(key=key, graph=None, used=False, imu_pin=torch.empty((B, live), dtype=torch.float32).pin_memory(),
                          imu_dev=torch.empty((B, live), dtype=torch.float32, device=dev),
                          res_pin=torch.empty((4, B), dtype=torch.float32).pin_memory(),
                          pred_pin=torch.empty((B,), dtype=torch.int64).pin_memory(),
                          loss_pin=torch.empty((), dtype=torch.float64).pin_memory(),
                          up=torch.cuda.Event(), done=torch.cuda.Event())
                if fmap_host is not None:
                    sl["fmap_pin"] = torch.empty_like(fmap_host).pin_memory()
                    sl["fmap_dev"] = (self._frame_buffer(tuple(fmap_host.shape), dev, i % depth) if fmap_host.dtype == torch.uint8
                                      else torch.empty(fmap_host.shape, dtype=fmap_host.dtype, device=dev))
            sl["imu_pin"].copy_(imu_host[:, 0, :live] if imu_host.dim() == 3 else imu_host[:, :live])
            if fmap_host is None and graphs:
                if sl.get("graph") is None or sl.get("generation") != pack_generation():      # weights re-packed: re-record
                    lane = torch.cuda.Stream(device=dev)
                    lane.wait_stream(main)
                    with torch.cuda.stream(lane):                # warm-up outside capture: packs weights, sizes allocations
                        self.run(sl["imu_dev"], None, window_stride=live)
                    lane.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=lane):
                        sl["imu_dev"].copy_(sl["imu_pin"], non_blocking=True)
                        out = self.run(sl["imu_dev"], None, window_stride=live)
                        sl["pred_pin"].copy_(out["pred"], non_blocking=True)
                        sl["res_pin"][0].copy_(out["msp"], non_blocking=True)
                        sl["res_pin"][1].copy_(out["energy"], non_blocking=True)
                        if "maha" in out:
                            sl["res_pin"][2].copy_(out["maha"], non_blocking=True)
                    sl.update(graph=g, lane=lane, has_maha="maha" in out, generation=pack_generation())
                with torch.cuda.stream(sl["lane"]):
                    sl["graph"].replay()
                    sl["done"].record(sl["lane"])
                sl["used_any"] = True
                res = {"pred": sl["pred_pin"], "msp": sl["res_pin"][0], "energy": sl["res_pin"][1]}
                if sl["has_maha"]:
                    res["maha"] = sl["res_pin"][2]
                sl["res"] = res
                pending.append(sl)
                continue
            if sl.get("used"):
                # the slot's device buffers were last read by batch i - depth: wait for THAT batch only (its `done`
                # event), not for the whole compute stream -- waiting on the stream serialised copy(i) behind
                # compute(i-1) and left the PCIe link idle for the length of a step (46 GB/s instead of 55)
                copy.wait_event(sl["done"])
            sl["used"] = sl["used_any"] = True
            with torch.cuda.stream(copy):
                sl["imu_dev"].copy_(sl["imu_pin"], non_blocking=True)
                fdev = None
                if fmap_host is not None:
                    src = fmap_host
                    if not fmap_host.is_pinned():
                        sl["fmap_pin"].copy_(fmap_host)
                        src = sl["fmap_pin"]
                    sl["fmap_dev"].copy_(src, non_blocking=True)
                    fdev = sl["fmap_dev"]
                sl["up"].record(copy)
            main.wait_event(sl["up"])
            if fdev is not None and fdev.dtype == torch.uint8:       # decoded frames: normalise + trunk graph of this ring slot
                fdev = self._frames_to_map(fdev, i % depth)
            out = self.run(sl["imu_dev"], fdev, window_stride=live)
            sl["pred_pin"].copy_(out["pred"], non_blocking=True)
            sl["res_pin"][0].copy_(out["msp"], non_blocking=True)
            sl["res_pin"][1].copy_(out["energy"], non_blocking=True)
            res = {"pred": sl["pred_pin"], "msp": sl["res_pin"][0], "energy": sl["res_pin"][1]}
            if "maha" in out:
                sl["res_pin"][2].copy_(out["maha"], non_blocking=True)
                res["maha"] = sl["res_pin"][2]
            if "loss" in out:
                sl["loss_pin"].copy_(out["loss"], non_blocking=True)
                res["loss"] = sl["loss_pin"]
            sl["done"].record(main)
            sl["res"] = res
            pending.append(sl)
        while pending:
            yield finish(pending.pop(0))

    def host_bytes_per_step(self, B: int, L: int, fmap_host: Optional[torch.Tensor]):
        live = 16 * (self.clf.imu_encoder._check_native_dims(L) - 1)
        h2d = B * live * 4 + (fmap_host.numel() * fmap_host.element_size() if fmap_host is not None else 0)
        scored = self.fusion if self.fusion is not None else self.clf
        d2h = B * (8 + 4 + 4 + (4 if scored._maha_state is not None else 0)) + (8 if fmap_host is not None else 0)
        return h2d, d2h
