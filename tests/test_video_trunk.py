"""Video trunk on the device (video_trunk.py).  CPU part: the host algebra (BatchNorm folding, zero-padded input
channels) against the unmodified torch trunk.  GPU part (-m gpu): the two kernels either side of the trunk against torch, the
channels-last route against the NCHW kernels bit for bit, and the frames-in pipeline against the feature-map pipeline.

  cmhar_frames_normalize   reference src/data/datasets.py:52-58 (ToTensor + Normalize), bit-exact vs torch in fp32 -> bf16
  cmhar_video_pool_nhwc    reference src/models/models.py:210-211,215 on the channels-last map
"""
import copy

import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from crossmodal_imu_video_ood_har_b200.models import operand_image

DEV = "cuda:0"
N = cm._native


def rows_of(img: torch.Tensor, n: int, dim: int) -> torch.Tensor:
    """bf16 operand image -> (n, dim) fp32 rows (host-side restatement of the SWIZZLE_128B chunk layout)."""
    tiles, kcs = (n + 127) // 128, dim // 64
    flat = img[: tiles * kcs * 16384].view(torch.bfloat16).view(tiles, kcs, 128 * 64)
    r = torch.arange(128, device=img.device)
    out = torch.zeros(tiles, 128, kcs, 64, dtype=torch.float32, device=img.device)
    for j in range(8):
        src = (r // 8) * 512 + (r % 8) * 64 + ((j ^ (r % 8)) * 8)
        for e in range(8):
            out[:, :, :, j * 8 + e] = flat[:, :, src + e].transpose(1, 2).float()
    return out.reshape(tiles * 128, dim)[:n]


def video_encoder(backbone="resnet18", seed=3):
    cfg = cm.default_config()
    cfg.model.video_backbone, cfg.model.video_pretrained = backbone, False
    torch.manual_seed(seed)
    ve = cm.VideoEncoder(cfg).eval()
    g = torch.Generator().manual_seed(seed)
    for m in ve.backbone.modules():                         # default BN statistics (0, 1) would make the folding vacuous
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(1.0 + 0.2 * torch.randn(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return ve


@pytest.mark.parametrize("backbone,cpad", [("resnet18", 8), ("resnet18", 3), ("mobilenet_v2", 4)])
def test_folded_padded_trunk_equals_torch_trunk_fp32(backbone, cpad):
    ve = video_encoder(backbone)
    trunk = cm.DeviceVideoTrunk(ve, pad_in_channels=cpad)
    assert trunk.folded >= 20
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in trunk.net.modules())
    x = torch.randn(3, 3, 48, 48, generator=torch.Generator().manual_seed(1))
    xp = torch.zeros(3, cpad, 48, 48)
    xp[:, :3] = x
    with torch.no_grad():
        want, got = ve.backbone(x), trunk.net(xp)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-5 * want.abs().max().item()
    # the module's own weights are untouched (the trunk works on a copy)
    assert any(isinstance(m, torch.nn.BatchNorm2d) for m in ve.backbone.modules())


def test_device_trunk_refuses_cpu_and_videomae_and_is_not_deep_copied():
    ve = video_encoder()
    trunk = ve.attach_device_trunk(True)
    with pytest.raises(RuntimeError, match="CUDA"):
        trunk(torch.zeros(1, 2, 16, 16, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="CUDA"):
        trunk.to("cpu")
    ve2 = copy.deepcopy(ve)
    assert ve2._device_trunk is None and ve2._device_trunk_kwargs == {}
    assert ve2._trunk() is not None and ve2._trunk() is not trunk        # rebuilt from the copy's own weights on demand
    ve.load_state_dict(ve.state_dict())                                   # parameters may have changed: snapshot is stale
    assert ve._device_trunk_stale and ve._trunk() is not trunk
    ve.attach_device_trunk(None)
    assert ve._trunk() is None


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("cpad", [3, 4, 8])
@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (1, 1, 7, 5), (1, 2, 112, 112)])
def test_frames_normalize_bit_exact(cpad, shape):
    B, T, H, W = shape
    g = torch.Generator().manual_seed(H * W + cpad)
    u8 = torch.randint(0, 256, (B * T, H, W, 3), dtype=torch.uint8, generator=g).to(DEV)
    trunk = cm.DeviceVideoTrunk(video_encoder(), pad_in_channels=cpad).to(DEV)
    out = torch.full((B * T, H, W, cpad), 7.0, dtype=torch.bfloat16, device=DEV)
    trunk.normalize_into(u8, out)
    torch.cuda.synchronize()
    want = trunk.reference_normalize(u8).to(torch.bfloat16).permute(0, 2, 3, 1)
    assert torch.equal(out[..., :3], want)
    assert cpad == 3 or float(out[..., 3:].abs().max()) == 0.0


@pytest.mark.gpu
def test_frames_normalize_rejects_bad_arguments_and_accepts_empty():
    lib = N.lib()
    import ctypes as C
    m, s = (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(1, 1, 1)
    buf = torch.zeros(64, dtype=torch.uint8, device=DEV)
    assert lib.cmhar_frames_normalize(buf.data_ptr(), 0, m, s, 8, buf.data_ptr(), N.stream_ptr(DEV)) == 0
    assert lib.cmhar_frames_normalize(buf.data_ptr(), 4, m, s, 5, buf.data_ptr(), N.stream_ptr(DEV)) != 0
    z = (C.c_float * 3)(1, 0, 1)
    assert lib.cmhar_frames_normalize(buf.data_ptr(), 4, m, z, 8, buf.data_ptr(), N.stream_ptr(DEV)) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,T,F,h,w", [(5, 16, 512, 4, 4), (3, 4, 1280, 2, 2), (130, 2, 64, 3, 3), (2, 7, 128, 1, 5)])
def test_pool_nhwc_matches_nchw_kernels(dtype, B, T, F, h, w):
    g = torch.Generator().manual_seed(B + T + F)
    fmap = torch.relu(torch.randn(B * T, F, h, w, generator=g)).to(dtype).to(DEV)
    cl = fmap.contiguous(memory_format=torch.channels_last)
    assert not cl.is_contiguous() or h * w == 1
    ve = video_encoder().to(DEV)
    want_rows, want_img = ve.pool_features(fmap, T, want_img=True)
    got_rows, got_img = ve.pool_features(cl, T, want_img=True)
    torch.cuda.synchronize()
    ref = fmap.float().view(B, T, F, h * w).mean(dim=(1, 3))
    assert (got_rows - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item())
    assert (got_rows - want_rows).abs().max().item() <= 2e-6
    # operand images (rows past n are never written): the bf16 rounding of the same fp32 means, up to the reduction order
    near = lambda x, y, n: (rows_of(x, n, F) - rows_of(y, n, F)).abs().max().item() <= 2 ** -7 * max(1e-6, rows_of(y, n, F).abs().max().item())
    assert near(got_img, want_img, B)
    assert (rows_of(got_img, B, F) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    clip_img, frame_img = ve.pool_features_frames(cl, T)
    clip_ref, frame_ref = ve.pool_features_frames(fmap, T)
    torch.cuda.synchronize()
    assert near(clip_img, clip_ref, B) and near(frame_img, frame_ref, B * T)
    fref = fmap.float().view(B * T, F, h * w).mean(dim=2)
    assert (rows_of(frame_img, B * T, F) - fref).abs().max().item() <= 2 ** -7 * fref.abs().max().item()
    if F == ve.feature_dim:                                    # the projection that follows is Linear(feature_dim -> video_d_model)
        per_frame = ve.forward_frame_features(cl)
        per_frame_ref = ve.forward_frame_features(fmap)
        assert (per_frame - per_frame_ref).abs().max().item() <= 1e-5 * max(1.0, per_frame_ref.abs().max().item())


@pytest.mark.gpu
def test_pool_nhwc_empty_and_unsupported():
    lib = N.lib()
    buf = torch.zeros(4096, dtype=torch.uint8, device=DEV)
    out = torch.zeros(64, device=DEV)
    st = N.stream_ptr(DEV)
    assert lib.cmhar_video_pool_nhwc(buf.data_ptr(), 1, 0, 16, 512, 16, out.data_ptr(), None, None, st) == 0
    assert lib.cmhar_video_pool_nhwc(buf.data_ptr(), 1, 1, 2, 12, 4, out.data_ptr(), None, None, st) != 0          # channels % 8
    assert lib.cmhar_video_pool_nhwc(buf.data_ptr(), 1, 1, 2, 32, 4, None, out.data_ptr(), None, st) != 0          # image needs % 64
    assert lib.cmhar_video_pool_nhwc(buf.data_ptr(), 1, 1, 2, 32, 4, None, None, None, st) != 0                    # no output


@pytest.mark.gpu
@pytest.mark.parametrize("backbone", ["resnet18", "mobilenet_v2"])
def test_device_trunk_matches_eager_fp32_trunk(backbone):
    """uint8 frames -> normalise -> channels-last bf16 trunk (CUDA graph) -> pooled features, against the eager fp32 module on the
    torch-normalised frames: the bf16 contract (2e-2 normwise), graph replay == eager bf16 run, second call reuses the graph."""
    ve = video_encoder(backbone).to(DEV)
    B, T, H = 3, 4, 64
    g = torch.Generator().manual_seed(9)
    u8 = torch.randint(0, 256, (B, T, H, H, 3), dtype=torch.uint8, generator=g).to(DEV)
    trunk = ve.attach_device_trunk(True)
    x32 = trunk.reference_normalize(u8.view(B * T, H, H, 3))
    with torch.no_grad():
        fmap_ref = ve.backbone(x32)
        want = ve.projection(fmap_ref.mean(dim=(2, 3)).view(B, T, -1)).mean(dim=1)
        got = ve(u8)                                                  # module forward, eval + no_grad: device trunk route
        fmap = trunk(u8).float()
        again = ve(u8)
        ungraphed = cm.DeviceVideoTrunk(ve, graphs=False).to(DEV)(u8).float()
    torch.cuda.synchronize()
    assert tuple(fmap.shape) == tuple(fmap_ref.shape)
    rel = lambda a, b: (a - b).abs().max().item() / b.abs().max().item()
    assert rel(fmap, fmap_ref) < 2e-2, rel(fmap, fmap_ref)
    assert rel(got, want) < 2e-2, rel(got, want)
    assert torch.equal(got, again)
    assert rel(ungraphed, fmap) < 1e-2          # cuDNN may pick another algorithm outside the capture: same contract, not same bits
    # the reference's float layout (B, T, 3, H, W), already normalised, takes the same route
    with torch.no_grad():
        got_f = ve(x32.view(B, T, 3, H, H))
    assert rel(got_f, want) < 2e-2


@pytest.mark.gpu
def test_pipeline_from_frames_matches_pipeline_from_maps():
    """CrossModalOODPipeline: run_frames / stream_host on uint8 frames == run on the trunk's own feature map, and every
    per-window output within the bf16 contract of the fp32-trunk + oracle-free fp32 pipeline."""
    from oracle import fusion_spec, weights as W
    tsd = lambda sd: {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    cfg = cm.default_config()
    cfg.model.video_backbone, cfg.model.video_pretrained = "resnet18", False
    torch.manual_seed(0)
    xm = cm.CrossModalModel(cfg)
    missing = xm.load_state_dict(tsd(W.cross_modal_state(41)), strict=False)
    assert all(k.startswith("video_encoder.backbone.") for k in missing.missing_keys) and not missing.unexpected_keys
    clf = cm.IMUClassifier(xm.imu_encoder, cfg)
    fus = cm.LateFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg)
    own = set(fus.state_dict().keys())
    fus.load_state_dict({k: v for k, v in tsd(fusion_spec.fusion_state(41)).items() if k in own}, strict=False)
    xm, clf, fus = xm.to(DEV).eval(), clf.to(DEV).eval(), fus.to(DEV).eval()
    B, T, H = 6, 16, 64
    pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=T, precision="bf16", fusion=fus)
    trunk = pipe.attach_trunk(True)
    imu = torch.from_numpy(W.imu_windows(5, B)).to(DEV)
    g = torch.Generator().manual_seed(2)
    batches = [torch.randint(0, 256, (B, T, H, H, 3), dtype=torch.uint8, generator=g) for _ in range(3)]
    outs = list(pipe.stream_host([(imu.cpu(), f) for f in batches], depth=2))
    assert len(outs) == 3
    for f, o in zip(batches, outs):
        direct = pipe.run_frames(imu, f.to(DEV), slot=1)
        fmap = trunk(f.to(DEV), slot=1)
        via_map = pipe.run(imu, fmap)
        torch.cuda.synchronize()
        for k in ("pred", "msp", "energy"):
            assert torch.equal(direct[k], via_map[k])
            assert torch.equal(o[k].to(DEV), direct[k].to(o[k].dtype)), k
        assert abs(float(o["loss"]) - float(direct["loss"])) < 1e-9
    assert len({int(o["pred"][0]) for o in outs} | {0}) >= 1
    hb = pipe.host_bytes_per_step(B, 250, batches[0])
    assert hb[0] == B * 240 * 4 + B * T * H * H * 3


def test_pipeline_frames_need_an_attached_trunk():
    """Host logic only: the frame entry points fail loudly before any device work when no trunk is attached / the layout is wrong."""
    cfg = cm.default_config()
    clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
    pipe = cm.CrossModalOODPipeline(clf, cm.CrossModalModel(cfg), None, frames=16)
    with pytest.raises(RuntimeError, match="attach_trunk"):
        pipe._frame_buffer((2, 16, 32, 32, 3), "cuda:0", 0)
    with pytest.raises(RuntimeError, match="attach_trunk"):
        pipe._frames_to_map(torch.zeros(2, 16, 32, 32, 3, dtype=torch.uint8))
    pipe.trunk = object()
    with pytest.raises(ValueError, match="H, W, 3"):
        pipe._frame_buffer((2, 16, 3, 32, 32), "cuda:0", 0)


@pytest.mark.parametrize("backbone,folded,fused", [("resnet18", 20, 9), ("mobilenet_v2", 52, 0)])
def test_video_encoder_with_cnn_trunk_vs_reference_golden(golden_dir, backbone, folded, fused):
    """Row a4 WITH the reference's trunks (src/models/models.py:163-173,208-216): the golden holds the output of the UNMODIFIED reference
    ``VideoEncoder`` constructed under a fixed torch seed (oracle/make_golden.py: video_encoder_with_trunk).  This module built under the
    same seed has the same parameters; its torch route (what the reference's trainers use) must give the reference's output, and the
    BatchNorm-folded, channel-padded, epilogue-fused trunk that DeviceVideoTrunk runs on the device must reproduce the reference
    trunk's map in fp32 (the bf16 device run is held to the eager trunk in the GPU test above)."""
    import os
    g = np.load(os.path.join(golden_dir, f"video_encoder_{backbone}.npz"))
    cfg = cm.default_config()
    cfg.model.video_backbone, cfg.model.video_pretrained = backbone, False
    torch.manual_seed(int(g["seed_init"]))
    ve = cm.VideoEncoder(cfg).eval()
    rs = np.random.RandomState(int(g["seed_bn"]))
    for m in ve.backbone.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.from_numpy((0.2 * rs.standard_normal(m.num_features)).astype(np.float32)))
            m.running_var.copy_(torch.from_numpy((0.5 + rs.rand(m.num_features)).astype(np.float32)))
            m.weight.data.copy_(torch.from_numpy((1.0 + 0.2 * rs.standard_normal(m.num_features)).astype(np.float32)))
            m.bias.data.copy_(torch.from_numpy((0.1 * rs.standard_normal(m.num_features)).astype(np.float32)))
    B, T, H = int(g["B"]), int(g["T"]), int(g["H"])
    x = torch.from_numpy(np.random.RandomState(int(g["seed_x"])).standard_normal((B, T, 3, H, H)).astype(np.float32))
    out = ve(x).detach()                                            # grad enabled: the differentiable torch route, on the CPU
    want = torch.from_numpy(g["out"])
    assert out.shape == want.shape and (out - want).abs().max().item() <= 2e-5 * want.abs().max().item()
    trunk = cm.DeviceVideoTrunk(ve)                                 # fp32 on the CPU until .to(cuda): the algebra, not the kernels
    assert trunk.folded == folded and trunk.fused == fused
    xp = torch.zeros(B * T, trunk.cpad, H, H)
    xp[:, :3] = x.view(B * T, 3, H, H)
    with torch.no_grad():
        fmap = trunk.net(xp)
    assert tuple(fmap.shape) == tuple(int(v) for v in g["fmap_shape"])
    fm = torch.from_numpy(g["frame_means"])
    assert (fmap.mean(dim=(2, 3)) - fm).abs().max().item() <= 5e-5 * fm.abs().max().item()
    # ... and the tail on that map (the reference's pool -> projection -> temporal mean, models.py:210-215)
    tail = ve.projection(fmap.mean(dim=(2, 3)).view(B, T, -1)).mean(dim=1).detach()
    assert (tail - want).abs().max().item() <= 5e-5 * want.abs().max().item()
