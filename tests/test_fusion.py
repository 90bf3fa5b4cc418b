"""Fusion classifiers (SURVEY.md section 8a row A6 -- SPEC-DEFINED, the reference has no fusion block):
late-fusion concat-MLP and cross-attention fusion against the in-repo spec ``oracle/fusion_spec.py``.
CPU: the autograd route; GPU: the native kernels through the C ABI.  Self-consistency, not reference parity."""
import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from oracle import fusion_spec, ood_spec, oracle, weights as W

T = 16


def tsd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def build(kind, seed=41):
    cfg = cm.default_config()
    sd = fusion_spec.fusion_state(seed)
    cls = cm.LateFusionClassifier if kind == "late" else cm.CrossAttentionFusionClassifier
    model = cls(cm.IMUEncoder(cfg), cm.VideoEncoder(cfg), cfg)
    own = set(model.state_dict().keys())
    missing = own - set(sd.keys())
    assert not missing, missing
    model.load_state_dict({k: v for k, v in tsd(sd).items() if k in own}, strict=True)
    return model.eval(), sd


def inputs(B, seed=3):
    return W.imu_windows(seed, B), W.video_feature_maps(seed + 1, B, T)


@pytest.mark.parametrize("kind", ["late", "xattn"])
def test_autograd_route_matches_spec(kind):
    model, sd = build(kind)
    imu, fmap = inputs(6)
    spec = fusion_spec.late_fusion if kind == "late" else fusion_spec.cross_attention_fusion
    want, _ = spec(imu, fmap, sd, T)
    video = torch.from_numpy(fmap).view(6, T, *fmap.shape[1:])          # identity trunk: "video" = feature maps
    got = model(torch.from_numpy(imu), video)
    assert got.requires_grad
    np.testing.assert_allclose(got.detach().numpy(), want.numpy(), atol=2e-4, rtol=0)
    got.sum().backward()
    assert model.imu_encoder.patch_embed.projections[0].weight.grad.abs().sum() > 0


def test_native_route_refuses_cpu_tensors():
    model, _ = build("late")
    imu, fmap = inputs(2)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.from_numpy(imu), torch.from_numpy(fmap).view(2, T, *fmap.shape[1:]))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["late", "xattn"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_native_fusion_matches_spec(kind, precision, tol):
    model, sd = build(kind)
    model = model.to("cuda:0")
    B = 77
    imu, fmap = inputs(B, 9)
    f_dev = torch.from_numpy(fmap).to("cuda:0").to(torch.bfloat16)
    f_r = f_dev.float().cpu().numpy()                       # the oracle sees the same bf16-rounded feature maps
    spec = fusion_spec.late_fusion if kind == "late" else fusion_spec.cross_attention_fusion
    want, want_f = spec(imu, f_r, sd, T, dtype=torch.float64)
    feats, labels = W.class_features(7, 2000)
    maha = cm.MahalanobisOOD(32, "cuda:0", ridge=1e-3).fit(torch.from_numpy(feats).cuda(), torch.from_numpy(labels).cuda())
    model.set_mahalanobis(maha)
    out = model.forward_scores(torch.from_numpy(imu).cuda(), f_dev, T, precision=precision)
    torch.cuda.synchronize()
    rel = lambda g, w: float(np.abs(g.detach().cpu().numpy().astype(np.float64) - w.numpy()).max() / np.abs(w.numpy()).max())
    assert rel(out["fused"], want_f) < tol
    assert rel(out["logits"], want) < tol
    got_logits = out["logits"].cpu().numpy()
    # scores are functions of the logits the kernel produced
    np.testing.assert_array_equal(out["pred"].cpu().numpy(), got_logits.argmax(1))
    np.testing.assert_allclose(out["energy"].cpu().numpy(), ood_spec.energy_score(got_logits), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["msp"].cpu().numpy(), ood_spec.msp_score(got_logits), rtol=1e-4, atol=1e-5)
    if precision == "fp32":
        assert np.array_equal(out["pred"].cpu().numpy(), oracle.predict(want))
        st = ood_spec.mahalanobis_finalize(*ood_spec.mahalanobis_sufficient_stats(feats, labels, 32), ridge=1e-3)
        want_m = ood_spec.mahalanobis_score(want_f.numpy(), st)
        assert float(np.abs(out["maha"].cpu().numpy() - want_m).max() / np.abs(want_m).max()) < 5e-3
    # the module's forward (video through the identity trunk) returns the same logits
    with torch.no_grad():
        cm.set_default_precision(precision)
        try:
            video = f_dev.view(B, T, *f_dev.shape[1:])
            logits = model(torch.from_numpy(imu).cuda(), video)
        finally:
            cm.set_default_precision("fp32")
    assert torch.equal(logits, out["logits"])


# ------------------------------------------------------------------ conv / BN / ReLU encoder (spec-defined)
def build_conv(seed=51):
    cfg = cm.default_config()
    sd = fusion_spec.conv_encoder_state(seed)
    model = cm.ConvIMUClassifier(cfg)
    model.load_state_dict(tsd(sd), strict=True)
    return model.eval(), sd


def test_conv_encoder_autograd_route_matches_spec():
    model, sd = build_conv()
    x = W.imu_windows(4, 5)
    want, _ = fusion_spec.conv_classifier(x, sd)
    got = model(torch.from_numpy(x))
    assert got.requires_grad
    np.testing.assert_allclose(got.detach().numpy(), want.numpy(), atol=2e-4, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("L,B", [(250, 77), (100, 33), (256, 300)])
def test_native_conv_encoder_matches_spec(L, B):
    model, sd = build_conv()
    model = model.to("cuda:0")
    x = W.imu_windows(6, B, W.Dims(imu_window=L))
    want, want_f = fusion_spec.conv_classifier(x, sd, dtype=torch.float64)
    out = model.forward_scores(torch.from_numpy(x).cuda(), precision="fp32")
    torch.cuda.synchronize()
    rel = lambda g, w: float(np.abs(g.detach().cpu().numpy().astype(np.float64) - w.numpy()).max() / np.abs(w.numpy()).max())
    assert rel(out["cls"], want_f) < 1e-5
    assert rel(out["logits"], want) < 1e-4
    assert np.array_equal(out["pred"].cpu().numpy(), oracle.predict(want))
    out16 = model.forward_scores(torch.from_numpy(x).cuda(), precision="bf16")      # tensor-core conv stack + tensor-core head
    assert rel(out16["cls"], want_f) < 2e-2 and rel(out16["logits"], want) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("L,B", [(250, 1), (250, 2), (250, 77), (100, 33), (16, 5), (254, 40), (255, 9), (256, 300), (250, 2 * 148 * 3 + 5)])
def test_tensor_core_conv_encoder_matches_spec(L, B):
    """conv_encoder_tc_kernel (implicit GEMMs over shared-memory im2col tiles, bf16 operands, split-precision input) against
    the float64 spec within the bf16 contract (2e-2; measured a few 1e-3) for every window length class -- one and two tiles of
    positions, the 255 / 256-sample right-padding corner, odd window counts (an empty second slot), several pairs per CTA --
    and against the fp32 CUDA-core kernel; a window's result does not depend on its slot or batch neighbours."""
    model, sd = build_conv()
    model = model.to("cuda:0")
    x = W.imu_windows(7, B, W.Dims(imu_window=L))
    xd = torch.from_numpy(x).cuda()
    _, want_f = fusion_spec.conv_classifier(x, sd, dtype=torch.float64)
    got = model.encoder.forward_native(xd, precision="bf16")
    ref32 = model.encoder.forward_native(xd, precision="fp32")
    torch.cuda.synchronize()
    w = want_f.numpy()
    err = float(np.abs(got.cpu().numpy().astype(np.float64) - w).max() / np.abs(w).max())
    print(f"L={L} B={B}: tensor-core conv encoder rel err {err:.2e}")
    assert torch.isfinite(got).all() and err < 2e-2
    assert float((got - ref32).abs().max() / ref32.abs().max()) < 2e-2
    again = model.encoder.forward_native(xd, precision="bf16")
    assert torch.equal(got, again)
    if B >= 5:
        part = model.encoder.forward_native(xd[3:5].contiguous(), precision="bf16")
        assert torch.equal(part, got[3:5])


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,dtype", [(5, 16, torch.bfloat16), (130, 16, torch.bfloat16), (9, 3, torch.float32)])
def test_single_pass_pooling_emits_clip_and_frame_images(B, T, dtype):
    """``pool_features_frames``: one read of the feature maps gives the clip means (temporal + spatial) AND the per-frame
    spatial means, both as bf16 operand images; each equals the fp32 mean rounded to bf16 (images decoded on the host)."""
    import test_gpu_fused_tail as tail
    cfg = cm.default_config()
    ve = cm.CrossModalModel(cfg).video_encoder.to("cuda:0").eval()
    g = torch.Generator().manual_seed(B * 31 + T)
    fmap = torch.relu(torch.randn(B * T, 512, 4, 4, generator=g)).to(dtype).cuda()
    clip_img, frame_img = ve.pool_features_frames(fmap, T)
    torch.cuda.synchronize()
    f32 = fmap.float()
    want_frame = f32.mean(dim=(2, 3))                                   # (B*T, 512)
    want_clip = want_frame.view(B, T, 512).mean(1)
    got_frame = tail.rows_of(frame_img, B * T, 512)
    got_clip = tail.rows_of(clip_img, B, 512)
    assert float((got_frame - want_frame).abs().max()) <= 2 ** -8 * float(want_frame.abs().max())
    assert float((got_clip - want_clip).abs().max()) <= 2 ** -8 * float(want_clip.abs().max())
    # the clip image carries the same values as the rows of the classic entry point
    pooled, img2 = ve.pool_features(fmap, T, want_img=True)
    assert torch.equal(tail.rows_of(img2, B, 512), got_clip)


@pytest.mark.gpu
def test_cross_attention_bf16_route_folds_projection_into_kv_and_refolds_after_reload():
    """bf16 ``forward_scores`` = single-pass pooling + the projection folded into the kv GEMM: within the bf16 contract of the
    float64 spec, and the fold follows a reload of the VIDEO encoder's weights (its cache is keyed by the pack generation)."""
    cfg = cm.default_config()
    sd = fusion_spec.fusion_state(61)
    xm = cm.CrossModalModel(cfg)
    xm.load_state_dict(tsd(W.cross_modal_state(61)), strict=True)
    model = cm.CrossAttentionFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg)
    own = set(model.state_dict().keys())
    model.load_state_dict({k: v for k, v in tsd(sd).items() if k in own}, strict=False)
    model = model.to("cuda:0").eval()
    B, T = 40, 16
    imu, fmap = W.imu_windows(3, B), W.video_feature_maps(4, B, T)
    f_dev = torch.from_numpy(fmap).cuda().to(torch.bfloat16)
    x = torch.from_numpy(imu).cuda()
    got = model.forward_scores(x, f_dev, T, precision="bf16")
    ref = model.forward_scores(x, f_dev, T, precision="fp32")
    torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(got["fused"], ref["fused"]) < 2e-2 and rel(got["logits"], ref["logits"]) < 2e-2
    # reload the video encoder's projection with other weights: the folded kv image must follow
    with torch.no_grad():
        new_sd = {k: v.clone() for k, v in model.video_encoder.state_dict().items()}
        new_sd["projection.weight"] = new_sd["projection.weight"] * 0.5
    model.video_encoder.load_state_dict(new_sd, strict=True)
    got2 = model.forward_scores(x, f_dev, T, precision="bf16")
    ref2 = model.forward_scores(x, f_dev, T, precision="fp32")
    assert rel(got2["fused"], ref2["fused"]) < 2e-2
    assert not torch.equal(got2["fused"], got["fused"])


@pytest.mark.gpu
@pytest.mark.parametrize("L,B", [(250, 1), (250, 7), (250, 8), (250, 9), (250, 300), (100, 45), (250, 8 * 148 + 3)])
def test_fused_cross_attention_kernel_matches_the_chained_route_and_the_spec(L, B):
    """``xattn_tc_kernel`` (q / k / v projections, 8-head attention with lane-masked MMAs, out-projection on top of the residual,
    LayerNorm, token mean -- one launch) against the chained bf16 route (linear_tc + cross_attention + residual_ln_pool kernels) and
    the float64 spec: ragged window counts, a short token sequence (S = 7, padded rows) and several tiles per CTA."""
    cfg = cm.default_config(imu_window_size=L)
    sd = fusion_spec.fusion_state(71, W.Dims(imu_window=L)) if L != 250 else fusion_spec.fusion_state(71)
    xm = cm.CrossModalModel(cfg)
    xm.load_state_dict(tsd(W.cross_modal_state(71, W.Dims(imu_window=L)) if L != 250 else W.cross_modal_state(71)), strict=True)
    model = cm.CrossAttentionFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg)
    own = set(model.state_dict().keys())
    model.load_state_dict({k: v for k, v in tsd(sd).items() if k in own}, strict=False)
    model = model.to("cuda:0").eval()
    T = 16
    imu = W.imu_windows(3, B, W.Dims(imu_window=L))
    f_dev = torch.from_numpy(W.video_feature_maps(4, B, T)).cuda().to(torch.bfloat16)
    x = torch.from_numpy(imu).cuda()
    from crossmodal_imu_video_ood_har_b200.models import imu_forward_native
    tokens = imu_forward_native(model.imu_encoder, None, None, x, want_tokens=True, precision="bf16")["tokens"]
    _, frame_img = model.video_encoder.pool_features_frames(f_dev, T, want_clip_img=False)
    fused = model.fuse_native_img(tokens, frame_img, T)
    chain = model.fuse_native_img(tokens, frame_img, T, fused_kernel=False)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(f"L={L} B={B}: fused cross-attention kernel vs chained route {rel(fused, chain):.2e}")
    assert torch.isfinite(fused).all() and rel(fused, chain) < 1e-2
    again = model.fuse_native_img(tokens, frame_img, T)
    assert torch.equal(fused, again)
    ref = model.forward_scores(x, f_dev, T, precision="fp32")["fused"]
    assert rel(fused, ref) < 2e-2
