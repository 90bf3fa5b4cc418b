"""GPU tests of the fused encode -> fuse -> score pass (pipeline.py): eager, host-buffer and
CUDA-graph routes must agree with each other and with the oracle."""
import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from oracle import ood_spec, oracle, weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tsd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def build(seed=31):
    cfg = cm.default_config()
    sd_x = W.cross_modal_state(seed)
    sd_c = W.classifier_state(seed)              # same encoder seed -> same imu_encoder.* tensors
    clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
    clf.load_state_dict(tsd(sd_c), strict=True)
    xm = cm.CrossModalModel(cfg)
    xm.load_state_dict(tsd(sd_x), strict=True)
    return clf.to(DEV).eval(), xm.to(DEV).eval(), sd_c, sd_x


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_pipeline_routes_agree_with_oracle(precision, tol):
    clf, xm, sd_c, sd_x = build()
    B, T = 40, 16
    imu = W.imu_windows(5, B)
    fmap = W.video_feature_maps(6, B, T)
    feats, labels = W.class_features(7, 2000)
    maha = cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV))
    pipe = cm.CrossModalOODPipeline(clf, xm, maha, frames=T, precision=precision)
    x, f = torch.from_numpy(imu).to(DEV), torch.from_numpy(fmap).to(DEV).to(torch.bfloat16)
    eager = pipe.run(x, f)
    torch.cuda.synchronize()
    # oracle (float64 on the same bf16-rounded feature maps)
    f_r = f.float().cpu().numpy()
    want_logits, want_cls = oracle.imu_classifier(imu, sd_c, dtype=torch.float64)
    ip, vp = oracle.cross_modal(imu, f_r, sd_x, T, dtype=torch.float64)
    want_loss = float(oracle.sigmoid_contrastive_loss(ip, vp, dtype=torch.float64))
    rel = lambda g, w: float(np.abs(g.detach().cpu().numpy().astype(np.float64) - w.numpy()).max() / np.abs(w.numpy()).max())
    assert rel(eager["logits"], want_logits) < tol
    assert rel(eager["imu_proj"], ip) < tol and rel(eager["video_proj"], vp) < tol
    assert abs(float(eager["loss"]) - want_loss) < tol * want_loss
    spec = ood_spec.mahalanobis_finalize(*ood_spec.mahalanobis_sufficient_stats(feats, labels, 32), ridge=1e-3)
    want_maha = ood_spec.mahalanobis_score(want_cls.numpy(), spec)
    assert float(np.abs(eager["maha"].cpu().numpy() - want_maha).max() / np.abs(want_maha).max()) < max(tol, 5e-3) * 3
    if precision == "fp32":
        assert np.array_equal(eager["pred"].cpu().numpy(), oracle.predict(want_logits))
    # host-buffer route
    host = pipe.run_host(torch.from_numpy(imu), f.cpu())
    assert torch.equal(host["pred"], eager["pred"].cpu())
    assert torch.allclose(host["energy"], eager["energy"].cpu()) and torch.allclose(host["maha"], eager["maha"].cpu())
    assert abs(float(host["loss"]) - float(eager["loss"])) < 1e-12
    # CUDA-graph route (two parallel branches)
    graph, out = pipe.capture(x, f)
    for v in out.values():
        if v.dtype.is_floating_point:
            v.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["pred"], eager["pred"]) and torch.equal(out["logits"], eager["logits"])
    assert abs(float(out["loss"]) - float(eager["loss"])) < 1e-12


def test_late_fusion_pipeline_matches_spec():
    """configs[1]: the pass with the (spec-defined) late-fusion classifier scored instead of the IMU-only head."""
    from oracle import fusion_spec
    cfg = cm.default_config()
    sd = fusion_spec.fusion_state(41)
    sd_x = W.cross_modal_state(41)
    xm = cm.CrossModalModel(cfg)
    xm.load_state_dict(tsd(sd_x), strict=True)
    clf = cm.IMUClassifier(xm.imu_encoder, cfg)
    fus = cm.LateFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg)
    own = set(fus.state_dict().keys())
    fus.load_state_dict({k: v for k, v in tsd(sd).items() if k in own}, strict=True)
    xm, clf, fus = xm.to(DEV).eval(), clf.to(DEV).eval(), fus.to(DEV).eval()
    B, T = 40, 16
    imu = W.imu_windows(5, B)
    f = torch.from_numpy(W.video_feature_maps(6, B, T)).to(DEV).to(torch.bfloat16)
    pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=T, precision="fp32", fusion=fus)
    out = pipe.run(torch.from_numpy(imu).to(DEV), f)
    torch.cuda.synchronize()
    want, want_f = fusion_spec.late_fusion(imu, f.float().cpu().numpy(), sd, T, dtype=torch.float64)
    rel = lambda g, w: float(np.abs(g.detach().cpu().numpy().astype(np.float64) - w.numpy()).max() / np.abs(w.numpy()).max())
    assert rel(out["logits"], want) < 1e-3 and rel(out["fused"], want_f) < 1e-3
    assert np.array_equal(out["pred"].cpu().numpy(), oracle.predict(want))
    ip, vp = oracle.cross_modal(imu, f.float().cpu().numpy(), sd_x, T, dtype=torch.float64)
    assert abs(float(out["loss"]) - float(oracle.sigmoid_contrastive_loss(ip, vp, dtype=torch.float64))) < 1e-3 * float(out["loss"])
    graph, gout = pipe.capture(torch.from_numpy(imu).to(DEV), f)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(gout["logits"], out["logits"]) and torch.equal(gout["pred"], out["pred"])


def test_stream_host_equals_run_host():
    clf, xm, _, _ = build()
    B, T = 24, 16
    pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=T, precision="fp32")
    batches = [(torch.from_numpy(W.imu_windows(50 + i, B)), torch.from_numpy(W.video_feature_maps(60 + i, B, T)).to(torch.bfloat16))
               for i in range(5)]
    want = []
    for imu, f in batches:
        r = pipe.run_host(imu, f)
        want.append({k: v.clone() for k, v in r.items()})
    got = [{k: v.clone() for k, v in r.items()} for r in pipe.stream_host(iter(batches), depth=2)]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert torch.equal(g["pred"], w["pred"]) and torch.equal(g["energy"], w["energy"])
        assert abs(float(g["loss"]) - float(w["loss"])) < 1e-12


def test_pooling_taken_out_of_the_step_and_streamed_batches():
    """``run(pooled=...)`` with the feature maps reduced beforehand -- by the default pooling kernel or by the
    co-resident ring kernel -- gives the same pass as ``run(imu, fmap)``; ``stream_host`` (copy ring that waits on the
    slot's own event) returns every batch's results in order and equal to the synchronous ``run_host``."""
    clf, xm, sd_c, sd_x = build()
    B, T = 48, 16
    x = torch.from_numpy(W.imu_windows(8, B)).to(DEV)
    f = torch.from_numpy(W.video_feature_maps(9, B, T)).to(DEV).to(torch.bfloat16)
    pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=T, precision="bf16")
    ref = pipe.run(x, f)
    for coresident in (False, True):
        pooled = xm.video_encoder.pool_features(f, T, coresident=coresident)
        got = pipe.run(x, None, pooled=pooled)
        torch.cuda.synchronize()
        assert torch.allclose(got["video_proj"], ref["video_proj"], rtol=0, atol=2e-6)
        assert torch.equal(got["pred"], ref["pred"]) and abs(float(got["loss"]) - float(ref["loss"])) < 1e-6
    # streamed host batches: 5 different batches through a 2-deep ring
    batches = [(torch.from_numpy(W.imu_windows(20 + i, B)), torch.from_numpy(W.video_feature_maps(30 + i, B, T)).to(torch.bfloat16)) for i in range(5)]
    sync = [{k: v.clone() for k, v in pipe.run_host(a, b).items()} for a, b in batches]
    streamed = [{k: v.clone() for k, v in r.items()} for r in pipe.stream_host(iter(batches), depth=2)]
    assert len(streamed) == len(sync)
    for s, w in zip(streamed, sync):
        assert torch.equal(s["pred"], w["pred"]) and torch.equal(s["energy"], w["energy"]) and float(s["loss"]) == float(w["loss"])


def test_local_mahalanobis_fit_never_enters_a_collective():
    """``fit(all_reduce=False)`` is what rank-0-only code (bench.py's scoring block) must use: same statistics as the
    default on a single process, and no torch.distributed call."""
    feats, labels = W.class_features(3, 3000)
    f, y = torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV)
    a = cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(f, y)
    b = cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(f, y, all_reduce=False)
    np.testing.assert_array_equal(a.fit_["whiten"], b.fit_["whiten"])
    q = torch.from_numpy(W.class_features(4, 777, ood_fraction=0.3)[0]).to(DEV)
    assert torch.equal(a.score(q, precision="bf16"), b.score(q, precision="bf16"))


@pytest.mark.parametrize("n", [5, 128, 300])
def test_projection_head_operand_image_chain_is_bit_identical(n):
    """bf16 path: the hidden activation handed from layer to layer as a bf16 operand image (cmhar_linear_forward_img)
    gives exactly the bits of the fp32-row hand-off -- both round the activation to bf16 once -- for both heads
    (K = 128 and K = 768), ragged and multi-tile row counts, and for the video feature handed over as an image."""
    clf, xm, sd_c, sd_x = build()
    rs = np.random.RandomState(n)
    for head, k in ((xm.imu_proj, 128), (xm.video_proj, 768)):
        x = torch.from_numpy(rs.standard_normal((n, k)).astype(np.float32)).to(DEV)
        l0, l1 = head._packed_layers(x.device)
        rows = l1(l0(x, relu=True, precision="bf16"), relu=False, precision="bf16")
        chained = head.forward_native(x, "bf16")
        torch.cuda.synchronize()
        assert torch.equal(rows, chained)
    pooled = torch.from_numpy(np.maximum(rs.standard_normal((n, 512)), 0).astype(np.float32)).to(DEV)
    vfeat, vimg = xm.video_encoder.project_pooled(pooled, precision="bf16", want_img=True)
    assert torch.equal(vfeat, xm.video_encoder.project_pooled(pooled, precision="bf16"))
    assert torch.equal(xm.video_proj.forward_native(vfeat, "bf16", x_img=vimg), xm.video_proj.forward_native(vfeat, "bf16"))


@pytest.mark.parametrize("n", [3, 129, 256])
def test_pooling_into_operand_image_matches_row_path(n):
    """cmhar_video_pool_img: the pooled features written directly as a bf16 operand image feed the projection layer with
    the same bits as the fp32-row route (the staging path rounds the same fp32 value to bf16)."""
    clf, xm, sd_c, sd_x = build()
    T = 16
    f = torch.from_numpy(W.video_feature_maps(40 + n, n, T)).to(DEV).to(torch.bfloat16)
    ve = xm.video_encoder
    rows = ve.pool_features(f, T)
    both_rows, img = ve.pool_features(f, T, want_img=True)
    none_rows, img2 = ve.pool_features(f, T, want_img=True, want_rows=False)
    torch.cuda.synchronize()
    assert none_rows is None and torch.equal(rows, both_rows)
    want = ve.project_pooled(rows, precision="bf16")
    assert torch.equal(ve.project_pooled(None, precision="bf16", x_img=img, n=n), want)
    assert torch.equal(ve.project_pooled(None, precision="bf16", x_img=img2, n=n), want)


def test_streamed_imu_only_batches_through_slot_graphs():
    """IMU-only ``stream_host``: every ring slot replays one CUDA graph [H2D, encoder + head + scores, D2H]; results in
    order and equal to the synchronous ``run_host`` and to the un-graphed streaming route, also when the batch shape
    changes mid-stream (the slot re-records its graph) and with a Mahalanobis scorer attached."""
    clf, xm, sd_c, sd_x = build()
    feats, labels = W.class_features(7, 2000)
    maha = cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV))
    pipe = cm.CrossModalOODPipeline(clf, xm, maha, frames=16, precision="bf16")
    sizes = [64, 64, 64, 64, 200, 200, 200, 64, 7]
    batches = [(torch.from_numpy(W.imu_windows(50 + i, b)), None) for i, b in enumerate(sizes)]
    want = [{k: v.clone() for k, v in pipe.run_host(a, None).items()} for a, _ in batches]
    for graphs in (True, False):
        got = [{k: v.clone() for k, v in r.items()} for r in pipe.stream_host(iter(batches), depth=2, graphs=graphs)]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert set(g) == set(w) and all(torch.equal(g[k], w[k]) for k in w)


def test_slot_graphs_follow_reloaded_weights():
    """The recorded per-slot graphs hold raw pointers into the packed weight blobs: after ``load_state_dict`` (which drops
    the blobs) the next ``stream_host`` call must re-record and return the NEW model's results."""
    clf, xm, sd_c, sd_x = build(31)
    pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=16, precision="bf16")
    batches = [(torch.from_numpy(W.imu_windows(70 + i, 64)), None) for i in range(4)]
    first = [r["energy"].clone() for r in pipe.stream_host(iter(batches))]
    clf.load_state_dict(tsd(W.classifier_state(32)), strict=True)
    want = [pipe.run_host(a, None)["energy"].clone() for a, _ in batches]
    got = [r["energy"].clone() for r in pipe.stream_host(iter(batches))]
    assert all(torch.equal(g, w) for g, w in zip(got, want))
    assert not all(torch.equal(g, f) for g, f in zip(got, first))


def test_stream_host_results_are_owned_and_early_close_is_safe():
    """``list(pipe.stream_host(...))`` WITHOUT cloning: every yielded result is an owned copy, not a view of the ring slot's
    pinned buffers (which later batches overwrite).  Closing the generator early waits for the work in flight; a following
    call with another batch shape re-keys the slots and still returns correct results."""
    clf, xm, sd_c, sd_x = build()
    pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=16, precision="bf16")
    batches = [(torch.from_numpy(W.imu_windows(90 + i, 64)), None) for i in range(7)]
    want = [{k: v.clone() for k, v in pipe.run_host(a, None).items()} for a, _ in batches]
    for graphs in (True, False):
        got = list(pipe.stream_host(iter(batches), depth=2, graphs=graphs))           # no clone on purpose
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert all(torch.equal(g[k], w[k]) for k in w)
    gen = pipe.stream_host(iter(batches), depth=2)
    first = next(gen)
    gen.close()                                                                        # abandons two batches in flight
    assert torch.equal(first["pred"], want[0]["pred"])
    other = [(torch.from_numpy(W.imu_windows(120 + i, 96)), None) for i in range(3)]
    want2 = [{k: v.clone() for k, v in pipe.run_host(a, None).items()} for a, _ in other]
    got2 = list(pipe.stream_host(iter(other), depth=2))
    for g, w in zip(got2, want2):
        assert all(torch.equal(g[k], w[k]) for k in w)


def test_ood_sweep_matches_spec_on_held_out_activity_split():
    """``OODSweep`` (configs[4]): fit on the ID rows only, one fused launch per batch, scores device resident, histogram
    AUROC / FPR95 -- against the float64 spec computed from the same path's own outputs, to 3 decimals; held-out rows never
    reach the fit (count check)."""
    from oracle import ood_spec
    clf, xm, sd_c, sd_x = build()
    held = [3, 17, 30]
    rs = np.random.RandomState(5)
    xs = [torch.from_numpy(W.imu_windows(200 + i, 500)).to(DEV) for i in range(4)]
    ys = [torch.from_numpy(rs.randint(0, 32, size=500).astype(np.int64)).to(DEV) for _ in range(4)]
    sw = cm.OODSweep(clf, held, precision="fp32", ridge=1e-3)
    maha = sw.fit(zip(xs[:2], ys[:2]))
    y_fit = torch.cat(ys[:2]).cpu().numpy()
    assert int(maha.fit_["count"].sum()) == int((~np.isin(y_fit, held)).sum())
    assert all(maha.fit_["count"][h] == 0 for h in held)
    sw.score(zip(xs[2:], ys[2:]))
    table = sw.metrics()
    res = [clf.forward_scores(x, precision="fp32") for x in xs[2:]]
    y = torch.cat(ys[2:]).cpu().numpy()
    ood = np.isin(y, held)
    for k in ("msp", "energy", "maha"):
        s = torch.cat([r[k] for r in res]).cpu().numpy()
        assert round(table[k]["auroc"], 3) == round(ood_spec.auroc(s[~ood], s[ood]), 3)
        assert round(table[k]["fpr95"], 3) == round(ood_spec.fpr_at_tpr_fast(s[~ood], s[ood]), 3)
    pred = torch.cat([r["pred"] for r in res]).cpu().numpy()
    assert abs(sw.accuracy - 100.0 * (pred[~ood] == y[~ood]).mean()) < 1e-9
    clf.set_mahalanobis(None)
