"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI on cuda:0, against
(1) golden vectors recorded from the unmodified reference, (2) the CPU oracle on seeded inputs,
(3) size-independent properties at BASELINE.json's full sizes.

Tolerance (BASELINE.json north_star): logits / embeddings within 1e-3 relative in fp32 and 2e-2 in
bf16, where "relative" is normwise: max|got - want| / max|want| over the tensor (an elementwise
ratio is meaningless on near-zero logits, SURVEY.md section 7 "Tolerance definition").
Predicted labels: bit-exact.  OOD AUROC / FPR95: identical to 3 decimals.
"""
import os

import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from oracle import ood_spec, oracle, weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"fp32": 1e-3, "bf16": 2e-2}


def rel_err(got, want):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else np.asarray(want)
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


def tsd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def make_classifier(seed, L=250):
    cfg = cm.default_config(imu_window_size=L)
    clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
    sd = W.classifier_state(seed, W.Dims(imu_window=L))
    clf.load_state_dict(tsd(sd), strict=True)
    return clf.to(DEV).eval(), sd


def bf16_available(clf):
    try:
        clf.forward_scores(torch.zeros(1, 6, 250, device=DEV), precision="bf16")
        return True
    except RuntimeError as e:
        if "not built" in str(e):
            return False
        raise


@pytest.mark.parametrize("name", ["imu_classifier_L250_B64.npz", "imu_classifier_L100_B64.npz",
                                  "imu_classifier_L250_B777.npz"])
def test_imu_classifier_fp32_vs_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    L, B = int(g["L"]), int(g["B"])
    clf, _ = make_classifier(int(g["seed_w"]), L)
    x = torch.from_numpy(W.imu_windows(int(g["seed_x"]), B, W.Dims(imu_window=L))).to(DEV)
    with torch.no_grad():
        logits = clf(x)                                   # IMUClassifier.forward surface
        cls, tokens = clf.imu_encoder(x)                  # IMUEncoder.forward surface
        sc = clf.forward_scores(x, want_cls=True)
    assert logits.shape == (B, 32) and tokens.shape == (B, W.Dims(imu_window=L).seq, 128)
    assert rel_err(logits, g["logits"]) < TOL["fp32"]
    assert rel_err(cls, g["cls"]) < TOL["fp32"]
    assert rel_err(tokens[:4], g["tokens_first4"]) < TOL["fp32"]
    assert torch.equal(tokens[:, 0], cls)
    assert torch.equal(sc["logits"], logits) and torch.equal(sc["cls"], cls)
    assert np.array_equal(sc["pred"].cpu().numpy(), g["preds"])                 # labels bit-exact
    assert np.array_equal(logits.max(1)[1].cpu().numpy(), g["preds"])
    # tighter than the contract: fp32 CUDA-core path sits at fp32 round-off from the reference
    assert rel_err(logits, g["logits"]) < 2e-5
    # logit-based OOD scores (spec rows A1/A2) against the float64 spec on the reference logits
    np.testing.assert_allclose(sc["msp"].cpu().numpy(), ood_spec.msp_score(g["logits"]), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sc["energy"].cpu().numpy(), ood_spec.energy_score(g["logits"]), rtol=1e-4, atol=1e-5)


def test_bf16_path_within_contract(golden_dir):
    g = np.load(os.path.join(golden_dir, "imu_classifier_L250_B777.npz"))
    clf, _ = make_classifier(int(g["seed_w"]))
    if not bf16_available(clf):
        pytest.skip("bf16 tcgen05 path not built yet")
    x = torch.from_numpy(W.imu_windows(int(g["seed_x"]), 777)).to(DEV)
    sc = clf.forward_scores(x, precision="bf16", want_cls=True)
    assert rel_err(sc["logits"], g["logits"]) < TOL["bf16"]
    assert rel_err(sc["cls"], g["cls"]) < TOL["bf16"]
    # labels: bit-exact wherever the reference's top-2 margin exceeds twice the logit error bound
    srt = np.sort(g["logits"], 1)
    margin = srt[:, -1] - srt[:, -2]
    err = np.abs(sc["logits"].cpu().numpy() - g["logits"]).max()
    safe = margin > 2 * err
    assert safe.mean() > 0.9
    assert np.array_equal(sc["pred"].cpu().numpy()[safe], g["preds"][safe])


def test_edge_cases_empty_single_ragged_and_dead_inputs():
    clf, sd = make_classifier(11)
    with torch.no_grad():
        empty = clf(torch.zeros(0, 6, 250, device=DEV))
        assert empty.shape == (0, 32)
        x = torch.from_numpy(W.imu_windows(3, 13)).to(DEV)                    # 13 = ragged last tile
        full = clf(x)
        one = clf(x[5:6])
        assert torch.equal(one[0], full[5])                                     # batch-composition invariant
        want, _ = oracle.imu_classifier(x.cpu().numpy(), sd)
        assert rel_err(full, want) < 2e-5
        # channels 1..5 and samples >= 240 are dead -- even NaN/Inf there is harmless
        x2 = x.clone()
        x2[:, 1:] = float("nan")
        x2[:, 0, 240:] = float("inf")
        assert torch.equal(clf(x2), full)
        # non-contiguous view of a larger buffer (stride handled without a copy)
        big = torch.zeros(13, 8, 300, device=DEV)
        big[:, :6, :250] = x
        assert torch.equal(clf(big[:, :6, :250]), full)
        # compact channel-0 buffer with explicit stride == what Evaluator uploads
        compact = x[:, 0, :240].contiguous()
        sc = clf.forward_scores(compact, window_stride=240)
        assert torch.equal(sc["logits"], full)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("n", [1001, 77, 20000])
def test_stored_feature_head_and_maha_scores(precision, n):
    """cmhar_head_forward from stored features: the fp32 CUDA-core kernel and the tcgen05 split-bf16 kernel
    (precision='bf16') both deliver fp32-grade logits, exact labels, and the Mahalanobis score of the spec."""
    clf, sd = make_classifier(12)
    rs = np.random.RandomState(n)
    feats, labels = W.class_features(7, 4000)
    feat = feats[rs.randint(0, 4000, size=n)] + 0.3 * rs.standard_normal((n, 128)).astype(np.float32)
    want = oracle.classifier_head(feat, sd, dtype=torch.float64).numpy()
    maha = cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV))
    f = torch.from_numpy(feat).to(DEV)
    logits = torch.empty(n, 32, device=DEV)
    pred = torch.empty(n, dtype=torch.int64, device=DEV)
    msp, energy, md = torch.empty(n, device=DEV), torch.empty(n, device=DEV), torch.empty(n, device=DEV)
    N = cm._native
    N.check(N.lib().cmhar_head_forward(clf._head_blob(f.device).data_ptr(), maha.blob(f.device).data_ptr(), f.data_ptr(), n,
                                       logits.data_ptr(), pred.data_ptr(), msp.data_ptr(), energy.data_ptr(), md.data_ptr(),
                                       N.BF16 if precision == "bf16" else N.FP32, N.stream_ptr(f.device)))
    torch.cuda.synchronize()
    assert rel_err(logits, want) < 2e-5
    got = logits.cpu().numpy()
    margin = np.sort(want, 1)[:, -1] - np.sort(want, 1)[:, -2]
    safe = margin > 1e-3 * np.abs(want).max()
    assert np.array_equal(pred.cpu().numpy()[safe], want.argmax(1)[safe])
    np.testing.assert_array_equal(pred.cpu().numpy(), got.argmax(1))
    np.testing.assert_allclose(energy.cpu().numpy(), ood_spec.energy_score(got), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(msp.cpu().numpy(), ood_spec.msp_score(got), rtol=1e-4, atol=1e-5)
    st = ood_spec.mahalanobis_finalize(*ood_spec.mahalanobis_sufficient_stats(feats, labels, 32), ridge=1e-3)
    want_m = ood_spec.mahalanobis_score(feat, st)
    assert float(np.abs(md.cpu().numpy() - want_m).max() / np.abs(want_m).max()) < 2e-4


def test_logit_scores_from_stored_logits():
    clf, sd = make_classifier(12)
    rs = np.random.RandomState(0)
    feat = rs.standard_normal((1001, 128)).astype(np.float32)
    want = oracle.classifier_head(feat, sd).numpy()
    s = cm.logit_scores(torch.from_numpy(want).to(DEV), temperature=2.0)
    np.testing.assert_allclose(s["energy"].cpu().numpy(), ood_spec.energy_score(want, T=2.0), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(s["msp"].cpu().numpy(), ood_spec.msp_score(want), rtol=1e-5, atol=1e-7)
    assert np.array_equal(s["pred"].cpu().numpy(), oracle.predict(want))
    # ties: arg-max is the FIRST maximal index (torch logits.max(1))
    z = torch.zeros(4, 40, device=DEV)
    z[1, 7] = z[1, 33] = 2.0
    assert cm.logit_scores(z)["pred"].tolist() == [0, 7, 0, 0]
    # C = 32: the TMA-ring kernel (one thread per row, rotated piece order) -- ties, ragged tail, many tiles per CTA
    z = torch.zeros(300, 32, device=DEV)
    z[1, 7] = z[1, 31] = 2.0
    z[2, 31] = 1.0
    z[129, 4] = z[129, 5] = z[129, 30] = 3.0
    p32 = cm.logit_scores(z)["pred"].tolist()
    assert p32[:3] == [0, 7, 31] and p32[129] == 4 and p32[299] == 0
    for n, T in ((250_001, 1.0), (3 * 148 * 128 * 5 + 17, 0.5)):
        big = (4.0 * np.random.RandomState(n % 1000).standard_normal((n, 32))).astype(np.float32)
        s = cm.logit_scores(torch.from_numpy(big).to(DEV), temperature=T)
        np.testing.assert_allclose(s["energy"].cpu().numpy(), ood_spec.energy_score(big, T=T), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(s["msp"].cpu().numpy(), ood_spec.msp_score(big), rtol=1e-5, atol=1e-7)
        assert np.array_equal(s["pred"].cpu().numpy(), big.argmax(1))


def test_cross_modal_and_losses_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cross_modal_B48.npz"))
    B, T = int(g["B"]), int(g["T"])
    cfg = cm.default_config()
    model = cm.CrossModalModel(cfg)
    model.load_state_dict(tsd(W.cross_modal_state(int(g["seed_w"]))), strict=True)
    model = model.to(DEV).eval()
    imu = torch.from_numpy(W.imu_windows(int(g["seed_x"]), B)).to(DEV)
    fmap = torch.from_numpy(W.video_feature_maps(int(g["seed_v"]), B, T)).to(DEV)
    with torch.no_grad():
        vfeat = model.video_encoder(fmap.view(B, T, 512, 4, 4))             # VideoEncoder.forward surface
        ip, vp = model(imu, fmap.view(B, T, 512, 4, 4))                     # CrossModalModel.forward surface
        ip2, vp2 = model.embed_from_features(imu, fmap, T)
        sig = cm.SigmoidContrastiveLoss().to(DEV)(ip, vp)
        nce = cm.InfoNCELoss()(ip, vp)
        sim = cm.similarity_native(ip, vp, materialize=True)["sim"]
    assert rel_err(vfeat, g["video_feat"]) < TOL["fp32"]
    assert rel_err(ip, g["imu_proj"]) < TOL["fp32"] and rel_err(vp, g["video_proj"]) < TOL["fp32"]
    assert torch.equal(ip, ip2) and torch.equal(vp, vp2)
    assert rel_err(sim, g["similarity"]) < TOL["fp32"]
    assert abs(float(sig) - float(g["sigmoid_loss"])) < 1e-3 * abs(float(g["sigmoid_loss"]))
    assert abs(float(nce) - float(g["info_nce_loss"])) < 1e-3 * abs(float(g["info_nce_loss"]))
    # bf16 feature maps (the layout config 2 names): same tail within the bf16 input rounding
    with torch.no_grad():
        v16 = model.video_encoder.forward_features(fmap.to(torch.bfloat16), T)
    assert rel_err(v16, g["video_feat"]) < TOL["bf16"]


def test_videomae_branch_projection(golden_dir):
    g = np.load(os.path.join(golden_dir, "videomae_projection.npz"))
    cfg = cm.default_config()
    ve = cm.VideoEncoder(cfg)
    sd = W.cross_modal_state(int(g["seed_w"]))
    ve.projection.load_state_dict({"weight": torch.from_numpy(sd["video_encoder.projection.weight"]),
                                   "bias": torch.from_numpy(sd["video_encoder.projection.bias"])})
    ve = ve.to(DEV).eval()
    x = torch.from_numpy(np.random.RandomState(int(g["seed"])).standard_normal((8, 512)).astype(np.float32)).to(DEV)
    out = ve._packed_projection(x.device)(x, relu=False)
    assert rel_err(out, g["out"]) < 2e-5


@pytest.mark.parametrize("hw,dtype", [(16, torch.bfloat16), (16, torch.float32), (49, torch.bfloat16), (1, torch.float32)])
def test_video_pool_shapes(hw, dtype):
    rs = np.random.RandomState(hw)
    n, T, Fd = 5, 3, 96
    fm = torch.from_numpy(np.maximum(rs.standard_normal((n * T, Fd, hw)), 0).astype(np.float32)).to(DEV).to(dtype)
    pooled = torch.empty(n, Fd, device=DEV)
    N = cm._native
    N.check(N.lib().cmhar_video_pool(fm.data_ptr(), int(dtype == torch.bfloat16), n, T, Fd, hw, pooled.data_ptr(), N.stream_ptr(fm.device)))
    want = fm.float().view(n, T, Fd, hw).mean(dim=(1, 3))
    assert rel_err(pooled, want) < 1e-5


@pytest.mark.parametrize("n,T,Fd,hw,dtype", [(1, 16, 512, 16, torch.bfloat16), (37, 16, 512, 16, torch.bfloat16),
                                             (300, 16, 512, 16, torch.bfloat16), (41, 3, 256, 16, torch.float32),
                                             (64, 5, 128, 8, torch.bfloat16), (19, 7, 384, 32, torch.bfloat16)])
def test_video_pool_ring_kernel(n, T, Fd, hw, dtype):
    """The co-resident TMA-ring pooling kernel (channels % 128 == 0, 16..64 bytes per channel, n <= 1024): one clip,
    fewer units than CTAs, many units per persistent CTA (ring stages and parities wrap), 1-, 2- and 4-vector slabs."""
    rs = np.random.RandomState(n + hw)
    fm = torch.from_numpy(np.maximum(rs.standard_normal((n * T, Fd, hw)), 0).astype(np.float32)).to(DEV).to(dtype)
    pooled = torch.full((n, Fd), float("nan"), device=DEV)
    N = cm._native
    N.check(N.lib().cmhar_video_pool_coresident(fm.data_ptr(), int(dtype == torch.bfloat16), n, T, Fd, hw, pooled.data_ptr(), N.stream_ptr(fm.device)))
    want = fm.float().view(n, T, Fd, hw).mean(dim=(1, 3))
    assert rel_err(pooled, want) < 1e-5


def test_similarity_losses_ragged_and_sharded():
    rs = np.random.RandomState(5)
    a = rs.standard_normal((300, 256)).astype(np.float32)
    b = rs.standard_normal((300, 256)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    ta, tb = torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)
    with torch.no_grad():
        sig = float(cm.SigmoidContrastiveLoss().to(DEV)(ta, tb))
        nce = float(cm.InfoNCELoss()(ta, tb))
    assert abs(sig - float(oracle.sigmoid_contrastive_loss(a, b, dtype=torch.float64))) < 1e-5 * sig
    assert abs(nce - float(oracle.info_nce_loss(a, b, dtype=torch.float64))) < 1e-5 * nce
    # row shards (what each rank computes): partial sums add up to the full-matrix sum, and the
    # diagonal offset picks each shard's positives
    full = cm.similarity_native(ta, tb, sigmoid=(10.0, -10.0), lse_scale=1 / 0.07)
    parts = [cm.similarity_native(ta[lo:hi], tb, sigmoid=(10.0, -10.0), lse_scale=1 / 0.07, diag_offset=lo)
             for lo, hi in (cm.shard_bounds(300, r, 4) for r in range(4))]
    assert abs(float(sum(p["sigmoid_sum"] for p in parts)) - float(full["sigmoid_sum"])) < 1e-9 * float(full["sigmoid_sum"])
    assert torch.allclose(torch.cat([p["row_lse"] for p in parts]), full["row_lse"], rtol=1e-6, atol=1e-6)
    assert torch.equal(torch.cat([p["diag"] for p in parts]), full["diag"])


def test_mahalanobis_fit_and_score_vs_spec():
    feats, labels = W.class_features(7, 20011)
    labels[::97] = -1                                            # unlabeled rows are skipped
    labels[5] = 99
    m = cm.MahalanobisOOD(32, DEV)
    f, y = torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV)
    m.accumulate(f[:7000], y[:7000])                             # streamed in ragged shards
    m.accumulate(f[7000:7001], y[7000:7001])
    m.accumulate(f[7001:], y[7001:])
    m.finalize()
    spec = ood_spec.mahalanobis_fit(feats, labels, 32)
    np.testing.assert_allclose(m.fit_["count"], spec["count"])
    np.testing.assert_allclose(m.fit_["mean"], spec["mean"], atol=1e-6)
    np.testing.assert_allclose(m.fit_["cov"], spec["cov"], atol=2e-5)
    q, ql = W.class_features(8, 5003, ood_fraction=0.5)
    got = m.score(torch.from_numpy(q).to(DEV)).cpu().numpy()
    want = ood_spec.mahalanobis_score(q, spec)
    assert rel_err(got, want) < TOL["fp32"]
    # AUROC / FPR95 identical to 3 decimals against the float64 spec
    r = cm.auroc_fpr95(torch.from_numpy(got[ql >= 0]).to(DEV), torch.from_numpy(got[ql < 0]).to(DEV))
    assert round(r["auroc"], 3) == round(ood_spec.auroc(want[ql >= 0], want[ql < 0]), 3)
    assert round(r["fpr"], 3) == round(ood_spec.fpr_at_tpr_fast(want[ql >= 0], want[ql < 0]), 3)
    # the tensor-core FIT kernel (split-bf16 GEMMs over the row dimension): same statistics
    m2 = cm.MahalanobisOOD(32, DEV)
    m2.accumulate(f[:7000], y[:7000], precision="bf16")
    m2.accumulate(f[7000:7001], y[7000:7001], precision="bf16")
    m2.accumulate(f[7001:], y[7001:], precision="bf16")
    m2.finalize()
    np.testing.assert_allclose(m2.fit_["count"], spec["count"])
    np.testing.assert_allclose(m2.fit_["mean"], spec["mean"], atol=2e-6)
    np.testing.assert_allclose(m2.fit_["cov"], spec["cov"], atol=2e-5)
    # the tensor-core score kernel (split-bf16 MMAs): same contract, same AUROC / FPR95
    got_tc = m.score(torch.from_numpy(q).to(DEV), precision="bf16").cpu().numpy()
    assert rel_err(got_tc, want) < 2e-4
    r2 = cm.auroc_fpr95(torch.from_numpy(got_tc[ql >= 0]).to(DEV), torch.from_numpy(got_tc[ql < 0]).to(DEV))
    assert round(r2["auroc"], 3) == round(r["auroc"], 3) and round(r2["fpr"], 3) == round(r["fpr"], 3)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5003, 148 * 128 * 3 + 5])
def test_streaming_mahalanobis_score_kernel(n):
    """maha_score_tc_kernel (one N=160 split-bf16 GEMM against the resident [W | G] image + per-row reduction):
    ragged tails, single-row input, more than two tiles per persistent CTA (both accumulators / staging buffers
    wrap), classes without training rows (their distance must never win)."""
    feats, labels = W.class_features(7, 6000)
    labels = labels.copy()
    labels[labels == 3] = 5                                       # class 3 has no rows
    m = cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV))
    spec = ood_spec.mahalanobis_finalize(*ood_spec.mahalanobis_sufficient_stats(feats, labels, 32), ridge=1e-3)
    q, _ = W.class_features(9, n, ood_fraction=0.5)
    want = ood_spec.mahalanobis_score(q, spec)
    qd = torch.from_numpy(q).to(DEV)
    got = m.score(qd, precision="bf16")
    torch.cuda.synchronize()
    assert got.shape == (n,)
    assert rel_err(got, want) < 2e-4
    ref32 = m.score(qd, precision="fp32")                         # the fp32 CUDA-core kernel: same contract
    assert rel_err(got, ref32) < 2e-4


def test_auroc_fpr95_vs_spec_continuous_tied_and_separated():
    rs = np.random.RandomState(9)
    cases = [
        (rs.standard_normal(50000), rs.standard_normal(30000) + 1.0),                # overlapping
        (np.round(rs.standard_normal(20000) * 4) / 4, np.round(rs.standard_normal(20000) * 4) / 4 + 0.5),  # tied
        (rs.standard_normal(1000) - 50, rs.standard_normal(1000) + 50),              # separated
        (-rs.rand(5000), -rs.rand(5000) * 0.5),                                      # negative scores (MSP)
    ]
    for a, b in cases:
        a, b = a.astype(np.float32), b.astype(np.float32)
        r = cm.auroc_fpr95(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV))
        want_auc, want_fpr = ood_spec.auroc(a, b), ood_spec.fpr_at_tpr_fast(a, b)
        assert abs(r["auroc"] - want_auc) <= r["auroc_bound"] + 1e-12
        assert round(r["auroc"], 3) == round(want_auc, 3)
        assert abs(r["fpr"] - want_fpr) < 1e-12, (r, want_fpr)                      # refined: exact


def test_fused_mahalanobis_and_evaluator_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "evaluator_n200.npz"))
    clf, sd = make_classifier(int(g["seed_w"]))
    n, bs = int(g["n"]), int(g["bs"])
    x = torch.from_numpy(W.imu_windows(int(g["seed_x"]), n))
    labels = torch.from_numpy(np.random.RandomState(int(g["seed_y"])).randint(0, 32, size=n).astype(np.int64))
    loader = [{"imu": x[i:i + bs], "label": labels[i:i + bs]} for i in range(0, n, bs)]
    ev = cm.Evaluator(clf, cm.default_config(), device=DEV)
    res = ev.evaluate(loader)                                     # Evaluator.evaluate surface (a8/a9)
    assert np.array_equal(res["predictions"], g["preds"]) and np.array_equal(res["labels"], g["labels"])
    assert res["predictions"].dtype == np.int64 and res["logits"].dtype == np.float32
    assert rel_err(res["logits"], g["logits"]) < 2e-5
    for k, v in zip(g["metric_names"], g["metrics"]):
        assert abs(res["metrics"][str(k)] - float(v)) < 1e-9
    # Mahalanobis fitted on the model's own CLS features, fused score == standalone score == spec
    maha = ev.fit_mahalanobis(loader, ridge=1e-3)
    out = ev.predict_scores(loader, want_cls=True)
    spec = ood_spec.mahalanobis_finalize(*ood_spec.mahalanobis_sufficient_stats(out["cls"], labels.numpy(), 32), ridge=1e-3)
    want = ood_spec.mahalanobis_score(out["cls"], spec)
    assert rel_err(out["maha"], want) < TOL["fp32"]
    alone = maha.score(torch.from_numpy(out["cls"]).to(DEV)).cpu().numpy()
    assert rel_err(alone, out["maha"]) < 1e-5
    table = ev.evaluate_ood(loader[:2], loader[2:])
    assert set(table) == {"msp", "energy", "maha"} and all(0 <= t["auroc"] <= 1 for t in table.values())


def test_full_size_properties_config3_and_streaming():
    """BASELINE configs 3/5 sizes, checked through size-independent properties."""
    clf, _ = make_classifier(13)
    n = 65536 + 37
    g = torch.Generator(device=DEV).manual_seed(1234)
    x = torch.randn(n, 6, 250, device=DEV, generator=g)
    sc = clf.forward_scores(x)
    assert torch.isfinite(sc["logits"]).all()
    idx = torch.randperm(n, device=DEV, generator=g)[:4096]
    sub = clf.forward_scores(x[idx].contiguous())
    assert torch.equal(sub["logits"], sc["logits"][idx])         # window results independent of batch/tile
    assert torch.equal(sub["pred"], sc["pred"][idx])
    assert torch.equal(sc["pred"], sc["logits"].max(1)[1])
    lse = torch.logsumexp(sc["logits"].double(), 1)
    assert torch.allclose(sc["energy"].double(), -lse, rtol=1e-5, atol=1e-5)
    # config 3: 4096 x 4096 similarity, fused loss == materialised matrix reduced by torch
    e = torch.nn.functional.normalize(torch.randn(4096, 256, device=DEV, generator=g), dim=1)
    v = torch.nn.functional.normalize(torch.randn(4096, 256, device=DEV, generator=g), dim=1)
    r = cm.similarity_native(e, v, materialize=True, sigmoid=(10.0, -10.0), lse_scale=1 / 0.07)
    want = torch.nn.functional.softplus(-(r["sim"].double() * 10 - 10)).sum()
    assert abs(float(r["sigmoid_sum"]) - float(want)) < 1e-6 * float(want)
    assert torch.allclose(r["row_lse"].double(), torch.logsumexp(r["sim"].double() / 0.07, 1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(r["col_lse"].double(), torch.logsumexp(r["sim"].double() / 0.07, 0), rtol=1e-5, atol=1e-5)
    assert rel_err(r["sim"], e.double() @ v.double().T) < 1e-5


def test_cuda_graph_capture_of_fused_forward():
    clf, _ = make_classifier(11)
    x = torch.from_numpy(W.imu_windows(2, 256)).to(DEV)
    eager = clf.forward_scores(x)
    out = {k: torch.empty_like(v) for k, v in eager.items()}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        clf.forward_scores(x, out=out)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        clf.forward_scores(x, out=out)
    for v in out.values():
        v.zero_()
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["logits"], eager["logits"]) and torch.equal(out["pred"], eager["pred"])


@pytest.mark.parametrize("na,nb", [(48, 48), (200, 333), (1024, 640)])
def test_similarity_bf16_tensor_core_path(na, nb):
    """tcgen05 path (precision='bf16'): against the float64 oracle within the bf16 contract, ragged shapes
    (zero-padded tiles are masked out of every reduction), sharded rows (diag_offset)."""
    g = torch.Generator(device=DEV).manual_seed(na * 7 + nb)
    a = torch.nn.functional.normalize(torch.randn(na, 256, device=DEV, generator=g), dim=1)
    b = torch.nn.functional.normalize(torch.randn(nb, 256, device=DEV, generator=g), dim=1)
    r = cm.similarity_native(a, b, materialize=True, sigmoid=(10.0, -10.0), lse_scale=1 / 0.07, precision="bf16")
    want = a.double() @ b.double().T
    assert rel_err(r["sim"], want) < TOL["bf16"]
    # the reductions are functions of the matrix the kernel produced (fp32 accumulators of bf16 products)
    sim = r["sim"].double()
    sp = torch.nn.functional.softplus(-(sim * 10 - 10)).sum()
    assert abs(float(r["sigmoid_sum"]) - float(sp)) < 2e-5 * float(sp)
    assert torch.allclose(r["row_lse"].double(), torch.logsumexp(sim / 0.07, 1), rtol=1e-4, atol=1e-4)
    assert torch.allclose(r["col_lse"].double(), torch.logsumexp(sim / 0.07, 0), rtol=1e-4, atol=1e-4)
    n = min(na, nb)
    assert torch.allclose(r["diag"][:n].double(), torch.diagonal(sim)[:n] / 0.07, rtol=1e-5, atol=1e-5)
    # against the oracle's loss values
    want_sig = torch.nn.functional.softplus(-(want * 10 - 10)).mean()
    assert abs(float(r["sigmoid_sum"]) / (na * nb) - float(want_sig)) < TOL["bf16"] * float(want_sig)
    # row shards with a diagonal offset reproduce the full result
    lo = na // 3
    part = cm.similarity_native(a[lo:], b, sigmoid=(10.0, -10.0), lse_scale=1 / 0.07, diag_offset=lo, precision="bf16")
    assert torch.allclose(part["row_lse"], r["row_lse"][lo:], rtol=1e-6, atol=1e-6)
    assert torch.allclose(part["diag"][:max(0, n - lo)], r["diag"][lo:n], rtol=1e-6, atol=1e-6)


def test_similarity_bf16_full_size_properties():
    """config 3 size (4096 x 4096 x 256) through the tensor-core path: fused reductions == reductions of the
    materialised matrix, and the timing of both paths (information)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    e = torch.nn.functional.normalize(torch.randn(4096, 256, device=DEV, generator=g), dim=1)
    v = torch.nn.functional.normalize(torch.randn(4096, 256, device=DEV, generator=g), dim=1)
    r = cm.similarity_native(e, v, materialize=True, sigmoid=(10.0, -10.0), lse_scale=1 / 0.07, precision="bf16")
    sim = r["sim"].double()
    want = torch.nn.functional.softplus(-(sim * 10 - 10)).sum()
    assert abs(float(r["sigmoid_sum"]) - float(want)) < 2e-5 * float(want)
    assert torch.allclose(r["row_lse"].double(), torch.logsumexp(sim / 0.07, 1), rtol=1e-4, atol=1e-4)
    assert torch.allclose(r["col_lse"].double(), torch.logsumexp(sim / 0.07, 0), rtol=1e-4, atol=1e-4)
    assert rel_err(r["sim"], e.double() @ v.double().T) < TOL["bf16"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for prec in ("fp32", "bf16"):
        cm.similarity_native(e, v, sigmoid=(10.0, -10.0), precision=prec)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(10):
            cm.similarity_native(e, v, sigmoid=(10.0, -10.0), precision=prec)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        print(f"similarity 4096x4096x256 fused sigmoid loss, {prec}: {ms * 1e3:.1f} us  ({2 * 4096 * 4096 * 256 / ms / 1e9:.1f} TFLOP/s)")


@pytest.mark.parametrize("n,k,m", [(256, 512, 768), (77, 128, 512), (300, 768, 512), (1000, 512, 256)])
def test_linear_tensor_core_path(n, k, m):
    """bf16 tcgen05 tiles of the small dense layers against float64, BatchNorm folded, ReLU, ragged rows/cols."""
    rs = np.random.RandomState(n + k)
    lin = torch.nn.Linear(k, m)
    bn = torch.nn.BatchNorm1d(m)
    with torch.no_grad():
        bn.running_mean.copy_(torch.from_numpy(rs.standard_normal(m).astype(np.float32)))
        bn.running_var.copy_(torch.from_numpy(rs.uniform(0.5, 2.0, m).astype(np.float32)))
        bn.weight.copy_(torch.from_numpy(rs.uniform(0.5, 1.5, m).astype(np.float32)))
        bn.bias.copy_(torch.from_numpy(rs.standard_normal(m).astype(np.float32)))
    from crossmodal_imu_video_ood_har_b200.models import _PackedLinear
    pl = _PackedLinear(lin.to(DEV), bn.to(DEV).eval(), DEV)
    x = torch.from_numpy(rs.standard_normal((n, k)).astype(np.float32)).to(DEV)
    want = torch.relu(bn.double()(lin.double()(x.double())))
    got32 = pl(x, relu=True, precision="fp32")
    got16 = pl(x, relu=True, precision="bf16")
    assert rel_err(got32, want) < 1e-5
    assert 1e-6 < rel_err(got16, want) < TOL["bf16"]          # really the bf16 path, and within its contract
    # concat variant: [x1 | x2] never materialised
    if k % 128 == 0:
        k1 = 128
        y = torch.empty(n, m, device=DEV)
        N = cm._native
        N.check(N.lib().cmhar_concat_linear_forward(pl.blob.data_ptr(), x[:, :k1].contiguous().data_ptr(), k1,
                                                    x[:, k1:].contiguous().data_ptr(), k - k1, n, m, 1, y.data_ptr(), None, 0,
                                                    N.BF16, N.stream_ptr(x.device))) if k > k1 else None
        if k > k1:
            assert torch.equal(y, got16)


def test_empty_shards_are_legal_for_the_scoring_entry_points():
    """A rank may hold no row of a population (e.g. all OOD rows of a contiguously sharded set live on the last rank): the
    scoring / histogram / fit entry points accept n = 0 (null data pointers included) and leave their outputs untouched."""
    e = torch.zeros(0, device=DEV)
    r = cm.auroc_fpr95(torch.randn(100, device=DEV), e)
    assert r["auroc"] != r["auroc"]                                   # NaN: no OOD rows anywhere
    assert cm.logit_scores(torch.zeros(0, 32, device=DEV))["msp"].numel() == 0
    feats, labels = W.class_features(3, 500)
    m = cm.MahalanobisOOD(32, DEV, ridge=1e-3)
    m.accumulate(torch.zeros(0, 128, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), precision="bf16")
    m.accumulate(torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV), precision="bf16")
    m.finalize(all_reduce=False)
    assert m.score(torch.zeros(0, 128, device=DEV), precision="bf16").numel() == 0


def test_round2_entry_points_accept_empty_batches_and_refuse_unserved_shapes():
    """n = 0 is a no-op for the new entry points (null pointers included); shapes the fused kernels do not serve come back as
    CMHAR_ERR_UNSUPPORTED (callers then run the chained route), never as a silent wrong answer."""
    lib, N = cm._native.lib(), cm._native
    st = N.stream_ptr(torch.device(DEV))
    assert lib.cmhar_xattn_forward(None, None, None, 0, 16, 16, 512, 1e-5, None, st) == 0
    assert lib.cmhar_conv_encoder_forward_ex(None, None, 0, 250, 1500, None, N.BF16, st) == 0
    assert lib.cmhar_video_pool_frames_img(None, 1, 0, 16, 512, 16, None, None, None, st) == 0
    assert lib.cmhar_near_tie_rows(None, 0, 32, 0.04, None, None, st) == 0
    assert lib.cmhar_xattn_blob_bytes(500) == 0 and lib.cmhar_xattn_blob_bytes(512) > 0
    blob = N.alloc_blob(lib.cmhar_xattn_blob_bytes(512), torch.device(DEV))
    tok = torch.zeros(8, 16, 128, device=DEV)
    img = N.alloc_blob(lib.cmhar_operand_image_bytes(8 * 16, 512), torch.device(DEV))
    out = torch.zeros(8, 128, device=DEV)
    rc = lib.cmhar_xattn_forward(blob.data_ptr(), tok.data_ptr(), img.data_ptr(), 8, 16, 8, 512, 1e-5, out.data_ptr(), st)   # 8 frames
    assert rc == N.UNSUPPORTED
    with pytest.raises(RuntimeError):
        N.check(lib.cmhar_debug_set_option(b"no_such_key", 1))
    assert lib.cmhar_blob_release(None) == 0 and lib.cmhar_blob_release(12345) == 0
