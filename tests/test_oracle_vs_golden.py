"""Pins the CPU oracle (oracle/oracle.py) against golden vectors recorded from the UNMODIFIED
reference modules (oracle/make_golden.py), and the spec-derived OOD oracle against scipy/sklearn.
Runs on CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import ood_spec, oracle, weights as W

FP32_ATOL = 2e-5      # two fp32 evaluation orders of the same algebra (reference vs restatement)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", ["imu_classifier_L250_B64.npz", "imu_classifier_L100_B64.npz",
                                  "imu_classifier_L250_B777.npz"])
def test_imu_classifier_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    dims = W.Dims(imu_window=int(g["L"]))
    sd = W.classifier_state(int(g["seed_w"]), dims)
    x = W.imu_windows(int(g["seed_x"]), int(g["B"]), dims)
    logits, cls = oracle.imu_classifier(x, sd, dims)
    _, tokens = oracle.imu_encoder(x[:4], sd, dims, "imu_encoder.")
    pe = oracle.patch_embed(torch.from_numpy(x[:4]), sd, dims, "imu_encoder.", torch.float32)
    assert tokens.shape[1] == dims.seq
    np.testing.assert_allclose(pe.numpy(), g["patch_embed_first4"], atol=FP32_ATOL, rtol=0)
    np.testing.assert_allclose(cls.numpy(), g["cls"], atol=FP32_ATOL, rtol=0)
    np.testing.assert_allclose(tokens.numpy(), g["tokens_first4"], atol=FP32_ATOL, rtol=0)
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=2e-4, rtol=0)
    assert np.array_equal(oracle.predict(logits), g["preds"])
    assert float(g["dead_input_delta"]) == 0.0


def test_dead_inputs_do_not_matter(golden_dir):
    """only channel 0, samples [0, 16*N) are live."""
    dims = W.Dims()
    sd = W.classifier_state(11, dims)
    x = W.imu_windows(21, 8, dims)
    a, _ = oracle.imu_classifier(x, sd, dims)
    x2 = x.copy()
    x2[:, 1:] = 1e3
    x2[:, 0, 16 * dims.num_patches:] = -1e3
    b, _ = oracle.imu_classifier(x2, sd, dims)
    assert torch.equal(a, b)


def test_cross_modal_and_losses_match_reference(golden_dir):
    g = _load(golden_dir, "cross_modal_B48.npz")
    dims = W.Dims()
    B, T = int(g["B"]), int(g["T"])
    sd = W.cross_modal_state(int(g["seed_w"]), dims)
    imu = W.imu_windows(int(g["seed_x"]), B, dims)
    fmap = W.video_feature_maps(int(g["seed_v"]), B, T, dims)
    vfeat = oracle.video_tail(fmap, sd, T)
    np.testing.assert_allclose(vfeat.numpy(), g["video_feat"], atol=FP32_ATOL, rtol=0)
    ip, vp = oracle.cross_modal(imu, fmap, sd, T, dims)
    np.testing.assert_allclose(ip.numpy(), g["imu_proj"], atol=FP32_ATOL, rtol=0)
    np.testing.assert_allclose(vp.numpy(), g["video_proj"], atol=FP32_ATOL, rtol=0)
    np.testing.assert_allclose(oracle.similarity_matrix(ip, vp).numpy(), g["similarity"], atol=FP32_ATOL, rtol=0)
    assert abs(float(oracle.sigmoid_contrastive_loss(g["imu_proj"], g["video_proj"])) - float(g["sigmoid_loss"])) < 1e-5
    assert abs(float(oracle.info_nce_loss(g["imu_proj"], g["video_proj"])) - float(g["info_nce_loss"])) < 1e-5
    # the reference's "SigLIP" loss is mean softplus(-z) over ALL pairs
    z = torch.from_numpy(g["similarity"]).double() * 10.0 - 10.0
    assert abs(float(torch.nn.functional.softplus(-z).mean()) - float(g["sigmoid_loss"])) < 1e-5


def test_videomae_projection_matches_reference(golden_dir):
    g = _load(golden_dir, "videomae_projection.npz")
    sd = W.cross_modal_state(int(g["seed_w"]))
    cls_tok = np.random.RandomState(int(g["seed"])).standard_normal((8, 512)).astype(np.float32)
    np.testing.assert_allclose(oracle.video_cls_projection(cls_tok, sd).numpy(), g["out"], atol=FP32_ATOL, rtol=0)


def test_evaluator_matches_reference(golden_dir):
    g = _load(golden_dir, "evaluator_n200.npz")
    dims = W.Dims()
    sd = W.classifier_state(int(g["seed_w"]), dims)
    x = W.imu_windows(int(g["seed_x"]), int(g["n"]), dims)
    labels = np.random.RandomState(int(g["seed_y"])).randint(0, 32, size=int(g["n"]))
    logits, _ = oracle.imu_classifier(x, sd, dims)
    preds = oracle.predict(logits)
    assert np.array_equal(preds, g["preds"]) and np.array_equal(labels, g["labels"])
    m = oracle.compute_metrics(labels, preds)
    for k, v in zip(g["metric_names"], g["metrics"]):
        assert abs(m[str(k)] - float(v)) < 1e-9, k
    rs = np.random.RandomState(int(g["m2_seed"]))
    yt = rs.randint(0, 32, size=5000)
    yp = np.where(rs.rand(5000) < 0.7, yt, rs.randint(0, 30, size=5000))
    m2 = oracle.compute_metrics(yt, yp)
    for k, v in zip(sorted(m2), g["m2"]):
        assert abs(m2[k] - float(v)) < 1e-9, k


def test_fp64_truth_is_close_to_fp32_reference(golden_dir):
    """Error bar: the reference's own fp32 result sits ~1e-6 from the float64 evaluation."""
    g = _load(golden_dir, "imu_classifier_L250_B64.npz")
    dims = W.Dims()
    sd = W.classifier_state(int(g["seed_w"]), dims)
    x = W.imu_windows(int(g["seed_x"]), 64, dims)
    l64, _ = oracle.imu_classifier(x, sd, dims, dtype=torch.float64)
    err = np.abs(l64.numpy() - g["logits"]).max() / np.abs(g["logits"]).max()
    assert err < 1e-5


# ----------------------------------------------------------------- spec-derived rows A1-A5
def test_msp_energy_against_scipy():
    from scipy.special import logsumexp, softmax
    z = np.random.RandomState(0).standard_normal((500, 32)) * 5
    np.testing.assert_allclose(ood_spec.msp_score(z), -softmax(z, axis=1).max(1), rtol=1e-12)
    np.testing.assert_allclose(ood_spec.energy_score(z), -logsumexp(z, axis=1), rtol=1e-12)
    np.testing.assert_allclose(ood_spec.energy_score(z, T=2.0), -2 * logsumexp(z / 2, axis=1), rtol=1e-12)


def test_auroc_fpr95_against_sklearn():
    from sklearn.metrics import roc_auc_score, roc_curve
    rs = np.random.RandomState(1)
    for quant in (None, 0.25):                      # continuous and heavily tied scores
        a = rs.standard_normal(3000)
        b = rs.standard_normal(2000) + 1.0
        if quant:
            a, b = np.round(a / quant) * quant, np.round(b / quant) * quant
        y = np.r_[np.zeros(len(a)), np.ones(len(b))]
        s = np.r_[a, b]
        assert abs(ood_spec.auroc(a, b) - roc_auc_score(y, s)) < 1e-12
        fpr, tpr, _ = roc_curve(y, s, drop_intermediate=False)
        want = fpr[np.argmax(tpr >= 0.95)]
        assert abs(ood_spec.fpr_at_tpr(a, b) - want) < 1e-12
        assert abs(ood_spec.fpr_at_tpr_fast(a, b) - want) < 1e-12


def test_mahalanobis_against_direct_formula():
    f, y = W.class_features(3, 4000, num_classes=32, dim=128)
    fit = ood_spec.mahalanobis_fit(f, y, 32)
    # direct tied covariance
    f64 = f.astype(np.float64)
    mu = np.stack([f64[y == c].mean(0) for c in range(32)])
    cov = sum(((f64[y == c] - mu[c]).T @ (f64[y == c] - mu[c])) for c in range(32)) / len(f64)
    np.testing.assert_allclose(fit["mean"], mu, atol=1e-10)
    np.testing.assert_allclose(fit["cov"], cov, atol=1e-9)
    np.testing.assert_allclose(fit["precision"], np.linalg.inv(cov), rtol=1e-6, atol=1e-8)
    q, _ = W.class_features(4, 300, ood_fraction=0.5)
    s = ood_spec.mahalanobis_score(q, fit)
    w = (q.astype(np.float64) @ fit["whiten"])[:, None, :] - fit["mean_whitened"][None]
    np.testing.assert_allclose(s, (w ** 2).sum(-1).min(1), rtol=1e-9)
    assert s[150:].mean() > s[:150].mean()          # OOD rows score higher
