"""CPU tests (-m "not gpu"): state_dict/key parity with the reference, the autograd route of the
drop-in modules against the oracle and the reference goldens, the C-ABI export list, host-side
metric / ROC / Mahalanobis algebra and the world_size-2 gloo reductions.  No CUDA compute."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from oracle import ood_spec, oracle, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torch_sd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def _classifier(seed=11, L=250):
    cfg = cm.default_config(imu_window_size=L)
    clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
    sd = W.classifier_state(seed, W.Dims(imu_window=L))
    clf.load_state_dict(_torch_sd(sd), strict=True)          # key + shape parity (Appendix B)
    return clf, sd


def test_state_dict_keys_match_reference_layout():
    clf, sd = _classifier()
    assert set(clf.state_dict().keys()) == set(sd.keys())
    cfg = cm.default_config()
    model = cm.CrossModalModel(cfg)
    ref = W.cross_modal_state(31)
    got = {k for k in model.state_dict() if not k.startswith("video_encoder.backbone")}
    assert got == set(ref.keys())
    model.load_state_dict(_torch_sd(ref), strict=True)
    n_enc = sum(p.numel() for p in clf.imu_encoder.parameters())
    n_head = sum(p.numel() for p in clf.classifier.parameters())
    assert (n_enc, n_head) == (808576, 70816)


@pytest.mark.parametrize("name", ["imu_classifier_L250_B64.npz", "imu_classifier_L100_B64.npz"])
def test_autograd_route_matches_reference_golden(golden_dir, name):
    """eval() with gradients enabled takes the differentiable torch route; it must equal the
    reference's CPU outputs (this is the route the reference's trainers use)."""
    g = np.load(os.path.join(golden_dir, name))
    L = int(g["L"])
    clf, _ = _classifier(int(g["seed_w"]), L)
    clf.eval()
    x = torch.from_numpy(W.imu_windows(int(g["seed_x"]), int(g["B"]), W.Dims(imu_window=L)))
    logits = clf(x)
    assert logits.requires_grad
    np.testing.assert_allclose(logits.detach().numpy(), g["logits"], atol=3e-4, rtol=0)
    cls, tok = clf.imu_encoder(x[:4])
    np.testing.assert_allclose(tok.detach().numpy(), g["tokens_first4"], atol=3e-5, rtol=0)
    logits.sum().backward()
    assert clf.imu_encoder.patch_embed.projections[0].weight.grad.abs().sum() > 0
    # dead channels get exactly zero gradient
    assert clf.imu_encoder.patch_embed.projections[3].weight.grad.abs().sum() == 0


def test_train_mode_runs_and_dropout_is_active():
    clf, _ = _classifier()
    clf.train()
    x = torch.from_numpy(W.imu_windows(5, 8))
    a, b = clf(x), clf(x)
    assert not torch.allclose(a, b)


def test_inference_route_refuses_cpu_tensors():
    clf, _ = _classifier()
    clf.eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        clf(torch.zeros(2, 6, 250))
    with pytest.raises(RuntimeError, match="CUDA"):
        cm.Evaluator(clf, cm.default_config(), device="cpu")
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        cm.InfoNCELoss()(torch.zeros(4, 8), torch.zeros(4, 8))


def test_cross_modal_autograd_route_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cross_modal_B48.npz"))
    cfg = cm.default_config()
    model = cm.CrossModalModel(cfg)
    model.load_state_dict(_torch_sd(W.cross_modal_state(int(g["seed_w"]))), strict=True)
    model.eval()
    B, T = int(g["B"]), int(g["T"])
    imu = torch.from_numpy(W.imu_windows(int(g["seed_x"]), B))
    video = torch.from_numpy(W.video_feature_maps(int(g["seed_v"]), B, T)).view(B, T, 512, 4, 4)
    ip, vp = model(imu, video)
    np.testing.assert_allclose(ip.detach().numpy(), g["imu_proj"], atol=3e-5, rtol=0)
    np.testing.assert_allclose(vp.detach().numpy(), g["video_proj"], atol=3e-5, rtol=0)
    sig = cm.SigmoidContrastiveLoss()(ip, vp)
    nce = cm.InfoNCELoss()(ip, vp)
    assert abs(float(sig) - float(g["sigmoid_loss"])) < 1e-4
    assert abs(float(nce) - float(g["info_nce_loss"])) < 1e-4
    assert {"temperature", "bias"} <= set(dict(cm.SigmoidContrastiveLoss().named_parameters()))


def test_module_lifecycle_deepcopy_and_unknown_backbone():
    clf, _ = _classifier()
    twin = copy.deepcopy(clf)                                   # main.py:166-167
    assert twin._packed == {} and twin.imu_encoder._packed is not clf.imu_encoder._packed
    assert not clf.freeze_encoder
    frozen = cm.IMUClassifier(cm.IMUEncoder(cm.default_config()), cm.default_config(), freeze_encoder=True)
    assert frozen.freeze_encoder
    frozen.unfreeze_encoder()
    assert not frozen.freeze_encoder
    with pytest.raises(ValueError, match="Backbone inconnu"):   # src/models/models.py:176
        cm.VideoEncoder(cm.default_config(video_backbone="vgg"))


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cmhar_b200.h")).read()
    declared = set(re.findall(r"\b(cmhar_[a-z0-9_]+)\s*\(", header))
    assert declared == set(cm._native.EXPORTED), declared ^ set(cm._native.EXPORTED)
    lib = ctypes.CDLL(cm._native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert cm._native.lib().cmhar_abi_version() == 3
    # argument validation happens before any CUDA call
    assert cm._native.lib().cmhar_imu_encoder_blob_bytes(16, 4) > 3_000_000
    assert cm._native.lib().cmhar_imu_encoder_blob_bytes(17, 4) == 0
    assert cm._native.lib().cmhar_head_blob_bytes(256, 128, 32) > 0
    rc = cm._native.lib().cmhar_imu_forward(None, None, None, None, 4, 1500, *([None] * 7), 0, None)
    assert rc == -1 and b"null" in cm._native.lib().cmhar_last_error()


def test_classification_metrics_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "evaluator_n200.npz"))
    m = cm.classification_metrics(g["labels"], g["preds"])
    for k, v in zip(g["metric_names"], g["metrics"]):
        assert abs(m[str(k)] - float(v)) < 1e-9
    rs = np.random.RandomState(int(g["m2_seed"]))
    yt = rs.randint(0, 32, size=5000)
    yp = np.where(rs.rand(5000) < 0.7, yt, rs.randint(0, 30, size=5000))
    m2 = cm.classification_metrics(yt, yp)
    for k, v in zip(sorted(m2), g["m2"]):
        assert abs(m2[k] - float(v)) < 1e-9


def test_roc_from_histograms_against_spec():
    rs = np.random.RandomState(2)
    a = np.round(rs.standard_normal(4000) * 8) / 8              # heavily tied scores
    b = np.round((rs.standard_normal(3000) + 1.2) * 8) / 8
    edges = np.unique(np.r_[a, b])
    ha = np.array([(a == e).sum() for e in edges])
    hb = np.array([(b == e).sum() for e in edges])
    r = cm.roc_from_histograms(ha, hb)
    assert abs(r["auroc"] - ood_spec.auroc(a, b)) < 1e-12       # one distinct value per bin: exact
    assert abs(r["fpr"] - ood_spec.fpr_at_tpr(a, b)) < 1e-12
    # coarse bins: the reported bound really bounds the error
    ha2, hb2 = ha.reshape(-1, 1)[: len(ha) // 4 * 4].reshape(-1, 4).sum(1), hb[: len(hb) // 4 * 4].reshape(-1, 4).sum(1)
    keep = len(ha) // 4 * 4
    a2, b2 = a[np.isin(a, edges[:keep])], b[np.isin(b, edges[:keep])]
    r2 = cm.roc_from_histograms(ha2, hb2)
    assert abs(r2["auroc"] - ood_spec.auroc(a2, b2)) <= r2["auroc_bound"] + 1e-12
    assert abs(r2["fpr"] - ood_spec.fpr_at_tpr(a2, b2)) <= r2["fpr_bound"] + 1e-12


def test_finalize_mahalanobis_against_spec():
    f, y = W.class_features(3, 3000)
    n, s, ff = ood_spec.mahalanobis_sufficient_stats(f, y, 32)
    got = cm.finalize_mahalanobis(n, s, ff)
    want = ood_spec.mahalanobis_finalize(n, s, ff)
    np.testing.assert_allclose(got["mean"], want["mean"], atol=1e-12)
    np.testing.assert_allclose(got["whiten"] @ got["whiten"].T, want["precision"], rtol=1e-8, atol=1e-10)


def test_shard_bounds_cover_everything_once():
    for n, w in ((10, 3), (1_000_003, 8), (5, 8), (0, 2)):
        spans = [cm.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_window_shard_round_trip(tmp_path):
    """Packed shard: live bytes are a contiguous slice, full windows rebuild bit for bit,
    per-window .npy files in the reference layout (T, C) pack to the same shard."""
    x = W.imu_windows(3, 37)
    y = np.arange(37) % 5
    p = str(tmp_path / "a.cmw")
    cm.write_shard(p, x, y)
    sh = cm.WindowShard(p)
    assert len(sh) == 37 and sh.live_len == 240 == cm.live_samples(250) and cm.live_samples(100) == 96
    np.testing.assert_array_equal(np.asarray(sh.live), x[:, 0, :240])
    for i in (0, 17, 36):
        np.testing.assert_array_equal(sh.window_full(i), x[i])
    got = list(sh.batches(16, lo=5, hi=30))
    assert [b["imu"].shape[0] for b in got] == [16, 9] and got[0]["imu"].shape[1] == 240
    np.testing.assert_array_equal(torch.cat([b["label"] for b in got]).numpy(), y[5:30])
    paths = []
    for i in range(4):
        f = tmp_path / f"w{i}.npy"
        np.save(f, x[i].T)                                   # reference layout on disk: (T, C)
        paths.append(str(f))
    assert cm.pack_npy_windows(paths, y[:4], str(tmp_path / "b.cmw")) == 4
    np.testing.assert_array_equal(np.asarray(cm.WindowShard(str(tmp_path / "b.cmw")).live), x[:4, 0, :240])


def test_ood_table_generator(tmp_path):
    """scorer x split table in the reference's "mean ± std" style, three output formats."""
    import pandas as pd
    rows = []
    for run in range(3):
        res = {"msp": {"auroc": 0.80 + 0.01 * run, "fpr95": 0.60}, "maha": {"auroc": 0.95, "fpr": 0.20 + 0.02 * run}}
        rows += cm.ood_rows("pretrained", "holdout_8_of_32", res, run)
    tables = cm.generate_ood_table(pd.DataFrame(rows))
    assert tables["ood_auroc"].loc["pretrained", ("holdout_8_of_32", "msp")] == "81.00 ± 1.00"
    assert tables["ood_fpr95"].loc["pretrained", ("holdout_8_of_32", "maha")] == "22.00 ± 2.00"
    files = cm.save_tables(tables, tmp_path, prefix="table_ood")
    assert len(files) == 9 and all(os.path.getsize(f) > 0 for f in files)


def test_pack_generation_moves_whenever_packed_blobs_are_dropped():
    """Recorded CUDA graphs (stream_host ring slots) key on this counter: it must move on every event that can change
    the packed blobs -- mode switches, load_state_dict, .to(), attaching a Mahalanobis scorer."""
    import crossmodal_imu_video_ood_har_b200 as cm
    from crossmodal_imu_video_ood_har_b200.models import pack_generation
    cfg = cm.default_config()
    clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
    g0 = pack_generation()
    clf.eval()
    g1 = pack_generation()
    clf.load_state_dict(clf.state_dict(), strict=True)
    g2 = pack_generation()
    clf.float()
    g3 = pack_generation()
    clf.set_mahalanobis(None)
    g4 = pack_generation()
    assert g0 < g1 < g2 < g3 < g4


def test_held_out_activity_split_numpy_and_torch():
    """ID = labelled rows outside the held-out activities, OOD = rows of the held-out activities; unlabelled rows (-1) are in
    neither population (the reference's datasets.py has no such split)."""
    import numpy as np
    import torch
    from crossmodal_imu_video_ood_har_b200.sweep import held_out_activity_split
    y = np.array([0, 5, 31, 5, -1, 7, 30, 30])
    idm, oodm = held_out_activity_split(y, [30, 31, 31])
    assert idm.tolist() == [True, True, False, True, False, True, False, False]
    assert oodm.tolist() == [False, False, True, False, False, False, True, True]
    idt, oodt = held_out_activity_split(torch.from_numpy(y), [30, 31])
    assert idt.tolist() == idm.tolist() and oodt.tolist() == oodm.tolist()
    idm2, oodm2 = held_out_activity_split(y, [])
    assert oodm2.sum() == 0 and idm2.sum() == 7


def test_mahalanobis_finalize_route_selection_on_cpu():
    """Statistics on the CPU take the host fp64 route by default (what the gloo test relies on); asking for the device route on CPU
    statistics fails loudly; `fit_` stays a plain dict of fp64 arrays; a failed fit leaves no blob behind."""
    feats, labels = W.class_features(5, 3000)
    m = cm.MahalanobisOOD(32, device="cpu", ridge=1e-3)
    n, s, ff = ood_spec.mahalanobis_sufficient_stats(feats, labels, 32)
    cnt, ssum, second = m._views()
    cnt.copy_(torch.from_numpy(n)); ssum.copy_(torch.from_numpy(s)); second.copy_(torch.from_numpy(ff))
    m.finalize(all_reduce=False)
    assert m._fit_dev is None and m.finalize_status() == 0
    spec = ood_spec.mahalanobis_fit(feats, labels, 32, ridge=1e-3)
    np.testing.assert_allclose(m.fit_["mean"], spec["mean"], atol=1e-12)
    np.testing.assert_allclose(m.fit_["cov"], spec["cov"], atol=1e-10)
    w = m.fit_["whiten"]
    np.testing.assert_allclose(w @ w.T @ m.fit_["cov"], np.eye(128), atol=1e-8)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.finalize(all_reduce=False, on_device=True)
    m.reset()
    with pytest.raises(ValueError, match="no labelled rows"):
        m.finalize(all_reduce=False)
    with pytest.raises(RuntimeError, match="fit"):
        m.blob("cpu")
