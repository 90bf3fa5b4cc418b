"""GPU tests (-m gpu) of the BENCHMARKED precision: what the bf16 tcgen05 path does and does not meet of the north-star
contract, measured against the reference goldens and asserted as hard bounds.

north_star: logits / embeddings within 2e-2 relative in bf16; predicted labels bit-exact; AUROC / FPR95 identical to 3
decimals.  Measured (tools/parity_probe.py, B200): bf16 logits 4.7e-3 .. 6.0e-3, 2 arg-max flips in 777 windows (both
rows have a reference top-2 margin below the bf16 logit error), AUROC within 2e-3 and FPR95 within 2e-2 of the
reference's at n = 777 -- i.e. plain bf16 meets the tolerance on logits but NOT label exactness / 3-decimal AUROC.
Both hold on the 'fp32' path (CUDA-core encoder, 2e-6 from the reference); 'bf16_refined' (bf16 tensor-core pass + fp32
re-run of the near-tie rows) restores the reference's labels exactly while most rows keep bf16 throughput -- its scores keep
bf16 accuracy, so the 3-decimal AUROC contract stays with 'fp32'.  Every statement is a test below.
"""
import os

import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from oracle import fusion_spec, ood_spec, oracle, weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDENS = ["imu_classifier_L250_B777.npz", "imu_classifier_L250_B64.npz", "imu_classifier_L100_B64.npz"]


def rel_err(got, want):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else np.asarray(want)
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


def tsd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def load_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    L, B = int(g["L"]), int(g["B"])
    cfg = cm.default_config(imu_window_size=L)
    clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
    clf.load_state_dict(tsd(W.classifier_state(int(g["seed_w"]), W.Dims(imu_window=L))), strict=True)
    x = torch.from_numpy(W.imu_windows(int(g["seed_x"]), B, W.Dims(imu_window=L))).to(DEV)
    return g, clf.to(DEV).eval(), x


@pytest.mark.parametrize("name", GOLDENS)
def test_bf16_encoder_vs_every_reference_golden_with_flip_count(golden_dir, name):
    """S = 16 (L = 250) and S = 7 (L = 100, the masked-softmax branch with padding rows) against the reference: tolerance
    on logits and CLS, and the NUMBER of arg-max flips over ALL rows with a hard bound (measured: 2 / 0 / 0)."""
    g, clf, x = load_golden(golden_dir, name)
    sc = clf.forward_scores(x, precision="bf16", want_cls=True)
    torch.cuda.synchronize()
    assert rel_err(sc["logits"], g["logits"]) < 2e-2
    assert rel_err(sc["cls"], g["cls"]) < 2e-2
    logits = sc["logits"].cpu().numpy()
    pred = sc["pred"].cpu().numpy()
    flips = np.flatnonzero(pred != g["preds"])
    srt = np.sort(g["logits"], 1)
    margin = srt[:, -1] - srt[:, -2]
    err = np.abs(logits - g["logits"]).max()
    print(f"{name}: bf16 logits rel err {rel_err(logits, g['logits']):.2e}, label flips {len(flips)}/{len(pred)}, "
          f"reference margins of the flipped rows {np.round(margin[flips], 4).tolist()}")
    assert len(flips) <= max(1, int(0.005 * len(pred)))                     # hard bound: <= 0.5 % of the rows (measured 0.26 %)
    assert np.all(margin[flips] < 2 * err)                                  # only genuine near-ties may flip
    np.testing.assert_array_equal(pred, logits.argmax(1))                   # pred is the arg-max of the logits it returns


def test_bf16_short_sequence_ignores_poisoned_dead_inputs(golden_dir):
    """S = 7: rows 7..15 of every 16-row group are padding -- their keys are masked out of the softmax, and the dead inputs
    (channels 1..5, samples >= 96) are never read, even as NaN / Inf."""
    g, clf, x = load_golden(golden_dir, "imu_classifier_L100_B64.npz")
    clean = clf.forward_scores(x, precision="bf16", want_cls=True)
    x2 = x.clone()
    x2[:, 1:] = float("nan")
    x2[:, 0, 96:] = float("inf")
    bad = clf.forward_scores(x2, precision="bf16", want_cls=True)
    torch.cuda.synchronize()
    assert torch.equal(bad["logits"], clean["logits"]) and torch.equal(bad["cls"], clean["cls"])
    assert torch.isfinite(clean["logits"]).all()
    # batch-composition invariance on the short sequence (a window's result does not depend on its tile neighbours)
    one = clf.forward_scores(x[37:38].contiguous(), precision="bf16")
    assert torch.equal(one["logits"][0], clean["logits"][37])


def _ood_metrics(scores, held):
    return {k: (ood_spec.auroc(v[~held], v[held]), ood_spec.fpr_at_tpr_fast(v[~held], v[held])) for k, v in scores.items()}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_auroc_fpr95_from_encoder_outputs_vs_reference_outputs(golden_dir, precision):
    """AUROC / FPR95 of MSP, energy and Mahalanobis-on-CLS computed from THIS path's encoder outputs against the same
    metrics from the reference's fp32 logits / CLS (held-out activities = the three highest predicted classes).
    fp32: identical to 3 decimals (the contract).  bf16: within 5e-3 (AUROC) / 3e-2 (FPR95 at 68 OOD rows, where one
    row is 1.5 % of TPR); measured 2e-3 / 2e-2 -- plain bf16 does NOT give 3-decimal identity, and the test says so."""
    g, clf, x = load_golden(golden_dir, "imu_classifier_L250_B777.npz")
    sc = clf.forward_scores(x, precision=precision, want_cls=True)
    torch.cuda.synchronize()
    held = g["preds"] >= np.sort(np.unique(g["preds"]))[-3]
    fit = ood_spec.mahalanobis_fit(g["cls"][~held], g["preds"][~held], 32, ridge=1e-3)
    want = _ood_metrics({"msp": ood_spec.msp_score(g["logits"]), "energy": ood_spec.energy_score(g["logits"]),
                         "maha": ood_spec.mahalanobis_score(g["cls"], fit)}, held)
    # device route: Mahalanobis scored by the tensor-core kernel on this path's CLS, metrics by the histogram kernels
    m = cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(torch.from_numpy(g["cls"][~held]).to(DEV), torch.from_numpy(g["preds"][~held]).to(DEV))
    got_scores = {"msp": sc["msp"], "energy": sc["energy"], "maha": m.score(sc["cls"], precision=precision)}
    hd = torch.from_numpy(held).to(DEV)
    for k, (a0, f0) in want.items():
        r = cm.auroc_fpr95(got_scores[k][~hd].contiguous(), got_scores[k][hd].contiguous())
        print(f"{precision} {k}: AUROC ref {a0:.5f} got {r['auroc']:.5f}  FPR95 ref {f0:.5f} got {r['fpr']:.5f}")
        if precision == "fp32":
            assert round(r["auroc"], 3) == round(a0, 3) and round(r["fpr"], 3) == round(f0, 3)
        else:
            assert abs(r["auroc"] - a0) < 5e-3 and abs(r["fpr"] - f0) < 3e-2


def _bench_step_modules(seed=41):
    """The modules of bench.py's step (configs[1]) with the deterministic test weights."""
    cfg = cm.default_config()
    sd, sd_x = fusion_spec.fusion_state(seed), W.cross_modal_state(seed)
    xm = cm.CrossModalModel(cfg)
    xm.load_state_dict(tsd(sd_x), strict=True)
    clf = cm.IMUClassifier(xm.imu_encoder, cfg)
    fus = cm.LateFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg)
    own = set(fus.state_dict().keys())
    fus.load_state_dict({k: v for k, v in tsd(sd).items() if k in own}, strict=True)
    return xm.to(DEV).eval(), clf.to(DEV).eval(), fus.to(DEV).eval(), sd, sd_x


def test_exact_bench_step_bf16_late_fusion_mahalanobis_graph_vs_spec():
    """The configuration bench.py times -- bf16, LateFusionClassifier, Mahalanobis attached, batch 256, one CUDA graph --
    against oracle/fusion_spec.py + oracle/ood_spec.py in float64, INCLUDING `maha` and `pred`."""
    xm, clf, fus, sd, sd_x = _bench_step_modules()
    B, T, NFIT = 256, 16, 512
    imu, fmap = W.imu_windows(5, B), W.video_feature_maps(6, B, T)
    f_dev = torch.from_numpy(fmap).to(DEV).to(torch.bfloat16)
    f_r = f_dev.float().cpu().numpy()                          # the oracle sees the same bf16-rounded feature maps
    # Mahalanobis fitted on FUSED features of a separate ID set: device fit (bf16 path) vs float64 spec fit
    imu_fit = W.imu_windows(15, NFIT)
    fit_dev = torch.from_numpy(W.video_feature_maps(16, NFIT, T)).to(DEV).to(torch.bfloat16)
    y_fit = np.random.RandomState(17).randint(0, 32, size=NFIT).astype(np.int64)
    maha = cm.MahalanobisOOD(32, DEV, ridge=1e-3)
    maha.accumulate(fus.forward_scores(torch.from_numpy(imu_fit).to(DEV), fit_dev, T, precision="bf16")["fused"],
                    torch.from_numpy(y_fit).to(DEV), precision="bf16")
    maha.finalize(all_reduce=False)
    _, fused_fit = fusion_spec.late_fusion(imu_fit, fit_dev.float().cpu().numpy(), sd, T, dtype=torch.float64)
    spec_fit = ood_spec.mahalanobis_fit(fused_fit.numpy(), y_fit, 32, ridge=1e-3)
    pipe = cm.CrossModalOODPipeline(clf, xm, maha, frames=T, precision="bf16", fusion=fus)
    x = torch.from_numpy(imu).to(DEV)
    graph, out = pipe.capture(x, f_dev)
    for v in out.values():
        if v.dtype.is_floating_point:
            v.fill_(float("nan"))
    graph.replay()
    torch.cuda.synchronize()
    want, want_f = fusion_spec.late_fusion(imu, f_r, sd, T, dtype=torch.float64)
    assert rel_err(out["fused"], want_f) < 2e-2
    assert rel_err(out["logits"], want) < 2e-2
    got = out["logits"].cpu().numpy()
    pred = out["pred"].cpu().numpy()
    np.testing.assert_array_equal(pred, got.argmax(1))
    wl = want.numpy()
    srt = np.sort(wl, 1)
    flips = np.flatnonzero(pred != wl.argmax(1))
    print(f"bench step bf16: logits rel err {rel_err(got, wl):.2e}, label flips {len(flips)}/{B}")
    assert len(flips) <= 2 and np.all((srt[:, -1] - srt[:, -2])[flips] < 2 * np.abs(got - wl).max())
    np.testing.assert_allclose(out["energy"].cpu().numpy(), ood_spec.energy_score(got), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["msp"].cpu().numpy(), ood_spec.msp_score(got), rtol=1e-4, atol=1e-5)
    want_m = ood_spec.mahalanobis_score(want_f.numpy(), spec_fit)
    m_err = float(np.abs(out["maha"].cpu().numpy() - want_m).max() / np.abs(want_m).max())
    print(f"bench step bf16: Mahalanobis rel err {m_err:.2e}")
    assert m_err < 5e-2                       # a squared distance of bf16 features against a bf16-fitted state
    ip, vp = oracle.cross_modal(imu, f_r, sd_x, T, dtype=torch.float64)
    assert rel_err(out["imu_proj"], ip) < 2e-2 and rel_err(out["video_proj"], vp) < 2e-2
    want_loss = float(oracle.sigmoid_contrastive_loss(ip, vp, dtype=torch.float64))
    assert abs(float(out["loss"]) - want_loss) < 2e-2 * want_loss
    # eager == graph
    eager = pipe.run(x, f_dev)
    torch.cuda.synchronize()
    assert torch.equal(eager["logits"], out["logits"]) and torch.equal(eager["maha"], out["maha"])


@pytest.mark.parametrize("name", GOLDENS)
def test_bf16_refined_restores_the_reference_labels_exactly(golden_dir, name):
    """precision='bf16_refined' = bf16 tensor-core pass + fp32 re-run of the near-tie rows (``cmhar_near_tie_rows``): ZERO label
    flips against the reference goldens (S = 16 and S = 7), the re-run rows carry the fp32 path's bits, every other row the
    bf16 path's; only a small fraction of the rows is re-run."""
    g, clf, x = load_golden(golden_dir, name)
    fast = {k: v.clone() for k, v in clf.forward_scores(x, precision="bf16", want_cls=True).items()}
    exact = clf.forward_scores(x, precision="fp32", want_cls=True)
    got = clf.forward_scores(x, precision="bf16_refined", want_cls=True)
    torch.cuda.synchronize()
    k = got["refined_rows"]
    pred = got["pred"].cpu().numpy()
    print(f"{name}: bf16_refined re-ran {k}/{len(pred)} rows; flips vs reference {int((pred != g['preds']).sum())} "
          f"(plain bf16: {int((fast['pred'].cpu().numpy() != g['preds']).sum())})")
    np.testing.assert_array_equal(pred, g["preds"])                         # bit-exact labels
    assert 0 < k < 0.6 * len(pred)        # measured: 9 % of the rows at S = 16 (B = 777), 44 % on the low-margin S = 7 golden
    same32 = (got["logits"] == exact["logits"]).all(1)
    same16 = (got["logits"] == fast["logits"]).all(1)
    assert bool((same32 | same16).all()) and int(same32.sum()) >= k
    assert rel_err(got["logits"], g["logits"]) < 2e-2
