"""GPU probe (development): how far is the bf16 tensor-core path from the label / AUROC contract?

For every reference golden of the IMU classifier: number of arg-max flips of the bf16 path against the golden
(fp32 reference) labels, the top-2 margins of the flipped rows, and AUROC / FPR95 of MSP, energy and the
Mahalanobis score computed from bf16-encoder outputs against the same metrics from the golden fp32 outputs
(ID = rows the reference predicts into the lower classes, OOD = the held-out upper classes).
    python tools/parity_probe.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_imu_video_ood_har_b200 as cm  # noqa: E402
from oracle import ood_spec, weights as W  # noqa: E402

DEV = "cuda:0"


def main():
    for name in ("imu_classifier_L250_B777.npz", "imu_classifier_L250_B64.npz", "imu_classifier_L100_B64.npz"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name))
        L, B = int(g["L"]), int(g["B"])
        cfg = cm.default_config(imu_window_size=L)
        clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
        sd = W.classifier_state(int(g["seed_w"]), W.Dims(imu_window=L))
        clf.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        clf = clf.to(DEV).eval()
        x = torch.from_numpy(W.imu_windows(int(g["seed_x"]), B, W.Dims(imu_window=L))).to(DEV)
        for prec in ("fp32", "bf16"):
            sc = clf.forward_scores(x, precision=prec, want_cls=True)
            torch.cuda.synchronize()
            logits = sc["logits"].cpu().numpy()
            pred = sc["pred"].cpu().numpy()
            err = np.abs(logits - g["logits"]).max()
            rel = err / np.abs(g["logits"]).max()
            srt = np.sort(g["logits"], 1)
            margin = srt[:, -1] - srt[:, -2]
            flips = np.flatnonzero(pred != g["preds"])
            print(f"{name} {prec}: logits rel err {rel:.2e} (abs {err:.3e}), cls rel err "
                  f"{np.abs(sc['cls'].cpu().numpy() - g['cls']).max() / np.abs(g['cls']).max():.2e}, label flips {len(flips)}/{B}"
                  f" margins of flipped rows {np.round(margin[flips], 4).tolist()}  rows with margin < 2*err: {(margin < 2 * err).sum()}")
            # bf16 margin of the flipped rows as the kernel sees them
            s2 = np.sort(logits, 1)
            m2 = s2[:, -1] - s2[:, -2]
            if len(flips):
                print(f"    kernel-side margins of flipped rows {np.round(m2[flips], 4).tolist()}; row max |logit| {np.round(np.abs(logits[flips]).max(1), 2).tolist()}")
            for frac in (0.005, 0.01, 0.02, 0.04):
                flagged = m2 < frac * np.abs(logits).max(1)
                print(f"    refine rule margin < {frac} * rowmax: flags {flagged.sum()} rows, catches {np.isin(flips, np.flatnonzero(flagged)).sum()}/{len(flips)} flips")
            if B < 200:
                continue
            held = g["preds"] >= np.sort(np.unique(g["preds"]))[-3]          # the three highest predicted classes = held-out "OOD"
            # Mahalanobis fitted on the golden CLS features of the ID rows (labels = reference predictions)
            fit = ood_spec.mahalanobis_fit(g["cls"][~held], g["preds"][~held], 32, ridge=1e-3)
            want = {"msp": ood_spec.msp_score(g["logits"]), "energy": ood_spec.energy_score(g["logits"]),
                    "maha": ood_spec.mahalanobis_score(g["cls"], fit)}
            got = {"msp": sc["msp"].cpu().numpy(), "energy": sc["energy"].cpu().numpy(),
                   "maha": ood_spec.mahalanobis_score(sc["cls"].cpu().numpy(), fit)}
            for k in want:
                a0, f0 = ood_spec.auroc(want[k][~held], want[k][held]), ood_spec.fpr_at_tpr_fast(want[k][~held], want[k][held])
                a1, f1 = ood_spec.auroc(got[k][~held], got[k][held]), ood_spec.fpr_at_tpr_fast(got[k][~held], got[k][held])
                print(f"    {k}: AUROC ref {a0:.5f} got {a1:.5f} (same to 3 dec: {round(a0, 3) == round(a1, 3)})  FPR95 ref {f0:.5f} got {f1:.5f} "
                      f"(same: {round(f0, 3) == round(f1, 3)})  n_id {int((~held).sum())} n_ood {int(held.sum())}")


if __name__ == "__main__":
    main()
