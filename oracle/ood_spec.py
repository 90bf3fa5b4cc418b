"""TEST INFRASTRUCTURE ONLY -- spec-derived oracle for the OOD rows A1-A6.

PARITY UNPINNED BY THE REFERENCE: /root/reference contains no OOD scorer, no Mahalanobis fit,
no AUROC/FPR95 and no fusion classifier.  These
functions restate the standard literature definitions that BASELINE.json's north_star names;
where a library definition exists they are cross-checked against it in
``tests/test_oracle_vs_golden.py`` (scipy ``logsumexp``/``softmax``,
``sklearn.metrics.roc_auc_score`` / ``roc_curve``).  Every result derived from them is
"self-consistency with the in-repo spec", not reference parity.

Sign convention used throughout the repo: every scorer returns an OOD SCORE, larger = more
out-of-distribution.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


# ---- A1: maximum softmax probability (Hendrycks & Gimpel 2017) ---------------------------
def msp_score(logits: np.ndarray) -> np.ndarray:
    """OOD score = -max_c softmax(logits)_c."""
    z = np.asarray(logits, dtype=np.float64)
    z = z - z.max(1, keepdims=True)
    p = np.exp(z)
    return -(p.max(1) / p.sum(1))


# ---- A2: energy (Liu et al. 2020) --------------------------------------------------------
def energy_score(logits: np.ndarray, T: float = 1.0) -> np.ndarray:
    """OOD score = E(x) = -T * logsumexp(logits / T) (computed with max subtraction)."""
    z = np.asarray(logits, dtype=np.float64) / T
    m = z.max(1)
    return -T * (m + np.log(np.exp(z - m[:, None]).sum(1)))


# ---- A3: Mahalanobis fit (Lee et al. 2018, tied covariance) ------------------------------
def mahalanobis_sufficient_stats(feats: np.ndarray, labels: np.ndarray, num_classes: int
                                 ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per-class counts n_c (C,), per-class sums (C,D) and the second moment sum f f^T (D,D),
    all in float64.  Rows with a label outside [0, C) are ignored.  These are exactly the
    quantities the multi-GPU fit all-reduces."""
    f = np.asarray(feats, dtype=np.float64)
    y = np.asarray(labels).astype(np.int64)
    keep = (y >= 0) & (y < num_classes)
    f, y = f[keep], y[keep]
    n = np.bincount(y, minlength=num_classes).astype(np.float64)
    s = np.zeros((num_classes, f.shape[1]))
    np.add.at(s, y, f)
    return n, s, f.T @ f


def mahalanobis_finalize(n: np.ndarray, s: np.ndarray, ff: np.ndarray, ridge: float = 0.0
                         ) -> Dict[str, np.ndarray]:
    """mu_c = s_c / n_c; tied Sigma = (sum f f^T - sum_c n_c mu_c mu_c^T) / N (+ ridge*I);
    precision factor via Cholesky: Sigma = G G^T,  W = G^{-T}  so that
    (f-mu)^T Sigma^{-1} (f-mu) = || W^T (f - mu) ||^2  ... stored as ``whiten`` (D,D) with
    rows-of-features convention:  dist = || f @ whiten - mu @ whiten ||^2."""
    N = n.sum()
    mu = s / np.maximum(n, 1.0)[:, None]
    sigma = (ff - (mu * n[:, None]).T @ mu) / N
    sigma = 0.5 * (sigma + sigma.T) + ridge * np.eye(sigma.shape[0])
    G = np.linalg.cholesky(sigma)                 # Sigma = G G^T
    whiten = np.linalg.inv(G).T                   # Sigma^{-1} = whiten @ whiten^T
    return {"mean": mu, "cov": sigma, "precision": whiten @ whiten.T, "whiten": whiten,
            "mean_whitened": mu @ whiten, "count": n}


def mahalanobis_fit(feats, labels, num_classes: int, ridge: float = 0.0):
    return mahalanobis_finalize(*mahalanobis_sufficient_stats(feats, labels, num_classes), ridge)


# ---- A4: Mahalanobis score --------------------------------------------------------------
def mahalanobis_score(feats: np.ndarray, fit: Dict[str, np.ndarray]) -> np.ndarray:
    """OOD score = min_c (f-mu_c)^T Sigma^{-1} (f-mu_c), evaluated directly in float64.
    Classes with zero training count are excluded."""
    f = np.asarray(feats, dtype=np.float64)
    out = np.full(f.shape[0], np.inf)
    for c in range(fit["mean"].shape[0]):
        if fit["count"][c] <= 0:
            continue
        d = f - fit["mean"][c]
        out = np.minimum(out, np.einsum("nd,de,ne->n", d, fit["precision"], d))
    return out


# ---- A5: AUROC / FPR95 -------------------------------------------------------------------
def auroc(scores_id: np.ndarray, scores_ood: np.ndarray) -> float:
    """P(score_ood > score_id) + 0.5 P(tie): Mann-Whitney U with mid-ranks, OOD = positive class.
    Equals sklearn.metrics.roc_auc_score([0]*n_id + [1]*n_ood, scores)."""
    a = np.asarray(scores_id, dtype=np.float64)
    b = np.asarray(scores_ood, dtype=np.float64)
    allv = np.concatenate([a, b])
    order = np.argsort(allv, kind="mergesort")
    sv = allv[order]
    ranks = np.empty(len(allv), dtype=np.float64)
    # mid-ranks for ties
    boundaries = np.flatnonzero(np.concatenate([[True], sv[1:] != sv[:-1], [True]]))
    for lo, hi in zip(boundaries[:-1], boundaries[1:]):
        ranks[order[lo:hi]] = 0.5 * (lo + hi - 1) + 1.0
    r_pos = ranks[len(a):].sum()
    n_pos, n_neg = len(b), len(a)
    return float((r_pos - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def fpr_at_tpr(scores_id: np.ndarray, scores_ood: np.ndarray, tpr: float = 0.95) -> float:
    """FPR at the first (highest) threshold whose TPR >= ``tpr``; OOD = positive class, decision
    ``score >= threshold``.  Equals ``fpr[np.argmax(tpr_curve >= tpr)]`` from
    sklearn.metrics.roc_curve (thresholds are the distinct score values, descending)."""
    a = np.asarray(scores_id, dtype=np.float64)
    b = np.asarray(scores_ood, dtype=np.float64)
    for thr in np.unique(np.concatenate([a, b]))[::-1]:
        if (b >= thr).mean() >= tpr:
            return float((a >= thr).mean())
    return 1.0


def fpr_at_tpr_fast(scores_id, scores_ood, tpr: float = 0.95) -> float:
    """Same definition, O(n log n): the smallest k such that k of the sorted-descending OOD scores
    reach the TPR gives threshold = that score."""
    a = np.sort(np.asarray(scores_id, dtype=np.float64))
    b = np.sort(np.asarray(scores_ood, dtype=np.float64))[::-1]
    k = int(np.ceil(tpr * len(b) - 1e-12))
    k = max(k, 1)
    thr = b[k - 1]
    return float((len(a) - np.searchsorted(a, thr, side="left")) / len(a))


# ---- A6: late-fusion (concat-MLP) classifier -- spec-defined, see DESIGN.md ---------------
def late_fusion_logits(imu_feat, video_feat, sd, prefix: str = "fusion."):
    """logits = W2 relu(BN_eval(W1 [imu_feat ; video_feat] + b1)) + b2.
    No reference definition exists; this IS the definition."""
    x = np.concatenate([np.asarray(imu_feat, np.float64), np.asarray(video_feat, np.float64)], 1)
    t = x @ sd[prefix + "net.0.weight"].astype(np.float64).T + sd[prefix + "net.0.bias"]
    t = (t - sd[prefix + "net.1.running_mean"]) / np.sqrt(sd[prefix + "net.1.running_var"] + 1e-5)
    t = np.maximum(t * sd[prefix + "net.1.weight"] + sd[prefix + "net.1.bias"], 0.0)
    return t @ sd[prefix + "net.3.weight"].astype(np.float64).T + sd[prefix + "net.3.bias"]
