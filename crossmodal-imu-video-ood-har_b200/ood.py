"""OOD scoring stage: MSP / energy, Mahalanobis fit + score, AUROC / FPR95.

SPEC-DERIVED (parity unpinned by the reference): /root/reference has no OOD code at all
(SURVEY.md F2, section 8a rows A1-A5).  The definitions implemented here are the literature ones,
restated on CPU in ``oracle/ood_spec.py``; results are "self-consistent with the in-repo spec".
Every scorer returns an OOD SCORE: larger = more out-of-distribution.

Multi-GPU: windows are sharded by rank with no data-path collective; the
only exchanges are the all-reduce of the Mahalanobis sufficient statistics (20 512 doubles) and of
the score histograms, both through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N

__all__ = ["logit_scores", "MahalanobisOOD", "ScoreHistogram", "auroc_fpr95", "finalize_mahalanobis",
           "roc_from_histograms"]

FEAT_DIM = 128


def _prec(precision: Optional[str]) -> int:
    from .models import _prec_code
    return _prec_code(precision)


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ------------------------------------------------------------------------------- A1 / A2
@torch.no_grad()
def logit_scores(logits: torch.Tensor, temperature: float = 1.0) -> Dict[str, torch.Tensor]:
    """From stored logits (n, C): 'pred' int64 arg-max, 'msp' = -max softmax, 'energy' =
    -T logsumexp(logits/T).  One warp per row, HBM-bound."""
    N.require_cuda(logits, "logit_scores")
    z = N.f32c(logits)
    n, c = z.shape
    dev = z.device
    pred = torch.empty(n, dtype=torch.int64, device=dev)
    msp = torch.empty(n, dtype=torch.float32, device=dev)
    energy = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().cmhar_logit_scores(z.data_ptr(), n, c, float(temperature), pred.data_ptr(), msp.data_ptr(),
                                           energy.data_ptr(), N.stream_ptr(dev)))
    return {"pred": pred, "msp": msp, "energy": energy}


# ------------------------------------------------------------------------------- A3 / A4
def _lower_triangular_inverse(chol: np.ndarray) -> np.ndarray:
    """G^{-1} of the Cholesky factor: LAPACK dtrtri (0.17 ms for 128 x 128) when scipy is importable, else the general solve
    (0.5-0.8 ms) -- the finalisation sits inside the timed fit of BASELINE configs[3]."""
    try:
        from scipy.linalg.lapack import dtrtri
        inv, info = dtrtri(np.asfortranarray(chol), lower=1)
        if info == 0:
            return np.tril(inv)
    except ImportError:
        pass
    return np.linalg.solve(chol, np.eye(chol.shape[0]))


def finalize_mahalanobis(count: np.ndarray, ssum: np.ndarray, second: np.ndarray, ridge: float = 0.0
                         ) -> Dict[str, np.ndarray]:
    """Host fp64 finalisation of the (all-reduced) sufficient statistics: class means, tied
    covariance Sigma = (sum f f^T - sum_c n_c mu_c mu_c^T)/N, Cholesky Sigma = G G^T and the
    whitening factor W = G^{-T} so that d_c(f) = ||f W - mu_c W||^2.  128x128: microseconds."""
    count = np.asarray(count, np.float64)
    ssum = np.asarray(ssum, np.float64)
    second = np.asarray(second, np.float64)
    total = count.sum()
    if total <= 0:
        raise ValueError("Mahalanobis fit saw no labelled rows")
    mean = ssum / np.maximum(count, 1.0)[:, None]
    cov = (second - (mean * count[:, None]).T @ mean) / total
    cov = 0.5 * (cov + cov.T)
    if ridge:
        cov = cov + ridge * np.eye(cov.shape[0])
    chol = np.linalg.cholesky(cov)
    whiten = _lower_triangular_inverse(chol).T
    return {"mean": mean, "cov": cov, "whiten": whiten, "mean_whitened": mean @ whiten, "count": count}


class MahalanobisOOD:
    """Class-conditional Gaussian with tied covariance on the 128-d CLS feature (Lee et al. 2018).

    ``accumulate`` may be called any number of times (streamed shards); ``finalize`` all-reduces
    the statistics across ranks when ``torch.distributed`` is initialised, solves the 128x128
    system on the host in fp64 (identically on every rank) and packs the scorer state."""

    def __init__(self, num_classes: int, device="cuda", ridge: float = 0.0):
        self.num_classes, self.ridge = int(num_classes), float(ridge)
        self.device = torch.device(device)
        self.reset()

    def reset(self):
        n = self.num_classes * (1 + FEAT_DIM) + FEAT_DIM * FEAT_DIM
        self._stats = torch.zeros(n, dtype=torch.float64, device=self.device)   # one buffer -> one all-reduce
        self._fit: Optional[Dict[str, np.ndarray]] = None
        self._fit_dev: Optional[Dict[str, torch.Tensor]] = None     # device finalisation: fit64 + the fp32 inputs of cmhar_maha_pack
        self._blobs: Dict[str, torch.Tensor] = {}

    @property
    def fit_(self) -> Optional[Dict[str, np.ndarray]]:
        """The fitted state as host fp64 arrays (mean, cov, whiten, mean_whitened, count); after a device finalisation it is
        read back on first access only (the scorer itself never needs it on the host)."""
        if self._fit is None and self._fit_dev is not None:
            c, d = self.num_classes, FEAT_DIM
            f = self._fit_dev["fit64"].cpu().numpy()
            o = [0, c * d, c * d + d * d, c * d + 2 * d * d, 2 * c * d + 2 * d * d]
            self._fit = {"mean": f[o[0]:o[1]].reshape(c, d), "cov": f[o[1]:o[2]].reshape(d, d), "whiten": f[o[2]:o[3]].reshape(d, d),
                         "mean_whitened": f[o[3]:o[4]].reshape(c, d), "count": self._fit_dev["count64"].cpu().numpy()}
        return self._fit

    @fit_.setter
    def fit_(self, value):
        self._fit, self._fit_dev = value, None

    def _views(self):
        c = self.num_classes
        return (self._stats[:c], self._stats[c:c + c * FEAT_DIM].view(c, FEAT_DIM),
                self._stats[c + c * FEAT_DIM:].view(FEAT_DIM, FEAT_DIM))

    @torch.no_grad()
    def accumulate(self, feats: torch.Tensor, labels: torch.Tensor, precision: Optional[str] = None) -> None:
        """Adds the rows' sufficient statistics.  precision 'bf16' = the tensor-core kernel (split-bf16 GEMMs over the
        row dimension, fp32-grade sums); 'fp32' = CUDA-core FMA tiles."""
        N.require_cuda(feats, "MahalanobisOOD.accumulate")
        f = N.f32c(feats)
        y = labels.to(device=f.device, dtype=torch.int64).contiguous()
        if f.shape[1] != FEAT_DIM or y.shape[0] != f.shape[0]:
            raise ValueError("features must be (n,128) with one label per row")
        cnt, ssum, second = self._views()
        with torch.cuda.device(f.device):
            N.check(N.lib().cmhar_maha_accumulate(f.data_ptr(), y.data_ptr(), f.shape[0], self.num_classes,
                                                  cnt.data_ptr(), ssum.data_ptr(), second.data_ptr(),
                                                  _prec(precision), N.stream_ptr(f.device)))

    def finalize(self, all_reduce: bool = True, on_device: Optional[bool] = None, check: bool = True) -> "MahalanobisOOD":
        """``all_reduce=False`` keeps the fit local to this rank (a scorer fitted on replicated data, or code that only
        one rank executes -- a collective entered by a single rank would hang the others).

        ``on_device`` (default: whenever the statistics live on a CUDA device and classes <= 64): means, tied covariance, Cholesky
        factor, whitening matrix and the packed scorer state are produced by ``cmhar_maha_finalize`` + ``cmhar_maha_pack`` on the
        current stream, right behind the accumulate kernels and the NCCL all-reduce -- no statistics cross PCIe and nothing
        synchronises except the read of one status word (``check``; False leaves even that to the caller: ``finalize_status()``).
        ``on_device=False`` is the host fp64 route (numpy / LAPACK), identical algebra."""
        stats = self._stats
        if all_reduce and _dist_on():
            stats = stats.clone()
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)       # 20 512 doubles over NCCL/NVLink
        if on_device is None:
            on_device = stats.is_cuda and self.num_classes <= 64
        if on_device:
            return self._finalize_device(stats, check)
        host = stats.cpu().numpy()
        c = self.num_classes
        self.fit_ = finalize_mahalanobis(host[:c], host[c:c + c * FEAT_DIM].reshape(c, FEAT_DIM),
                                         host[c + c * FEAT_DIM:].reshape(FEAT_DIM, FEAT_DIM), self.ridge)
        self._blobs.clear()
        from .models import _PACK_GENERATION
        _PACK_GENERATION[0] += 1          # a re-fit replaces the packed scorer state recorded graphs point into
        return self

    def _finalize_device(self, stats: torch.Tensor, check: bool) -> "MahalanobisOOD":
        N.require_cuda(stats, "MahalanobisOOD.finalize(on_device=True)")
        lib, dev, c = N.lib(), stats.device, self.num_classes
        fit64 = torch.empty(lib.cmhar_maha_fit64_doubles(c), dtype=torch.float64, device=dev)
        w32 = torch.empty((FEAT_DIM, FEAT_DIM), dtype=torch.float32, device=dev)
        mw32 = torch.empty((c, FEAT_DIM), dtype=torch.float32, device=dev)
        cnt32 = torch.empty(c, dtype=torch.float32, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        blob = N.alloc_blob(lib.cmhar_maha_blob_bytes(c), dev)
        with torch.cuda.device(dev):
            st = N.stream_ptr(dev)
            N.check(lib.cmhar_maha_finalize(stats.data_ptr(), c, float(self.ridge), fit64.data_ptr(), w32.data_ptr(), mw32.data_ptr(),
                                            cnt32.data_ptr(), info.data_ptr(), st))
            N.check(lib.cmhar_maha_pack(w32.data_ptr(), mw32.data_ptr(), cnt32.data_ptr(), c, blob.data_ptr(), st))
        self._fit, self._blobs = None, {str(dev): blob}
        self._fit_dev = {"fit64": fit64, "count64": stats[:c].clone(), "w32": w32, "mw32": mw32, "cnt32": cnt32, "info": info}
        from .models import _PACK_GENERATION
        _PACK_GENERATION[0] += 1          # a re-fit replaces the packed scorer state recorded graphs point into
        if check:
            self.finalize_status()
        return self

    def finalize_status(self) -> int:
        """Status word of the last device finalisation (one 4-byte D2H read; synchronises).  Raises like the host route."""
        if self._fit_dev is None:
            return 0
        code = int(self._fit_dev["info"].item())
        if code == -1:
            self._fit_dev, self._blobs = None, {}
            raise ValueError("Mahalanobis fit saw no labelled rows")
        if code != 0:
            self._fit_dev, self._blobs = None, {}
            raise np.linalg.LinAlgError(f"Mahalanobis fit: covariance not positive definite (pivot {code}); use a ridge")
        return code

    def fit(self, feats: torch.Tensor, labels: torch.Tensor, all_reduce: bool = True) -> "MahalanobisOOD":
        self.reset()
        self.accumulate(feats, labels)
        return self.finalize(all_reduce=all_reduce)

    def blob(self, device) -> torch.Tensor:
        if self._fit is None and self._fit_dev is None:
            raise RuntimeError("MahalanobisOOD: call fit()/finalize() first")
        key = str(torch.device(device))
        if key not in self._blobs:
            lib = N.lib()
            w = torch.from_numpy(self.fit_["whiten"].astype(np.float32)).to(device).contiguous()
            mw = torch.from_numpy(self.fit_["mean_whitened"].astype(np.float32)).to(device).contiguous()
            cnt = torch.from_numpy(self.fit_["count"].astype(np.float32)).to(device).contiguous()
            blob = N.alloc_blob(lib.cmhar_maha_blob_bytes(self.num_classes), device)
            with torch.cuda.device(device):
                N.check(lib.cmhar_maha_pack(w.data_ptr(), mw.data_ptr(), cnt.data_ptr(), self.num_classes,
                                            blob.data_ptr(), N.stream_ptr(device)))
                torch.cuda.current_stream(device).synchronize()       # w/mw/cnt may be freed after return
            self._blobs[key] = blob
        return self._blobs[key]

    @torch.no_grad()
    def score(self, feats: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
        """min_c Mahalanobis^2 for stored features (n,128).  precision 'bf16' = the tensor-core kernel
        (split-bf16 operands, fp32-grade result)."""
        N.require_cuda(feats, "MahalanobisOOD.score")
        f = N.f32c(feats)
        out = torch.empty(f.shape[0], dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device):
            from .models import _prec_code
            N.check(N.lib().cmhar_maha_score(self.blob(f.device).data_ptr(), f.data_ptr(), f.shape[0],
                                             out.data_ptr(), _prec_code(precision), N.stream_ptr(f.device)))
        return out


# ------------------------------------------------------------------------------- A5
def _float_key(v: np.ndarray) -> np.ndarray:
    b = np.asarray(v, np.float32).view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)


def roc_from_histograms(h_id: np.ndarray, h_ood: np.ndarray, tpr_level: float = 0.95) -> Dict[str, float]:
    """AUROC and FPR@TPR from two aligned histograms over an order-preserving binning of the
    score (bin index increases with the score; OOD = positive class; decision score >= thr).

    Pairs that fall into the same bin are counted as ties (1/2), so
    |AUROC_hist - AUROC_exact| <= tie_mass = sum_b id_b*ood_b / (2 n_id n_ood); FPR is evaluated
    at bin granularity, |FPR_hist - FPR_exact| <= id_{b*}/n_id for the crossing bin b*.  Both
    bounds are returned; they are 0 when every bin holds a single distinct score."""
    h_id = np.asarray(h_id, np.float64)
    h_ood = np.asarray(h_ood, np.float64)
    n_id, n_ood = h_id.sum(), h_ood.sum()
    if n_id == 0 or n_ood == 0:
        return {"auroc": float("nan"), "fpr": float("nan"), "auroc_bound": 0.0, "fpr_bound": 0.0}
    below = np.cumsum(h_id) - h_id                       # ID mass strictly below each bin
    auroc = float(((below + 0.5 * h_id) * h_ood).sum() / (n_id * n_ood))
    tie = float((h_id * h_ood).sum() / (2.0 * n_id * n_ood))
    tp_ge = np.cumsum(h_ood[::-1])[::-1]                 # OOD mass at or above each bin
    fp_ge = np.cumsum(h_id[::-1])[::-1]
    need = tpr_level * n_ood - 1e-9
    ok = np.flatnonzero(tp_ge >= need)
    b = int(ok.max()) if ok.size else 0                  # highest threshold bin reaching the TPR
    return {"auroc": auroc, "fpr": float(fp_ge[b] / n_id), "auroc_bound": tie,
            "fpr_bound": float(h_id[b] / n_id), "crossing_bin": b}


class ScoreHistogram:
    """Order-preserving histogram of float scores accumulated on the device, one instance per
    population (ID / OOD).  ``bins`` buckets span [key_lo, key_hi] of the monotone uint32 image of
    the float; ranks share ``key_lo``/``shift`` so histograms add across ranks."""

    def __init__(self, key_lo: int, shift: int, bins: int, device="cuda"):
        self.key_lo, self.shift, self.bins = int(key_lo), int(shift), int(bins)
        self.hist = torch.zeros(bins, dtype=torch.int64, device=device)

    @torch.no_grad()
    def add(self, scores: torch.Tensor) -> None:
        s = N.f32c(scores)
        with torch.cuda.device(s.device):
            N.check(N.lib().cmhar_score_histogram(s.data_ptr(), s.numel(), self.key_lo, self.shift, self.bins,
                                                  self.hist.data_ptr(), N.stream_ptr(s.device)))


@torch.no_grad()
def score_key_range(scores: torch.Tensor, mm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Running (min,max) of the monotone uint32 keys; returns the int64[2] device tensor."""
    s = N.f32c(scores)
    if mm is None:
        mm = torch.tensor([0xFFFFFFFF, 0], dtype=torch.int64, device=s.device)
    tmp = torch.tensor([-1, 0], dtype=torch.int32, device=s.device)     # {0xffffffff, 0} as uint32
    with torch.cuda.device(s.device):
        N.check(N.lib().cmhar_score_key_range(s.data_ptr(), s.numel(), tmp.data_ptr(), N.stream_ptr(s.device)))
    lohi = tmp.to(torch.int64) & 0xFFFFFFFF
    mm[0] = torch.minimum(mm[0], lohi[0])
    mm[1] = torch.maximum(mm[1], lohi[1])
    return mm


@torch.no_grad()
def auroc_fpr95(scores_id: torch.Tensor, scores_ood: torch.Tensor, bins: int = 1 << 16,
                tpr_level: float = 0.95, refine: bool = True) -> Dict[str, float]:
    """AUROC / FPR95 of device-resident score shards (OOD = positive).

    Pass 1 finds the global key range (all-reduced MIN/MAX when distributed), pass 2 builds the two
    histograms (all-reduced SUM).  With ``refine`` the bin where the TPR crosses is re-histogrammed
    at full key resolution so FPR95 is exact, and the AUROC error bound (tie mass) is returned in
    'auroc_bound' -- at 2^16 bins it is ~1e-5 for continuous scores, far below the 3-decimal
    contract."""
    N.require_cuda(scores_id, "auroc_fpr95")
    dev = scores_id.device
    mm = score_key_range(scores_id)
    mm = score_key_range(scores_ood, mm)
    if _dist_on():
        lo, hi = mm[0:1].clone(), mm[1:2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        mm = torch.cat([lo, hi])
    lo, hi = (int(v) for v in mm.tolist())
    if hi < lo:
        return {"auroc": float("nan"), "fpr": float("nan"), "auroc_bound": 0.0, "fpr_bound": 0.0}
    span = hi - lo + 1
    shift = 0
    while (span >> shift) > bins:
        shift += 1

    def hist_pair(key_lo, sh, nb):
        h_id, h_ood = ScoreHistogram(key_lo, sh, nb, dev), ScoreHistogram(key_lo, sh, nb, dev)
        h_id.add(scores_id)
        h_ood.add(scores_ood)
        both = torch.stack([h_id.hist, h_ood.hist])
        if _dist_on():
            dist.all_reduce(both, op=dist.ReduceOp.SUM)
        both = both.cpu().numpy()
        return both[0], both[1]

    nb = (span >> shift) + 1
    a, b = hist_pair(lo, shift, nb)
    res = roc_from_histograms(a, b, tpr_level)
    if refine and shift > 0 and res["fpr_bound"] > 0:
        # exact FPR: resolve the crossing bin at key granularity (2^shift sub-bins, values outside
        # the bin clamp into sentinel bins 0 and last, which are discarded)
        cb = res["crossing_bin"]
        sub_lo = lo + (cb << shift)
        sub_n = (1 << shift) + 2
        if sub_n <= (1 << 22) and sub_lo >= 1:
            sa, sb = hist_pair(sub_lo - 1, 0, sub_n)
            sa, sb = sa[1:-1], sb[1:-1]
            n_id, n_ood = a.sum(), b.sum()
            tp_above, fp_above = b[cb + 1:].sum(), a[cb + 1:].sum()
            tp_ge = tp_above + np.cumsum(sb[::-1])[::-1]
            fp_ge = fp_above + np.cumsum(sa[::-1])[::-1]
            ok = np.flatnonzero((tp_ge >= tpr_level * n_ood - 1e-9) & ((sa + sb) > 0))
            if ok.size:
                k = int(ok.max())
                res["fpr"], res["fpr_bound"] = float(fp_ge[k] / n_id), 0.0
    res.pop("crossing_bin", None)
    res["bins"], res["shift"] = int(nb), int(shift)
    return res
