"""Full evaluation sweep: ID vs held-out-activity OOD splits, all scorers, results table (BASELINE.json configs[4];
SURVEY.md section 8a rows A1-A5, 8f.2).

SPEC-DERIVED (parity unpinned by the reference): the reference has no OOD evaluation and its dataset code has no
held-out-activity split (src/data/datasets.py:256-337 builds few-shot splits only).  ``held_out_activity_split`` defines
the split the north star names: windows whose activity label is in ``held_out`` are the OOD population and never reach
the Mahalanobis fit; everything else is in-distribution.

``OODSweep`` runs the sweep with every score device resident until the metrics are final:

    fit pass    windows -> encoder + head launch (CLS) -> Mahalanobis accumulate (ID rows only) -> all-reduce + finalise
    score pass  windows -> ONE encoder + head launch emits pred / MSP / energy / Mahalanobis -> kept on the device
    metrics     per scorer: key range + histograms on the device, all-reduced across ranks, AUROC / FPR95
    table       rows for ``tables.generate_ood_table`` (rank 0 writes ood_results.csv + the csv / tex / md tables)

Windows are sharded by rank (``evaluator.shard_bounds``); the only exchanges are the statistics all-reduce of the fit
and the MIN/MAX + SUM all-reduces of the histograms.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .models import IMUClassifier
from .ood import MahalanobisOOD, auroc_fpr95

__all__ = ["held_out_activity_split", "OODSweep"]


def held_out_activity_split(labels, held_out: Sequence[int]):
    """(id_mask, ood_mask) for activity ``labels`` (numpy array or torch tensor, any device): rows whose label is in
    ``held_out`` are out-of-distribution.  Labels < 0 (unlabelled) are in neither population."""
    held = sorted(set(int(h) for h in held_out))
    if isinstance(labels, torch.Tensor):
        y = labels.reshape(-1)
        ood = torch.zeros_like(y, dtype=torch.bool)
        for h in held:
            ood |= (y == h)
        return (~ood) & (y >= 0), ood
    y = np.asarray(labels).reshape(-1)
    ood = np.isin(y, held)
    return (~ood) & (y >= 0), ood


class OODSweep:
    """``OODSweep(classifier, held_out=[...])``; ``fit(train_batches)``, ``score(test_batches)``, ``metrics()``.

    Batches are ``(windows, labels)`` pairs of CUDA tensors -- windows in the reference layout (B,6,L) or compact
    (B, live) with ``window_stride`` -- i.e. this rank's shard of the data, already on the device (an evaluator loop
    over host data feeds ``CrossModalOODPipeline.stream_host`` / ``Evaluator`` instead)."""

    SCORERS = ("msp", "energy", "maha")

    def __init__(self, classifier: IMUClassifier, held_out: Sequence[int], precision: Optional[str] = None,
                 ridge: float = 1e-3, window_stride: Optional[int] = None):
        self.clf = classifier.eval()
        self.held_out = sorted(set(int(h) for h in held_out))
        self.precision, self.ridge, self.window_stride = precision, ridge, window_stride
        self.device = next(classifier.parameters()).device
        self.maha: Optional[MahalanobisOOD] = None
        self._scores: Dict[str, List[torch.Tensor]] = {}
        self._labels: List[torch.Tensor] = []
        self._correct = torch.zeros(2, dtype=torch.int64, device=self.device)      # (hits, ID rows) of the classifier
        self.windows_seen = 0

    @torch.no_grad()
    def fit(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> MahalanobisOOD:
        """Mahalanobis class statistics from the ID rows of the training shard (held-out activities excluded)."""
        maha = MahalanobisOOD(self.clf.num_classes, self.device, self.ridge)
        self.clf.set_mahalanobis(None)
        for x, y in batches:
            res = self.clf.forward_scores(x, precision=self.precision, want_cls=True, want_logits=False,
                                          window_stride=self.window_stride)
            id_mask, _ = held_out_activity_split(y, self.held_out)
            lab = torch.where(id_mask, y, torch.full_like(y, -1))          # rows labelled -1 are skipped by the kernel
            maha.accumulate(res["cls"], lab, precision=self.precision)
            self.windows_seen += int(x.shape[0])
        maha.finalize()                                                     # all-reduce across ranks inside
        self.clf.set_mahalanobis(maha)
        self.maha = maha
        return maha

    @torch.no_grad()
    def score(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> None:
        """One fused launch pair per batch; scores, labels and the hit count stay on the device.  Nothing here synchronises
        (no boolean indexing per batch): the populations are separated once, in ``metrics``."""
        for x, y in batches:
            res = self.clf.forward_scores(x, precision=self.precision, window_stride=self.window_stride)
            id_mask, _ = held_out_activity_split(y, self.held_out)
            for k in self.SCORERS:
                if k in res:
                    self._scores.setdefault(k, []).append(res[k])
            self._labels.append(y)
            self._correct[0] += ((res["pred"] == y) & id_mask).sum()
            self._correct[1] += id_mask.sum()
            self.windows_seen += int(x.shape[0])

    @torch.no_grad()
    def metrics(self) -> Dict[str, Dict[str, float]]:
        """{scorer: {auroc, fpr95, auroc_bound}} over ALL ranks' shards (collective when distributed) + 'accuracy'."""
        import torch.distributed as dist
        y = torch.cat(self._labels) if self._labels else torch.zeros(0, dtype=torch.int64, device=self.device)
        id_mask, ood = held_out_activity_split(y, self.held_out)
        table: Dict[str, Dict[str, float]] = {}
        for k in self.SCORERS:
            if k not in self._scores:
                continue
            s = torch.cat(self._scores[k])
            r = auroc_fpr95(s[id_mask].contiguous(), s[ood].contiguous())
            table[k] = {"auroc": r["auroc"], "fpr95": r["fpr"], "auroc_bound": r["auroc_bound"]}
        c = self._correct.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(c)
        hits, n = (int(v) for v in c.tolist())
        self.accuracy = 100.0 * hits / max(n, 1)
        return table
