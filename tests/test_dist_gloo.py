"""world_size-2 gloo tests (CPU) of the only two exchanges on the path:
the all-reduce of the Mahalanobis sufficient statistics and of the ROC score histograms.  Rows are
sharded by rank with ``shard_bounds``; the reductions must reproduce the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import crossmodal_imu_video_ood_har_b200 as cm
from crossmodal_imu_video_ood_har_b200 import ood
from oracle import ood_spec, weights as W


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- Mahalanobis fit: each rank accumulates its shard, finalize() all-reduces
        feats, labels = W.class_features(7, 4001)
        lo, hi = cm.shard_bounds(len(feats), rank, world)
        m = cm.MahalanobisOOD(32, device="cpu")
        n, s, ff = ood_spec.mahalanobis_sufficient_stats(feats[lo:hi], labels[lo:hi], 32)  # stands in for the CUDA accumulate
        cnt, ssum, second = m._views()
        cnt.copy_(torch.from_numpy(n)); ssum.copy_(torch.from_numpy(s)); second.copy_(torch.from_numpy(ff))
        m.finalize()
        # ---- ROC histograms: per-rank histograms over a shared binning, summed over ranks
        q, ql = W.class_features(8, 3000, ood_fraction=0.4)
        fit = ood_spec.mahalanobis_fit(feats, labels, 32)
        sc = ood_spec.mahalanobis_score(q, fit).astype(np.float32)
        lo2, hi2 = cm.shard_bounds(len(sc), rank, world)
        keys = ood._float_key(sc[lo2:hi2]).astype(np.int64)
        is_ood = (ql[lo2:hi2] < 0)
        mm = torch.tensor([keys.min(), keys.max()])
        lo_k, hi_k = mm[0:1].clone(), mm[1:2].clone()
        dist.all_reduce(lo_k, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_k, op=dist.ReduceOp.MAX)
        lo_k, hi_k = int(lo_k), int(hi_k)
        shift = 0
        while ((hi_k - lo_k + 1) >> shift) > (1 << 16):
            shift += 1
        nb = ((hi_k - lo_k + 1) >> shift) + 1
        h = torch.zeros(2, nb, dtype=torch.int64)
        h[0] = torch.from_numpy(np.bincount((keys[~is_ood] - lo_k) >> shift, minlength=nb))
        h[1] = torch.from_numpy(np.bincount((keys[is_ood] - lo_k) >> shift, minlength=nb))
        dist.all_reduce(h)
        roc = cm.roc_from_histograms(h[0].numpy(), h[1].numpy())
        if rank == 0:
            np.savez(os.path.join(out_dir, "r0.npz"), mean=m.fit_["mean"], whiten=m.fit_["whiten"],
                     auroc=roc["auroc"], fpr=roc["fpr"], bound=roc["auroc_bound"], fbound=roc["fpr_bound"])
    finally:
        dist.destroy_process_group()


def test_two_rank_reductions_match_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "r0.npz")
    feats, labels = W.class_features(7, 4001)
    fit = ood_spec.mahalanobis_fit(feats, labels, 32)
    np.testing.assert_allclose(got["mean"], fit["mean"], atol=1e-10)
    np.testing.assert_allclose(got["whiten"] @ got["whiten"].T, fit["precision"], rtol=1e-7, atol=1e-9)
    q, ql = W.class_features(8, 3000, ood_fraction=0.4)
    sc = ood_spec.mahalanobis_score(q, fit).astype(np.float32)
    want_auc = ood_spec.auroc(sc[ql >= 0], sc[ql < 0])
    want_fpr = ood_spec.fpr_at_tpr_fast(sc[ql >= 0], sc[ql < 0])
    assert abs(float(got["auroc"]) - want_auc) <= float(got["bound"]) + 1e-12
    assert abs(float(got["auroc"]) - want_auc) < 5e-4          # 3-decimal contract
    assert abs(float(got["fpr"]) - want_fpr) <= float(got["fbound"]) + 1e-12
