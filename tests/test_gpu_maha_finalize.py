"""GPU tests (-m gpu) of the device-side Mahalanobis finalisation (cmhar_maha_finalize: means, tied covariance, Cholesky, whitening
factor in one fp64 CTA, stream-ordered in front of cmhar_maha_pack) against the host fp64 route it replaces (ood.finalize_mahalanobis)
and the float64 spec (oracle/ood_spec.py).  Row A3 -- no reference implementation exists: parity unpinned, spec oracle."""
import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from crossmodal_imu_video_ood_har_b200.ood import finalize_mahalanobis
from oracle import ood_spec, weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("classes,ridge,n", [(32, 0.0, 20011), (32, 1e-3, 4000), (7, 0.0, 3001), (64, 1e-2, 9000)])
def test_device_finalize_matches_host_and_spec(classes, ridge, n):
    feats, labels = W.class_features(classes + 1, n, num_classes=classes)
    labels[::53] = -1
    if classes > 4:
        labels[labels == 3] = 2                                  # an empty class: excluded from the scores, means stay finite
    f, y = torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV)
    dev_fit = cm.MahalanobisOOD(classes, DEV, ridge=ridge)
    dev_fit.accumulate(f, y)
    host_fit = cm.MahalanobisOOD(classes, DEV, ridge=ridge)
    host_fit._stats.copy_(dev_fit._stats)
    dev_fit.finalize(all_reduce=False)                           # default on a CUDA device: the device route
    host_fit.finalize(all_reduce=False, on_device=False)
    assert dev_fit._fit_dev is not None and host_fit._fit_dev is None
    d, h = dev_fit.fit_, host_fit.fit_
    np.testing.assert_array_equal(d["count"], h["count"])
    for k in ("mean", "cov", "whiten", "mean_whitened"):
        assert d[k].shape == h[k].shape and rel(d[k], h[k]) < 1e-10, (k, rel(d[k], h[k]))
    assert np.allclose(d["whiten"], np.triu(d["whiten"]))       # G^-T is upper triangular
    # W W^T = Sigma^-1
    assert rel(d["whiten"] @ d["whiten"].T @ d["cov"], np.eye(128)) < 1e-8
    spec = ood_spec.mahalanobis_fit(feats, labels, classes, ridge=ridge)
    assert rel(d["mean"], spec["mean"]) < 1e-5 and rel(d["cov"], spec["cov"]) < 1e-4
    q, _ = W.class_features(99, 3000, num_classes=classes, ood_fraction=0.3)
    qd = torch.from_numpy(q).to(DEV)
    for prec, tol in (("fp32", 1e-3), ("bf16", 1e-3)) if classes <= 32 else (("fp32", 1e-3),):
        got, ref = dev_fit.score(qd, precision=prec), host_fit.score(qd, precision=prec)
        assert rel(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5
        assert rel(got.cpu().numpy(), ood_spec.mahalanobis_score(q, spec)) < tol
    # deterministic: a second finalisation of the same statistics gives the same bits
    again = cm.MahalanobisOOD(classes, DEV, ridge=ridge)
    again._stats.copy_(dev_fit._stats)
    again.finalize(all_reduce=False)
    np.testing.assert_array_equal(again.fit_["whiten"], d["whiten"])
    assert torch.equal(again.score(qd), dev_fit.score(qd))


def test_device_finalize_errors_and_deferred_status():
    m = cm.MahalanobisOOD(32, DEV)
    with pytest.raises(ValueError, match="no labelled rows"):
        m.finalize(all_reduce=False)
    feats = np.ones((500, 128), np.float32)                      # rank-deficient covariance, no ridge
    labels = (np.arange(500) % 32).astype(np.int64)
    m.reset()
    m.accumulate(torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV))
    m.finalize(all_reduce=False, check=False)                    # nothing synchronises here ...
    with pytest.raises(np.linalg.LinAlgError):
        m.finalize_status()                                      # ... the status word is read on demand
    with pytest.raises(RuntimeError, match="fit"):
        m.blob(DEV)
    m.ridge = 1e-2
    m.finalize(all_reduce=False)
    assert m.finalize_status() == 0 and np.isfinite(m.fit_["whiten"]).all()
    lib = cm._native.lib()
    assert lib.cmhar_maha_fit64_doubles(65) == 0 and lib.cmhar_maha_fit64_doubles(32) == 2 * 32 * 128 + 2 * 128 * 128
    buf = torch.zeros(64, dtype=torch.float64, device=DEV)
    assert lib.cmhar_maha_finalize(buf.data_ptr(), 65, 0.0, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(),
                                   cm._native.stream_ptr(DEV)) != 0
    assert lib.cmhar_maha_finalize(buf.data_ptr(), 8, -1.0, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(),
                                   cm._native.stream_ptr(DEV)) != 0
