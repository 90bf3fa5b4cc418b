"""GPU tests (-m gpu) of the fused tail kernels of the bf16 step: each must reproduce the chain of launches it replaces --
bit for bit where the arithmetic is the same sequence of MMAs, to round-off where only a reduction order differs -- and the
float64 oracle within the bf16 contract.

  cmhar_imu_forward_ex(cls_img)   CLS rows also as a bf16 operand image
  cmhar_mlp2_forward_img          Linear -> BN -> ReLU -> Linear -> L2 normalise       (reference models.py:226-234, 288-289)
  cmhar_fused_head_forward        late-fusion layer + head + scores                    (spec row A6)
  cmhar_similarity_img            sigmoid loss from operand images, partitioned B       (reference losses.py:37-52)
  cmhar_peer_barrier              single-rank degenerate case (the 2-GPU case: tests/test_gpu_multi.py)
"""
import numpy as np
import pytest
import torch

import crossmodal_imu_video_ood_har_b200 as cm
from crossmodal_imu_video_ood_har_b200.losses import similarity_img_native, similarity_img_work
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native, l2_normalize_native, operand_image
from oracle import fusion_spec, oracle, weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N = cm._native


def tsd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def modules(seed=41):
    cfg = cm.default_config()
    sd, sd_x = fusion_spec.fusion_state(seed), W.cross_modal_state(seed)
    xm = cm.CrossModalModel(cfg)
    xm.load_state_dict(tsd(sd_x), strict=True)
    fus = cm.LateFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg)
    own = set(fus.state_dict().keys())
    fus.load_state_dict({k: v for k, v in tsd(sd).items() if k in own}, strict=True)
    return xm.to(DEV).eval(), fus.to(DEV).eval(), sd, sd_x


def image_of(x: torch.Tensor) -> torch.Tensor:
    """fp32 rows -> bf16 operand image through a kernel that is tested elsewhere (cmhar_linear_forward_img's writer is not
    usable without a layer, so: the similarity pre-pass's layout restated on the host)."""
    n, dim = x.shape
    tiles, kcs = (n + 127) // 128, dim // 64
    xb = torch.zeros(tiles * 128, dim, dtype=torch.bfloat16, device=x.device)
    xb[:n] = x.to(torch.bfloat16)
    r = torch.arange(128, device=x.device)
    img = torch.zeros(tiles, kcs, 128 * 64, dtype=torch.bfloat16, device=x.device)
    for j in range(8):                                  # 16-byte piece j of row r sits at piece (j ^ (r & 7))
        dst = (r // 8) * 512 + (r % 8) * 64 + ((j ^ (r % 8)) * 8)
        for e in range(8):
            img[:, :, dst + e] = xb.view(tiles, 128, kcs, 64)[:, :, :, j * 8 + e].transpose(1, 2)
    out = operand_image(n, dim, x.device)
    out.copy_(img.view(torch.uint8).reshape(-1))
    return out


def rows_of(img: torch.Tensor, n: int, dim: int) -> torch.Tensor:
    """Inverse of ``image_of``: bf16 operand image -> (n, dim) fp32 rows (host-side restatement of the layout)."""
    tiles, kcs = (n + 127) // 128, dim // 64
    flat = img[: tiles * kcs * 16384].view(torch.bfloat16).view(tiles, kcs, 128 * 64)
    r = torch.arange(128, device=img.device)
    out = torch.zeros(tiles, 128, kcs, 64, dtype=torch.float32, device=img.device)
    for j in range(8):
        src = (r // 8) * 512 + (r % 8) * 64 + ((j ^ (r % 8)) * 8)
        for e in range(8):
            out[:, :, :, j * 8 + e] = flat[:, :, src + e].transpose(1, 2).float()
    return out.reshape(tiles * 128, dim)[:n]


@pytest.mark.parametrize("n", [5, 128, 300])
def test_encoder_cls_image_and_fused_projection_head(n):
    xm, fus, sd, sd_x = modules()
    x = torch.from_numpy(W.imu_windows(n, n)).to(DEV)
    out = imu_forward_native(xm.imu_encoder, None, None, x, want_cls=True, precision="bf16", want_cls_img=True)
    torch.cuda.synchronize()
    assert _images_equal_on_valid_rows(out["cls_img"], image_of(out["cls"]), n, 128)
    for head, k, src in ((xm.imu_proj, 128, out["cls"]), (xm.video_proj, 768, None)):
        if src is None:
            src = torch.from_numpy(np.random.RandomState(n).standard_normal((n, k)).astype(np.float32)).to(DEV)
        img = image_of(src)
        chain_rows = head.forward_native(src, "bf16")                       # two linear_tc launches
        rows, y_img = head.forward_fused(img, n, normalize=False)
        torch.cuda.synchronize()
        assert torch.equal(rows, chain_rows)                                # same MMAs in the same order
        nrm, n_img = head.forward_fused(img, n, normalize=True)
        want = l2_normalize_native(chain_rows)
        torch.cuda.synchronize()
        assert torch.allclose(nrm, want, rtol=0, atol=2e-7)
        assert _images_equal_on_valid_rows(n_img, image_of(nrm), n, 256)
        assert torch.allclose(nrm.norm(dim=1), torch.ones(n, device=DEV), atol=1e-5)


def _images_equal_on_valid_rows(a, b, n, dim):
    """Operand images agree on the rows < n (rows past n inside the last tile are zero or untouched)."""
    kcs = dim // 64
    av = a.view(torch.int16).view(-1, kcs, 128 * 64)
    bv = b.view(torch.int16).view(-1, kcs, 128 * 64)
    r = torch.arange(128, device=a.device)
    ok = True
    for j in range(8):
        base = (r // 8) * 512 + (r % 8) * 64 + ((j ^ (r % 8)) * 8)
        for e in range(8):
            ca, cb = av[:, :, base + e], bv[:, :, base + e]                # (tiles, kcs, 128 rows)
            rows = (torch.arange(av.shape[0], device=a.device)[:, None, None] * 128 + r[None, None, :]) < n
            ok = ok and bool(torch.equal(ca[rows.expand_as(ca)], cb[rows.expand_as(cb)]))
    return ok


@pytest.mark.parametrize("n,with_maha", [(77, True), (256, True), (300, False)])
def test_fused_head_equals_chain_and_spec(n, with_maha):
    xm, fus, sd, sd_x = modules()
    T = 16
    imu, fmap = W.imu_windows(3, n), W.video_feature_maps(4, n, T)
    f_dev = torch.from_numpy(fmap).to(DEV).to(torch.bfloat16)
    if with_maha:
        feats, labels = W.class_features(7, 2000)
        fus.set_mahalanobis(cm.MahalanobisOOD(32, DEV, ridge=1e-3).fit(torch.from_numpy(feats).to(DEV), torch.from_numpy(labels).to(DEV)))
    x = torch.from_numpy(imu).to(DEV)
    enc = imu_forward_native(xm.imu_encoder, None, None, x, want_cls=True, precision="bf16", want_cls_img=True)
    ve = xm.video_encoder
    _, pimg = ve.pool_features(f_dev, T, want_img=True, want_rows=False)
    vfeat, vimg = ve._packed_projection(DEV).forward_img(n, False, x_img=pimg, want_rows=True, want_img=True)
    chain = fus.forward_scores(None, None, T, precision="bf16", imu_cls=enc["cls"], video_feat=vfeat)    # concat-linear + head_tc
    got = fus.forward_scores_img(enc["cls_img"], vimg, n)
    torch.cuda.synchronize()
    assert got is not None
    for k in ("fused", "logits", "pred", "msp", "energy") + (("maha",) if with_maha else ()):
        assert torch.equal(got[k], chain[k]), k
    f_r = f_dev.float().cpu().numpy()
    want, want_f = fusion_spec.late_fusion(imu, f_r, sd, T, dtype=torch.float64)
    rel = lambda g, w: float(np.abs(g.detach().cpu().numpy().astype(np.float64) - w.numpy()).max() / np.abs(w.numpy()).max())
    assert rel(got["fused"], want_f) < 2e-2 and rel(got["logits"], want) < 2e-2


@pytest.mark.parametrize("na,nb", [(256, 256), (300, 200), (1024, 4096)])
def test_similarity_from_operand_images(na, nb):
    g = torch.Generator(device=DEV).manual_seed(na + nb)
    a = torch.nn.functional.normalize(torch.randn(na, 256, device=DEV, generator=g), dim=1)
    b = torch.nn.functional.normalize(torch.randn(nb, 256, device=DEV, generator=g), dim=1)
    a_img, b_img = image_of(a), image_of(b)
    ref = cm.similarity_native(a, b, sigmoid=(10.0, -10.0), precision="bf16")["sigmoid_sum"] / (na * nb)
    work = similarity_img_work(na, nb, DEV)
    loss = similarity_img_native(a_img, na, b_img, nb, 256, work=work)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 1e-9 * abs(float(ref))
    first = float(loss)
    for _ in range(3):                                   # the ticket re-arms itself; the reduction order is fixed -> bit-stable
        loss2 = similarity_img_native(a_img, na, b_img, nb, 256, work=work)
        torch.cuda.synchronize()
        assert float(loss2) == first
    # float64 restatement for a rectangular block (the oracle's literal form needs a square matrix): mean softplus(-(t s + b))
    want = float(torch.nn.functional.softplus(-(a.double() @ b.double().T * 10.0 - 10.0)).mean())
    if na == nb:
        assert abs(want - float(oracle.sigmoid_contrastive_loss(a.cpu().numpy(), b.cpu().numpy(), dtype=torch.float64))) < 1e-12
    assert abs(first - want) < 2e-2 * want
    # B operand partitioned over shards (what the ranks' peer-mapped buffers look like), any shard count
    if nb % 256 == 0:
        for parts in (2, 4):
            rpp = nb // parts
            if rpp % 128:
                continue
            bytes_per_part = (rpp // 128) * 4 * 16384
            ptrs = [b_img.data_ptr() + i * bytes_per_part for i in range(parts)]
            lp = similarity_img_native(a_img, na, ptrs, nb, 256, rows_per_part=rpp, work=work)
            torch.cuda.synchronize()
            assert float(lp) == first
    # row shards (what each rank computes) add up to the whole
    if na % 256 == 0:
        halves = []
        for lo in (0, na // 2):
            sub = operand_image(na // 2, 256, DEV)
            sub.copy_(a_img[(lo // 128) * 4 * 16384:((lo + na // 2) // 128) * 4 * 16384])
            halves.append(float(similarity_img_native(sub, na // 2, b_img, nb, 256, out_scale=1.0 / (na * nb))))
        torch.cuda.synchronize()
        assert abs(sum(halves) - first) < 1e-12 * abs(first)


def test_single_rank_peer_barrier_and_slot_sum():
    lib = N.lib()
    flags = torch.zeros(8, dtype=torch.int64, device=DEV)
    epoch = torch.zeros(1, dtype=torch.int64, device=DEV)
    slots = torch.tensor([1.5, 2.25, -0.75], dtype=torch.float64, device=DEV)
    out = torch.zeros((), dtype=torch.float64, device=DEV)
    for it in range(3):
        N.check(lib.cmhar_peer_barrier(N.ptr_array([flags.data_ptr()]), 0, 1, epoch.data_ptr(), slots.data_ptr(), 3, 0.5,
                                       out.data_ptr(), N.stream_ptr(torch.device(DEV))))
        torch.cuda.synchronize()
        assert int(epoch) == it + 1 and int(flags[0]) == it + 1
        assert float(out) == 1.5
    assert lib.cmhar_peer_barrier(N.ptr_array([flags.data_ptr()]), 1, 1, epoch.data_ptr(), None, 0, 1.0, None, None) == -1


def test_head_kernel_kind_is_queryable():
    """The bf16 head falls back to the fp32 CUDA-core kernel for layouts the tensor-core kernel does not serve; the
    library says which kernel a call launches instead of hiding it."""
    from crossmodal_imu_video_ood_har_b200.models import pack_head_blob
    lib = N.lib()
    xm, fus, _, _ = modules()
    ref_head = fus._head_blob(torch.device(DEV))
    assert lib.cmhar_head_kernel_kind(ref_head.data_ptr(), None, N.BF16) == 1
    assert lib.cmhar_head_kernel_kind(ref_head.data_ptr(), None, N.FP32) == 0
    odd = torch.nn.Sequential(torch.nn.Linear(128, 64), torch.nn.BatchNorm1d(64), torch.nn.ReLU(), torch.nn.Dropout(0.1),
                              torch.nn.Linear(64, 32), torch.nn.BatchNorm1d(32), torch.nn.ReLU(), torch.nn.Dropout(0.1),
                              torch.nn.Linear(32, 10)).to(DEV).eval()
    odd_blob = pack_head_blob(odd, torch.device(DEV))
    assert lib.cmhar_head_kernel_kind(odd_blob.data_ptr(), None, N.BF16) == 0
    moved = ref_head.clone()                                                # a copied blob is unknown to the registry
    assert lib.cmhar_head_kernel_kind(moved.data_ptr(), None, N.BF16) == 0
