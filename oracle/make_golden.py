"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in THIS container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz

It imports the reference modules from /root/reference (read-only, nothing is copied), loads the
deterministic parameters of ``oracle/weights.py`` into them with ``strict=True`` and records
their CPU fp32 outputs.  The fixtures hold only seeds + outputs; the GPU-side tests rebuild the
same parameters and inputs from the seeds.

Two environment work-arounds, neither of which touches reference code:
  * ``matplotlib`` / ``seaborn`` are not installed, and ``src/eval/evaluator.py:13-14`` imports
    them at module scope for its plotting helpers; empty stand-in modules are registered in
    ``sys.modules`` so that ``Evaluator.predict`` / ``compute_metrics`` can be imported.
  * the third-party video trunk (torchvision resnet18, ``src/models/models.py:163-167``) is out
    of scope, so ``video_encoder.backbone`` is swapped for ``nn.Identity()`` on the constructed
    reference object and the synthetic feature map is fed as the "video"; everything after the
    trunk (``models.py:210-216``) runs as written.
"""
from __future__ import annotations

import copy
import os
import sys
import tempfile
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference not present; goldens can only be generated in the build container")
    os.chdir(tempfile.mkdtemp(prefix="ref_cwd_"))       # configs/config.py:33-46 creates ./outputs
    sys.path.insert(0, REF)
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torch  # noqa
    from configs.config import CONFIG
    from src.models import models as ref_models
    from src.models import losses as ref_losses
    from src.eval import evaluator as ref_eval
    return torch, CONFIG, ref_models, ref_losses, ref_eval


def _to_torch_sd(torch, sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def video_encoder_with_trunk(torch, CONFIG, M, backbone="resnet18"):
    """a4 WITH the reference's CNN trunks (resnet18: src/models/models.py:163-167, mobilenet_v2: :169-173; tail :208-216), random init: the module is constructed under
    torch.manual_seed(seed_init) -- the tests construct theirs the same way, so no 45 MB state dict has to be stored -- and its
    BatchNorm statistics are randomised from a numpy stream.  Records the output and the per-frame spatial means of the trunk's map."""
    cfg = copy.deepcopy(CONFIG)
    cfg.model.video_backbone = backbone
    cfg.model.video_pretrained = False
    seed_init, seed_bn, seed_x, B, T, H = 1234, 71, 72, 2, 4, 64
    torch.manual_seed(seed_init)
    ve = M.VideoEncoder(cfg).eval()
    rs = np.random.RandomState(seed_bn)
    for m in ve.backbone.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.from_numpy((0.2 * rs.standard_normal(m.num_features)).astype(np.float32)))
            m.running_var.copy_(torch.from_numpy((0.5 + rs.rand(m.num_features)).astype(np.float32)))
            m.weight.data.copy_(torch.from_numpy((1.0 + 0.2 * rs.standard_normal(m.num_features)).astype(np.float32)))
            m.bias.data.copy_(torch.from_numpy((0.1 * rs.standard_normal(m.num_features)).astype(np.float32)))
    x = torch.from_numpy(np.random.RandomState(seed_x).standard_normal((B, T, 3, H, H)).astype(np.float32))
    with torch.no_grad():
        out = ve(x)
        fmap = ve.backbone(x.view(B * T, 3, H, H))
    np.savez_compressed(os.path.join(OUT, f"video_encoder_{backbone}.npz"), seed_init=seed_init, seed_bn=seed_bn, seed_x=seed_x, B=B, T=T, H=H,
                        out=out.numpy(), frame_means=fmap.mean(dim=(2, 3)).numpy(), fmap_shape=np.array(fmap.shape))
    print(f"video_encoder_{backbone}: out absmax {out.abs().max():.4f}, map {tuple(fmap.shape)}")


def main():
    sys.path.insert(0, REPO)
    from oracle import weights as W
    torch, CONFIG, M, LS, EV = _import_reference()
    if "--only-trunk" in sys.argv:                          # add this fixture without rewriting the others
        os.makedirs(OUT, exist_ok=True)
        for bb in ("resnet18", "mobilenet_v2"):
            video_encoder_with_trunk(torch, CONFIG, M, bb)
        return
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- a1-a3, a8: IMU classifier
    for L, B, seed_w, seed_x in ((250, 64, 11, 21), (100, 64, 12, 22), (250, 777, 13, 23)):
        cfg = copy.deepcopy(CONFIG)
        cfg.data.imu_window_size = L
        dims = W.Dims(imu_window=L)
        sd = W.classifier_state(seed_w, dims)
        enc = M.IMUEncoder(cfg)
        clf = M.IMUClassifier(enc, cfg).eval()
        clf.load_state_dict(_to_torch_sd(torch, sd), strict=True)
        x = torch.from_numpy(W.imu_windows(seed_x, B, dims))
        with torch.no_grad():
            cls, tokens = clf.imu_encoder(x)
            logits = clf(x)
            pe = clf.imu_encoder.patch_embed(x[:4])
        # dead-input property: channels 1..5 and the tail samples do not matter
        x2 = x.clone()
        x2[:, 1:] = 123.0
        x2[:, 0, 16 * dims.num_patches:] = -77.0
        with torch.no_grad():
            dead_delta = float((clf(x2) - logits).abs().max())
        np.savez_compressed(
            os.path.join(OUT, f"imu_classifier_L{L}_B{B}.npz"),
            L=L, B=B, seed_w=seed_w, seed_x=seed_x,
            logits=logits.numpy(), cls=cls.numpy(), tokens_first4=tokens[:4].numpy(),
            patch_embed_first4=pe.numpy(), preds=logits.max(1)[1].numpy().astype(np.int64),
            dead_input_delta=dead_delta)
        print(f"imu_classifier L={L} B={B}: logits absmax {logits.abs().max():.3f} dead_delta {dead_delta}")

    # ---------------------------------------------------------------- a4-a7: cross-modal + losses
    cfg = copy.deepcopy(CONFIG)
    cfg.model.video_backbone = "resnet18"
    cfg.model.video_pretrained = False
    dims = W.Dims()
    B, T, seed_w, seed_x, seed_v = 48, 16, 31, 41, 51
    sd = W.cross_modal_state(seed_w, dims)
    model = M.CrossModalModel(cfg)
    model.video_encoder.backbone = torch.nn.Identity()            # trunk out of scope (see header)
    missing = model.load_state_dict(_to_torch_sd(torch, sd), strict=True)
    model.eval()
    imu = torch.from_numpy(W.imu_windows(seed_x, B, dims))
    fmap = torch.from_numpy(W.video_feature_maps(seed_v, B, T, dims))          # (B*T, 512, 4, 4)
    video = fmap.view(B, T, dims.video_feature_dim, 4, 4)
    with torch.no_grad():
        vfeat = model.video_encoder(video)
        ip, vp = model(imu, video)
        sig = LS.SigmoidContrastiveLoss()(ip, vp)
        nce = LS.InfoNCELoss(temperature=0.07)(ip, vp)
        sim = ip @ vp.T
        imu_feat, _ = model.imu_encoder(imu)
        head_only = model.imu_proj(imu_feat)
    np.savez_compressed(
        os.path.join(OUT, "cross_modal_B48.npz"),
        B=B, T=T, seed_w=seed_w, seed_x=seed_x, seed_v=seed_v,
        video_feat=vfeat.numpy(), imu_proj=ip.numpy(), video_proj=vp.numpy(),
        imu_proj_unnormalized=head_only.numpy(),
        similarity=sim.numpy(), sigmoid_loss=float(sig), info_nce_loss=float(nce))
    print(f"cross_modal: sigmoid {float(sig):.6f} infonce {float(nce):.6f}")

    # VideoMAE branch after the HF trunk = Linear on the CLS token (models.py:201-205): pinned via
    # the same projection weights with nn.Linear directly.
    rs = np.random.RandomState(61)
    cls_tok = rs.standard_normal((8, dims.video_feature_dim)).astype(np.float32)
    with torch.no_grad():
        vm = model.video_encoder.projection(torch.from_numpy(cls_tok)).numpy()
    np.savez_compressed(os.path.join(OUT, "videomae_projection.npz"), seed=61, out=vm, seed_w=seed_w)

    # ---------------------------------------------------------------- a8-a9: Evaluator
    cfg = copy.deepcopy(CONFIG)
    dims = W.Dims()
    sd = W.classifier_state(14, dims)
    clf = M.IMUClassifier(M.IMUEncoder(cfg), cfg)
    clf.load_state_dict(_to_torch_sd(torch, sd), strict=True)
    n, bs = 200, 64
    x = torch.from_numpy(W.imu_windows(24, n, dims))
    labels = torch.from_numpy(np.random.RandomState(25).randint(0, 32, size=n).astype(np.int64))
    loader = [{"imu": x[i:i + bs], "label": labels[i:i + bs]} for i in range(0, n, bs)]
    EV.tqdm = lambda it, **kw: it
    ev = EV.Evaluator(clf, cfg, device="cpu")
    res = ev.evaluate(loader)
    # metrics on a second, less trivial label pair (predictions correlated with labels)
    rs = np.random.RandomState(26)
    yt = rs.randint(0, 32, size=5000)
    yp = np.where(rs.rand(5000) < 0.7, yt, rs.randint(0, 30, size=5000))
    m2 = ev.compute_metrics(yt, yp)
    np.savez_compressed(
        os.path.join(OUT, "evaluator_n200.npz"),
        n=n, bs=bs, seed_w=14, seed_x=24, seed_y=25,
        preds=res["predictions"].astype(np.int64), labels=res["labels"].astype(np.int64),
        logits=res["logits"],
        metric_names=np.array(sorted(res["metrics"])),
        metrics=np.array([res["metrics"][k] for k in sorted(res["metrics"])], dtype=np.float64),
        m2_seed=26, m2=np.array([m2[k] for k in sorted(m2)], dtype=np.float64))
    print("evaluator:", res["metrics"])
    for bb in ("resnet18", "mobilenet_v2"):
        video_encoder_with_trunk(torch, CONFIG, M, bb)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
