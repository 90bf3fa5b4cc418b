"""Development aid: the device video trunk -- normalise kernel, channels-last bf16 resnet18 under a CUDA graph, NHWC
pooling kernel -- each alone and as the frames-in pipeline through stream_host.   python tools/bench_trunk.py [clips]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T, H = 16, 112
cfg = cm.default_config()
cfg.model.video_backbone, cfg.model.video_pretrained = "resnet18", False
torch.manual_seed(0)
ve = cm.VideoEncoder(cfg).to(dev).eval()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


u8 = torch.randint(0, 256, (B, T, H, H, 3), dtype=torch.uint8, device=dev)
FLOP = 2 * 0.5607e9 / 4 * 1.0      # resnet18 trunk at 112x112: ~0.28 GMAC per frame (1/4 of the 224x224 figure of 1.8 GFLOP less the fc)
for cpad, fused in ((8, True), (4, True), (3, True), (8, False), (3, False)):
    trunk = cm.DeviceVideoTrunk(ve, pad_in_channels=cpad, fused_epilogues=fused).to(dev)
    fmap = trunk(u8)
    ms = timeit(lambda: trunk(u8))
    print(f"trunk graph cpad={cpad} fused={trunk.fused} ({trunk.fused_error}): {ms:8.3f} ms per {B} clips = {B / ms * 1e3:9.0f} clips/s   map {tuple(fmap.shape)} {fmap.dtype} channels_last={fmap.is_contiguous(memory_format=torch.channels_last)}")
    if not fused:
        continue
    x = trunk._slot(B * T, H, H, 0)["x"]
    ms = timeit(lambda: trunk.normalize_into(u8.view(B * T, H, H, 3), x), 20)
    nb = u8.numel() + x.numel() * 2
    print(f"   frames_normalize cpad={cpad}: {ms * 1e3:8.1f} us  {nb / ms / 1e6:8.1f} GB/s")
with torch.no_grad():
    x32 = trunk.reference_normalize(u8[:32].reshape(32 * T, H, H, 3))
    ms = timeit(lambda: ve.backbone(x32), 5)
    print(f"eager fp32 NCHW trunk (the reference's call): {ms:8.3f} ms per 32 clips = {32 / ms * 1e3:9.0f} clips/s")
trunk = cm.DeviceVideoTrunk(ve).to(dev)
fmap = trunk(u8)
for name, f in (("nhwc", lambda: ve.pool_features(fmap, T, want_img=True, want_rows=False)),
                ("nchw", None)):
    if f is None:
        fm2 = fmap.contiguous()
        f = lambda: ve.pool_features(fm2, T, want_img=True, want_rows=False)
    ms = timeit(f, 20)
    print(f"pool {name}: {ms * 1e3:8.1f} us  {fmap.numel() * 2 / ms / 1e6:8.1f} GB/s")
# frames-in pipeline through stream_host (host uint8 frames, pinned)
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
xm = cm.CrossModalModel(cfg)
fus = cm.LateFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg)
xm, clf, fus = xm.to(dev).eval(), clf.to(dev).eval(), fus.to(dev).eval()
pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=T, precision="bf16", fusion=fus)
pipe.attach_trunk(True)
imu = torch.randn(B, 6, 250).pin_memory()
host = [torch.randint(0, 256, (B, T, H, H, 3), dtype=torch.uint8).pin_memory() for _ in range(3)]
for _ in pipe.stream_host([(imu, host[i % 3]) for i in range(4)]):
    pass
torch.cuda.synchronize()
K = 12
t0 = time.perf_counter()
n = 0
for r in pipe.stream_host([(imu, host[i % 3]) for i in range(K)]):
    n += int(r["pred"].numel())
torch.cuda.synchronize()
dt = time.perf_counter() - t0
h2d, d2h = pipe.host_bytes_per_step(B, 250, host[0])
print(json.dumps({"e2e_from_frames_windows_per_s": n / dt, "ms_per_step": dt / K * 1e3, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "h2d_gbs": h2d * K / dt / 1e9}))
