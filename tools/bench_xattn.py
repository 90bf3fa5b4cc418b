"""GPU tool: the cross-attention fusion block at 4096 windows -- fused tcgen05 kernel vs the chained route (launch times), and
an ncu target.   python tools/bench_xattn.py [windows]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import crossmodal_imu_video_ood_har_b200 as cm
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
cfg, clf, xm, fus = bench.build_modules(dev)
torch.manual_seed(5)
xf = cm.CrossAttentionFusionClassifier(xm.imu_encoder, xm.video_encoder, cfg).to(dev).eval()
imu, fmap = bench.synth_inputs(dev, B, 1, 0)[0]
tokens = imu_forward_native(clf.imu_encoder, None, None, imu, want_tokens=True, precision="bf16")["tokens"]
_, frame_img = xm.video_encoder.pool_features_frames(fmap, 16, want_clip_img=False)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
a = t(lambda: xf.fuse_native_img(tokens, frame_img, 16))
b = t(lambda: xf.fuse_native_img(tokens, frame_img, 16, fused_kernel=False))
p = t(lambda: xm.video_encoder.pool_features_frames(fmap, 16))
nbytes = B * 16 * 512 * 2 + B * 16 * 128 * 4 + B * 512
print(f"{B} windows: fused cross-attention kernel {a:.1f} us ({nbytes / a / 1e3:.0f} GB/s of frame tokens + IMU tokens), chained route {b:.1f} us, x{b / a:.1f}; "
      f"single-pass pooling (clip + frame images) {p:.1f} us = {B * 16 * 512 * 16 * 2 / p / 1e3:.0f} GB/s")
