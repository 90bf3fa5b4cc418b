"""Profiling driver: the two kernels either side of the device video trunk, for `ncu --set full -k regex:"frames_normalize|video_pool_nhwc"`.
   python tools/profile_frames.py [clips]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T, H = 16, 112
cfg = cm.default_config()
cfg.model.video_backbone, cfg.model.video_pretrained = "resnet18", False
torch.manual_seed(0)
ve = cm.VideoEncoder(cfg).to("cuda").eval()
trunk = cm.DeviceVideoTrunk(ve, graphs=False).to("cuda")
u8 = torch.randint(0, 256, (B, T, H, H, 3), dtype=torch.uint8, device="cuda")
x = trunk._slot(B * T, H, H, 0)["x"]
fmap = torch.relu(torch.randn(B * T, 512, 4, 4, device="cuda")).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
for _ in range(3):
    trunk.normalize_into(u8.view(B * T, H, H, 3), x)
    ve.pool_features(fmap, T, want_img=True, want_rows=False)
torch.cuda.synchronize()
