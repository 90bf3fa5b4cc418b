"""Profiling driver: a few launches of the tensor-core conv encoder, for `ncu --set full -k regex:conv_encoder_tc`.
   python tools/profile_conv.py [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.manual_seed(0)
enc = cm.ConvIMUEncoder(cm.default_config()).to("cuda").eval()
x = torch.randn(nb, 6, 250, device="cuda")
for _ in range(3):
    enc.forward_native(x, precision="bf16")
torch.cuda.synchronize()
