"""Development aid: does the HBM-bound pooling kernel become resident next to the encoder's whole-SM CTAs?
Launches a long encoder kernel (65 536 windows, ~3 ms, one persistent CTA per SM) on one stream and the pooling
kernel of a 256-clip batch on another, and times the pooling kernel from its own stream's point of view.
   python tools/coresidency_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native
N = cm._native
dev = torch.device("cuda:0")
cfg = cm.default_config()
torch.manual_seed(0)
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to(dev).eval()
x = torch.randn(65536, 6, 250, device=dev)
B = 256
fm = torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev)).to(torch.bfloat16)
pooled = torch.empty(B, 512, device=dev)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

def pool():
    N.check(N.lib().cmhar_video_pool(fm.data_ptr(), 1, B, 16, 512, 16, pooled.data_ptr(), N.stream_ptr(dev)))

with torch.no_grad():
    for _ in range(3):
        imu_forward_native(clf.imu_encoder, None, None, x, want_cls=True, precision="bf16"); pool()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); pool(); e[1].record(); torch.cuda.synchronize()
    print(f"pool alone                         {e[0].elapsed_time(e[1]) * 1e3:8.1f} us")
    for rep in range(3):
        start = torch.cuda.Event(enable_timing=True); start.record()
        sa.wait_event(start); sb.wait_event(start)
        with torch.cuda.stream(sa):
            imu_forward_native(clf.imu_encoder, None, None, x, want_cls=True, precision="bf16")
            e[2].record(sa)
        with torch.cuda.stream(sb):
            torch.cuda._sleep(200_000)          # let the encoder CTAs take their SMs first (~0.1 ms)
            s0 = torch.cuda.Event(enable_timing=True); s0.record(sb)
            pool()
            e[3].record(sb)
        torch.cuda.synchronize()
        print(f"encoder {start.elapsed_time(e[2]) * 1e3:8.1f} us | pool launched at {start.elapsed_time(s0) * 1e3:8.1f} us, "
              f"took {s0.elapsed_time(e[3]) * 1e3:8.1f} us (co-resident if << encoder time)")
