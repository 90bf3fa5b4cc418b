import sys, os, torch
sys.path.insert(0, os.getcwd())
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native
dev = torch.device("cuda:0")
cfg = cm.default_config(); torch.manual_seed(0)
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to(dev).eval()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for n in (256, 4096, 65536):
        x = torch.randn(n, 6, 250, device=dev)
        for _ in range(3): imu_forward_native(clf.imu_encoder, None, None, x, want_cls=True, precision="bf16")
        torch.cuda.synchronize(); reps = 50
        e0.record()
        for _ in range(reps): imu_forward_native(clf.imu_encoder, None, None, x, want_cls=True, precision="bf16")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"encoder n={n}: {ms*1e3:.1f} us  {n/ms/1e3:.2f} M windows/s  {n*25751552/ms/1e9:.0f} TFLOP/s")
