"""Profiling driver (test infrastructure): a few launches of the fused IMU kernel at a given batch,
for `ncu --set full -k regex:imu_forward`.   python tools/profile_imu.py [batch] [precision] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
nohead = len(sys.argv) > 4 and sys.argv[4] == "nohead"      # encoder only (CLS out), no classifier head / scores
if os.environ.get("ENC_KERNEL"):            # 1 = single-tile kernel, 2 = two-tiles-in-flight kernel (development switch of the library)
    cm._native.check(cm._native.lib().cmhar_debug_set_option(b"enc_kernel", int(os.environ["ENC_KERNEL"])))
torch.manual_seed(0)
cfg = cm.default_config()
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to("cuda").eval()
x = torch.randn(batch, 6, 250, device="cuda")
out = {}
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run = (lambda: clf.imu_encoder.encode_cls(x, precision=prec)) if nohead else (lambda: clf.forward_scores(x, precision=prec, out=out))
run()
torch.cuda.synchronize()
ev0.record()
for i in range(reps):
    run()
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
print(f"batch {batch} {prec}{" nohead" if nohead else ""} ablate={os.environ.get("CMHAR_ABLATE", "0")}: {ms:.3f} ms/launch -> {batch / ms * 1e3:.0f} windows/s, {25890816 * batch / ms / 1e9:.1f} TFLOP/s")
