"""Development aid: head + scores (+ Mahalanobis) launch time by batch size, both kernels.
   python tools/bench_head.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
from oracle import weights as W
N = cm._native
dev = torch.device("cuda:0")
lib = N.lib()
cfg = cm.default_config()
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to(dev).eval()
hb = clf._head_blob(dev)
feats, labels = W.class_features(1, 20000)
maha = cm.MahalanobisOOD(32, dev, ridge=1e-3).fit(torch.from_numpy(feats).to(dev), torch.from_numpy(labels).to(dev))
mb = maha.blob(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (256, 1024, 4096, 16384, 65536, 1 << 20):
    feat = torch.randn(n, 128, device=dev)
    lg = torch.empty(n, 32, device=dev); pr = torch.empty(n, dtype=torch.int64, device=dev)
    m1, m2, m3 = torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev)
    row = [f"n={n:8d}"]
    for prec, name in ((0, "fp32"), (1, "tc")):
        f = lambda: N.check(lib.cmhar_head_forward(hb.data_ptr(), mb.data_ptr(), feat.data_ptr(), n, lg.data_ptr(), pr.data_ptr(), m1.data_ptr(), m2.data_ptr(), m3.data_ptr(), prec, N.stream_ptr(dev)))
        for _ in range(3): f()
        torch.cuda.synchronize()
        reps = 20 if n <= 65536 else 5
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        row.append(f"{name} {us:9.1f} us ({n / us:7.2f} M rows/s)")
    print("   ".join(row))
