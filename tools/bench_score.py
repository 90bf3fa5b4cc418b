"""Development aid: maha_score_tc_kernel alone at 2 M / 1 M rows (CMHAR_MAHA_VARIANT 0 | 3, CMHAR_L2_PREFETCH)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
from oracle import weights as W
N = cm._native; N.enable_dev_env(); lib = N.lib(); dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
feats, labels = W.class_features(1, 20000)
maha = cm.MahalanobisOOD(32, dev, ridge=1e-3).fit(torch.from_numpy(feats).to(dev), torch.from_numpy(labels).to(dev))
for n in (2_000_000, 1_000_000):
    feat = torch.randn(n, 128, device=dev); score = torch.empty(n, device=dev)
    f = lambda: N.check(lib.cmhar_maha_score(maha.blob(dev).data_ptr(), feat.data_ptr(), n, score.data_ptr(), 1, N.stream_ptr(dev)))
    for _ in range(3): f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"VARIANT={os.environ.get('CMHAR_MAHA_VARIANT','dflt')} PF={os.environ.get('CMHAR_L2_PREFETCH','dflt')} n={n}: {ms*1e3:7.1f} us  {n*516/ms/1e6:7.1f} GB/s  {n*516/ms/1e6/65.399:5.1f} %")
    del feat
