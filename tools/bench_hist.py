"""Development aid: score_histogram_kernel on concentrated (MSP-like) and spread scores.   python tools/bench_hist.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
dev = torch.device("cuda:0"); n = 16_000_000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, sc in (("msp-like (-softmax max, 32 classes)", -torch.softmax(torch.randn(n, 32, device=dev), 1).max(1)[0]),
                 ("spread (N(0,1) * 50)", torch.randn(n, device=dev) * 50)):
    a, b = sc[: n // 2].contiguous(), sc[n // 2:].contiguous()
    cm.auroc_fpr95(a, b)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): r = cm.auroc_fpr95(a, b)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: auroc_fpr95 of {n} scores {e0.elapsed_time(e1) / 3:.2f} ms  auroc {r['auroc']:.4f}")
