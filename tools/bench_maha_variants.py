"""Development aid: the streaming Mahalanobis score kernel's role-layout variants (CMHAR_MAHA_VARIANT, read once per
process) -- time and check against the fp32 CUDA-core kernel.   CMHAR_MAHA_VARIANT=k python tools/bench_maha_variants.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
from oracle import weights as W
N = cm._native
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6539.9
feats, labels = W.class_features(1, 20000)
maha = cm.MahalanobisOOD(32, dev, ridge=1e-3).fit(torch.from_numpy(feats).to(dev), torch.from_numpy(labels).to(dev))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (2_000_000, 1_000_003, 65_536):
    q, _ = W.class_features(3, min(n, 200_000), ood_fraction=0.5)
    feat = torch.from_numpy(q).to(dev).repeat((n + q.shape[0] - 1) // q.shape[0], 1)[:n].contiguous()
    ref = maha.score(feat, precision="fp32")
    worst = 0.0
    for rep in range(5):                      # repeated: a hand-off race would show up as run-to-run differences
        got = maha.score(feat, precision="bf16")
        worst = max(worst, float((got - ref).abs().max() / ref.abs().max()))
    score = torch.empty(n, device=dev)
    f = lambda: N.check(N.lib().cmhar_maha_score(maha.blob(dev).data_ptr(), feat.data_ptr(), n, score.data_ptr(), 1, N.stream_ptr(dev)))
    for _ in range(3): f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = n * 516 / ms / 1e6
    print(f"variant {os.environ.get('CMHAR_MAHA_VARIANT', '0')} n={n:8d}: {ms * 1e3:8.1f} us  {gbs:7.1f} GB/s  {100 * gbs / peak:5.1f} %   max rel diff vs fp32 kernel {worst:.2e}")
