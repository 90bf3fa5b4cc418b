"""ncu target: the Mahalanobis fit kernel (2 M rows) and the pooling kernel (2 048 clips), a few launches each.
   ncu --set full -k regex:maha_fit_tc|video_pool_kernel ... python tools/profile_fit_pool.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
N = cm._native; dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
n = 2_000_000
feat = torch.randn(n, 128, device=dev, generator=g)
lab = torch.randint(0, 32, (n,), device=dev, generator=g)
cnt = torch.zeros(32, dtype=torch.float64, device=dev); ssum = torch.zeros(32, 128, dtype=torch.float64, device=dev); sec = torch.zeros(128, 128, dtype=torch.float64, device=dev)
B = 2048
fm = torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev, generator=g)).to(torch.bfloat16)
pooled = torch.empty(B, 512, device=dev)
for _ in range(3):
    N.check(N.lib().cmhar_maha_accumulate(feat.data_ptr(), lab.data_ptr(), n, 32, cnt.data_ptr(), ssum.data_ptr(), sec.data_ptr(), N.BF16, N.stream_ptr(dev)))
    N.check(N.lib().cmhar_video_pool(fm.data_ptr(), 1, B, 16, 512, 16, pooled.data_ptr(), N.stream_ptr(dev)))
torch.cuda.synchronize()
print("ok", float(cnt.sum()), float(pooled[0, 0]))
