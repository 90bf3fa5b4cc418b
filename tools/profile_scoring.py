"""ncu target: the standalone scoring kernels at bandwidth-saturating sizes (one launch each after warm-up).
   ncu --set full -k regex:maha_score_tc|logit_scores_ring ... python tools/profile_scoring.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
N = cm._native; dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(7)
n_rows, n_feat = 4_000_000, 2_000_000
logits = torch.randn(n_rows, 32, device=dev, generator=g)
pr = torch.empty(n_rows, dtype=torch.int64, device=dev); s1 = torch.empty(n_rows, device=dev); s2 = torch.empty(n_rows, device=dev)
mu = 2.0 * torch.randn(32, 128, device=dev, generator=g)
y = torch.randint(0, 32, (20000,), device=dev, generator=g)
maha = cm.MahalanobisOOD(32, dev, ridge=1e-3).fit(mu[y] + torch.randn(20000, 128, device=dev, generator=g), y)
feat = mu[torch.randint(0, 32, (n_feat,), device=dev, generator=g)] + torch.randn(n_feat, 128, device=dev, generator=g)
sc = torch.empty(n_feat, device=dev)
for _ in range(3):
    N.check(N.lib().cmhar_logit_scores(logits.data_ptr(), n_rows, 32, 1.0, pr.data_ptr(), s1.data_ptr(), s2.data_ptr(), N.stream_ptr(dev)))
    N.check(N.lib().cmhar_maha_score(maha.blob(dev).data_ptr(), feat.data_ptr(), n_feat, sc.data_ptr(), N.BF16, N.stream_ptr(dev)))
torch.cuda.synchronize()
print("ok", float(sc[:4].sum()), float(s1[:4].sum()))
