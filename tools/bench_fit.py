"""Development aid: maha_fit_tc_kernel alone at 2 M / 1 M rows (CMHAR_L2_PREFETCH = tiles ahead, CMHAR_L2_PF_LANES)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
N = cm._native; N.enable_dev_env(); lib = N.lib(); dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (2_000_000, 1_000_000):
    feat = torch.randn(n, 128, device=dev); lab = torch.randint(0, 32, (n,), device=dev)
    cnt = torch.zeros(32, dtype=torch.float64, device=dev); ssum = torch.zeros(32, 128, dtype=torch.float64, device=dev); sec = torch.zeros(128, 128, dtype=torch.float64, device=dev)
    f = lambda: N.check(lib.cmhar_maha_accumulate(feat.data_ptr(), lab.data_ptr(), n, 32, cnt.data_ptr(), ssum.data_ptr(), sec.data_ptr(), 1, N.stream_ptr(dev)))
    for _ in range(3): f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"PF={os.environ.get('CMHAR_L2_PREFETCH','dflt')} LANES={os.environ.get('CMHAR_L2_PF_LANES','dflt')} n={n}: {ms*1e3:7.1f} us  {n*520/ms/1e6:7.1f} GB/s  {n*520/ms/1e6/65.399:5.1f} %")
    del feat
