"""Development aid: device time of the Mahalanobis finalisation kernel and of the pack kernels behind it.   python tools/bench_maha_finalize.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
N = cm._native; lib = N.lib(); dev = torch.device("cuda:0")
C = 32
g = torch.Generator(device=dev).manual_seed(1)
mu = 2.0 * torch.randn(C, 128, device=dev, generator=g)
y = torch.randint(0, C, (200000,), device=dev, generator=g)
f = mu[y] + torch.randn(200000, 128, device=dev, generator=g)
m = cm.MahalanobisOOD(C, dev, ridge=1e-3)
m.accumulate(f, y, precision="bf16")
stats = m._stats
fit64 = torch.empty(lib.cmhar_maha_fit64_doubles(C), dtype=torch.float64, device=dev)
w32 = torch.empty(128, 128, device=dev); mw32 = torch.empty(C, 128, device=dev); c32 = torch.empty(C, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
blob = N.alloc_blob(lib.cmhar_maha_blob_bytes(C), dev)
st = N.stream_ptr(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
fin = lambda: N.check(lib.cmhar_maha_finalize(stats.data_ptr(), C, 1e-3, fit64.data_ptr(), w32.data_ptr(), mw32.data_ptr(), c32.data_ptr(), info.data_ptr(), st))
pack = lambda: N.check(lib.cmhar_maha_pack(w32.data_ptr(), mw32.data_ptr(), c32.data_ptr(), C, blob.data_ptr(), st))
print(f"cmhar_maha_finalize: {t(fin):7.1f} us   cmhar_maha_pack: {t(pack):7.1f} us   info {int(info.item())}")
t0 = time.perf_counter()
for _ in range(20): m.finalize(all_reduce=False)
print(f"MahalanobisOOD.finalize (device route, incl. status read): {(time.perf_counter() - t0) / 20 * 1e6:7.1f} us wall")
t0 = time.perf_counter()
for _ in range(20): m.finalize(all_reduce=False, on_device=False); m.blob(dev)
torch.cuda.synchronize()
print(f"MahalanobisOOD.finalize (host route + pack): {(time.perf_counter() - t0) / 20 * 1e6:7.1f} us wall")
