"""Development aid: per-CTA trace (kernel id, SM, start, end) of the pipelined 256-window step -- how do the CTAs of
the concurrently running kernels of 8 lanes share the SMs?   python tools/cta_trace.py [full|encpool]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native
N = cm._native
dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
torch.manual_seed(0)
cfg = cm.default_config()
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to(dev).eval()
xm = cm.CrossModalModel(cfg).to(dev).eval()
fus = cm.LateFusionClassifier(clf.imu_encoder, xm.video_encoder, cfg).to(dev).eval()
B, L, STEPS = 256, int(os.environ.get("LANES", "16")), 400
xs = [torch.randn(B, 6, 250, device=dev) for _ in range(L)]
fs = [torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev)).to(torch.bfloat16) for _ in range(L)]
pooled = [torch.empty(B, 512, device=dev) for _ in range(L)]
pipes = [cm.CrossModalOODPipeline(clf, xm, None, frames=16, precision="bf16", fusion=fus) for _ in range(L)]
side = torch.cuda.Stream()

def encpool(i):
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        N.check(N.lib().cmhar_video_pool(fs[i].data_ptr(), 1, B, 16, 512, 16, pooled[i].data_ptr(), N.stream_ptr(dev)))
    o = imu_forward_native(clf.imu_encoder, None, None, xs[i], want_cls=True, precision="bf16")
    main.wait_stream(side)
    return o

fn = (lambda i: pipes[i].run(xs[i], fs[i])) if mode == "full" else encpool
graphs = []
with torch.no_grad():
    for i in range(L):
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): fn(i)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): out = fn(i)
        graphs.append((g, out))
lanes = [torch.cuda.Stream() for _ in range(L)]
CAP = 2_000_000
buf = torch.zeros(2 + 4 * CAP, dtype=torch.int64, device=dev)

def run(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for ln in lanes: ln.wait_event(e0)
    for i in range(steps):
        with torch.cuda.stream(lanes[i % L]): graphs[i % L][0].replay()
    for ln in lanes:
        ev = torch.cuda.Event(); ev.record(ln); torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

run(200)
print(f"{mode}: untraced {run(STEPS):.2f} us/step")
N.check(N.lib().cmhar_debug_cta_trace(buf.data_ptr(), CAP))
us = run(STEPS)
N.check(N.lib().cmhar_debug_cta_trace(None, 0))
h = buf.cpu().numpy()
n = int(min(h[0], CAP)); rec = h[2:2 + 4 * n].reshape(n, 4)
print(f"{mode}: traced {us:.2f} us/step, {n} CTA records over {STEPS} steps")
os.makedirs("gpurun_out", exist_ok=True); np.save(f"gpurun_out/cta_trace_{mode}.npy", rec)
t0, t1 = rec[:, 2].min(), rec[:, 3].max()
span = (t1 - t0) / 1e3
names = {1: "encoder", 2: "pool", 3: "head_tc", 4: "linear_tc", 5: "similarity"}
print(f"span {span:.0f} us = {span / STEPS:.2f} us/step; SM-time shares (148 SMs x span = 100 %):")
for k, nm in names.items():
    r = rec[rec[:, 0] == k]
    if len(r) == 0: continue
    d = (r[:, 3] - r[:, 2]) / 1e3
    print(f"  {nm:10s} {len(r) / STEPS:7.1f} CTAs/step  duration mean {d.mean():7.1f} us  p10 {np.percentile(d, 10):7.1f}  p90 {np.percentile(d, 90):7.1f}  "
          f"sum {d.sum() / STEPS:8.0f} SM-us/step = {100 * d.sum() / (148 * span):5.1f} % of SM-time")
# per-SM: fraction of the span during which an encoder CTA is resident; idle gaps between consecutive encoder CTAs
enc = rec[rec[:, 0] == 1]
mid = (enc[:, 2] > t0 + 0.2 * (t1 - t0)) & (enc[:, 3] < t0 + 0.8 * (t1 - t0))
gaps = []
for sm in range(148):
    e = enc[(enc[:, 1] == sm) & mid]
    e = e[np.argsort(e[:, 2])]
    if len(e) > 1: gaps += list((e[1:, 2] - e[:-1, 3]) / 1e3)
gaps = np.array(gaps)
if len(gaps):
    print(f"encoder CTAs per SM back to back: gap mean {gaps.mean():.1f} us  median {np.median(gaps):.1f}  p90 {np.percentile(gaps, 90):.1f}  (negative = overlap impossible)")
# what runs on an SM during the encoder gaps?
pool = rec[rec[:, 0] == 2]
if len(pool):
    # pool CTAs concurrent with an encoder CTA on the same SM
    co = 0
    for sm in range(0, 148, 8):
        e = enc[enc[:, 1] == sm]; p = pool[pool[:, 1] == sm]
        for ps, pe in p[:, 2:4][:2000]:
            co += bool(np.any((e[:, 2] < pe) & (e[:, 3] > ps)))
    tot = sum(min(2000, int((pool[:, 1] == sm).sum())) for sm in range(0, 148, 8))
    print(f"pool CTAs that overlapped an encoder CTA on their SM: {100 * co / max(tot, 1):.0f} % (sampled)")
    # time-resolved: number of SMs hosting an encoder CTA, sampled every 5 us over the middle of the run
    ts = np.arange(t0 + 0.3 * (t1 - t0), t0 + 0.7 * (t1 - t0), 5000)
    occ = np.array([((enc[:, 2] <= t) & (enc[:, 3] > t)).sum() for t in ts])
    pocc = np.array([((pool[:, 2] <= t) & (pool[:, 3] > t)).sum() for t in ts])
    print(f"SMs hosting an encoder CTA: mean {occ.mean():.1f} of 148 (min {occ.min()}, max {occ.max()}); resident pool CTAs: mean {pocc.mean():.0f} (max {pocc.max()})")
