"""Development aid: achieved GB/s of the HBM-bound stages against the measured copy bandwidth
(algorithmic bytes per unit from SURVEY.md section 8d).   python tools/bench_hbm_kernels.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
from oracle import weights as W
import numpy as np

N = cm._native
N.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
dev = torch.device("cuda:0")
peak = 6539.9
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:44s} {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f} % of {peak:.0f}")


lib = N.lib()
st = lambda: N.stream_ptr(dev)
# video pooling: bf16 feature maps (B*16, 512, 4, 4) -> (B, 512)
for B in (256, 2048):
    sets = [torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev)).to(torch.bfloat16) for _ in range(max(2, 600_000_000 // (B * 262144)))]
    pooled = torch.empty(B, 512, device=dev)
    it = [0]
    def f():
        x = sets[it[0] % len(sets)]; it[0] += 1
        N.check(lib.cmhar_video_pool(x.data_ptr(), 1, B, 16, 512, 16, pooled.data_ptr(), st()))
    report(f"video_pool B={B} (bf16 fmap)", timeit(f), B * (16 * 512 * 16 * 2 + 512 * 4))
    del sets
# MSP / energy from stored logits
n = 8_000_000
logits = torch.randn(n, 32, device=dev)
pred = torch.empty(n, dtype=torch.int64, device=dev); msp = torch.empty(n, device=dev); en = torch.empty(n, device=dev)
f = lambda: N.check(lib.cmhar_logit_scores(logits.data_ptr(), n, 32, 1.0, pred.data_ptr(), msp.data_ptr(), en.data_ptr(), st()))
report(f"logit_scores n={n}", timeit(f), n * (128 + 16))
del logits
# head + scores / Mahalanobis from stored features
n = 2_000_000
feats, labels = W.class_features(1, 20000)
maha = cm.MahalanobisOOD(32, dev, ridge=1e-3).fit(torch.from_numpy(feats).to(dev), torch.from_numpy(labels).to(dev))
feat = torch.randn(n, 128, device=dev)
lab = torch.randint(0, 32, (n,), device=dev)
score = torch.empty(n, device=dev)
for PREC, pname in ((0, "fp32 FMA"), (1, "tcgen05 split-bf16")):
    f = lambda: N.check(lib.cmhar_maha_score(maha.blob(dev).data_ptr(), feat.data_ptr(), n, score.data_ptr(), PREC, st()))
    report(f"maha_score n={n} [{pname}]", timeit(f, 10), n * (512 + 4))
cnt = torch.zeros(32, dtype=torch.float64, device=dev); ssum = torch.zeros(32, 128, dtype=torch.float64, device=dev); sec = torch.zeros(128, 128, dtype=torch.float64, device=dev)
for PREC, pname in ((0, "fp32 FMA"), (1, "tcgen05 split-bf16")):
    f = lambda: N.check(lib.cmhar_maha_accumulate(feat.data_ptr(), lab.data_ptr(), n, 32, cnt.data_ptr(), ssum.data_ptr(), sec.data_ptr(), PREC, st()))
    report(f"maha_accumulate n={n} [{pname}]", timeit(f, 10), n * (512 + 8))
cfg = cm.default_config()
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to(dev).eval()
hb = clf._head_blob(dev)
lg = torch.empty(n, 32, device=dev); pr = torch.empty(n, dtype=torch.int64, device=dev); m1 = torch.empty(n, device=dev); m2 = torch.empty(n, device=dev); m3 = torch.empty(n, device=dev)
for PREC, pname in ((0, 'fp32 FMA'), (1, 'tcgen05 split-bf16')):
    f = lambda: N.check(lib.cmhar_head_forward(hb.data_ptr(), maha.blob(dev).data_ptr(), feat.data_ptr(), n, lg.data_ptr(), pr.data_ptr(), m1.data_ptr(), m2.data_ptr(), m3.data_ptr(), PREC, st()))
    ms = timeit(f, 5)
    report(f"head+scores+maha n={n} [" + pname + "]", ms, n * (512 + 128 + 20))
    print(f"   head: {n * (139264 + 2 * 128 * 128 + 3 * 128 * 32) / ms / 1e9:.1f} TFLOP/s fp32")
    f = lambda: N.check(lib.cmhar_head_forward(hb.data_ptr(), None, feat.data_ptr(), n, lg.data_ptr(), pr.data_ptr(), m1.data_ptr(), m2.data_ptr(), None, PREC, st()))
    ms = timeit(f, 5)
    report(f"head+scores n={n} [" + pname + "]", ms, n * (512 + 128 + 16))
    print(f"   head: {n * 139264 / ms / 1e9:.1f} TFLOP/s fp32")
# ROC histograms
n = 16_000_000
s = torch.randn(n, device=dev)
rng = torch.empty(2, dtype=torch.int32, device=dev)
init = torch.tensor([-1, 0], dtype=torch.int32, device=dev)
def f():
    rng.copy_(init)
    N.check(lib.cmhar_score_key_range(s.data_ptr(), n, rng.data_ptr(), st()))
report(f"score_key_range n={n}", timeit(f, 10), n * 4)
hist = torch.zeros(65536, dtype=torch.int64, device=dev)
f = lambda: N.check(lib.cmhar_score_histogram(s.data_ptr(), n, 0, 16, 65536, hist.data_ptr(), st()))
report(f"score_histogram n={n} (65536 bins)", timeit(f, 10), n * 4)
