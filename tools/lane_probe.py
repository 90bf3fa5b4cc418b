"""Development aid: where does the pipelined 256-window step saturate?  Replays CUDA graphs of sub-sets of the
step over 8 streams and prints us/step.   python tools/lane_probe.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native, l2_normalize_native
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = cm.default_config()
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to(dev).eval()
xm = cm.CrossModalModel(cfg).to(dev).eval()
B, L = 256, int(os.environ.get("LANES", "8"))
xs = [torch.randn(B, 6, 250, device=dev) for _ in range(L)]
fs = [torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev)).to(torch.bfloat16) for _ in range(L)]

def capture(fn, hi_prio=False):
    graphs = []
    for i in range(L):
        s = torch.cuda.Stream(priority=-1) if hi_prio else torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(i)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s if hi_prio else None):
            out = fn(i)
        graphs.append((g, out))
    return graphs

def bench(name, fn, steps=2000, hi_prio=False):
    graphs = capture(fn, hi_prio)
    lanes = [torch.cuda.Stream(priority=-1) if hi_prio else torch.cuda.Stream() for _ in range(L)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        for ln in lanes: ln.wait_event(e0)
        for i in range(steps):
            with torch.cuda.stream(lanes[i % L]):
                graphs[i % L][0].replay()
        evs = []
        for ln in lanes:
            ev = torch.cuda.Event(); ev.record(ln); torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / steps * 1e3:7.2f} us/step")

with torch.no_grad():
    bench("encoder only (1 launch)", lambda i: imu_forward_native(clf.imu_encoder, None, None, xs[i], want_cls=True, precision="bf16"))
    bench("encoder + head_tc (2 launches)", lambda i: clf.forward_scores(xs[i], precision="bf16"))
    bench("video pool only (1 launch)", lambda i: xm.video_encoder.forward_frame_features(fs[i][:16], precision="bf16") if False else cm._native.check(cm._native.lib().cmhar_video_pool(fs[i].data_ptr(), 1, B, 16, 512, 16, torch.empty(B, 512, device=dev).data_ptr(), cm._native.stream_ptr(dev))))
    bench("video tail (pool + projection, 2 launches)", lambda i: xm.video_encoder.forward_features(fs[i], 16, precision="bf16"))
    N = cm._native
    pooled = [torch.empty(B, 512, device=dev) for _ in range(L)]
    side = torch.cuda.Stream()
    def enc_and_pool(i):
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            N.check(N.lib().cmhar_video_pool(fs[i].data_ptr(), 1, B, 16, 512, 16, pooled[i].data_ptr(), N.stream_ptr(dev)))
        o = imu_forward_native(clf.imu_encoder, None, None, xs[i], want_cls=True, precision="bf16")
        main.wait_stream(side)
        return o
    bench("encoder || pool (2 branches)", enc_and_pool)
    lo_prio = torch.cuda.Stream(priority=0)
    hi_lanes = None
    def enc_and_pool_prio(i):      # pooling on a low-priority side stream: pending encoder CTAs are dispatched first
        main = torch.cuda.current_stream()
        lo_prio.wait_stream(main)
        with torch.cuda.stream(lo_prio):
            N.check(N.lib().cmhar_video_pool(fs[i].data_ptr(), 1, B, 16, 512, 16, pooled[i].data_ptr(), N.stream_ptr(dev)))
        o = imu_forward_native(clf.imu_encoder, None, None, xs[i], want_cls=True, precision="bf16")
        main.wait_stream(lo_prio)
        return o
    bench("encoder (hi prio lanes) || pool (lo prio)", enc_and_pool_prio, hi_prio=True)
    def pool_then_enc(i):
        N.check(N.lib().cmhar_video_pool(fs[i].data_ptr(), 1, B, 16, 512, 16, pooled[i].data_ptr(), N.stream_ptr(dev)))
        return imu_forward_native(clf.imu_encoder, None, None, xs[i], want_cls=True, precision="bf16")
    bench("pool -> encoder (one branch)", pool_then_enc)
    if os.environ.get("LANE_PROBE_SHORT"):
        fus = cm.LateFusionClassifier(clf.imu_encoder, xm.video_encoder, cfg).to(dev).eval()
        pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=16, precision="bf16", fusion=fus)
        bench("full pipeline (14 launches)", lambda i: pipe.run(xs[i], fs[i]))
        sys.exit(0)
    fus = cm.LateFusionClassifier(clf.imu_encoder, xm.video_encoder, cfg).to(dev).eval()
    pipe = cm.CrossModalOODPipeline(clf, xm, None, frames=16, precision="bf16", fusion=fus)
    bench("full pipeline (14 launches)", lambda i: pipe.run(xs[i], fs[i]))
    # all L steps as parallel branches of ONE graph
    def mega():
        pipes = [cm.CrossModalOODPipeline(clf, xm, None, frames=16, precision="bf16", fusion=fus) for _ in range(L)]
        streams = [torch.cuda.Stream() for _ in range(L)]
        def run_all():
            main = torch.cuda.current_stream()
            outs = []
            for i in range(L):
                streams[i].wait_stream(main)
                with torch.cuda.stream(streams[i]):
                    outs.append(pipes[i].run(xs[i], fs[i]))
            for st in streams:
                main.wait_stream(st)
            return outs
        s0 = torch.cuda.Stream(); s0.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s0):
            run_all()
        torch.cuda.current_stream().wait_stream(s0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = run_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            e0.record()
            for _ in range(250):
                g.replay()
            e1.record(); torch.cuda.synchronize()
        print(f"{'8 steps as branches of one graph':44s} {e0.elapsed_time(e1) / (250 * L) * 1e3:7.2f} us/step")
    mega()
    cls = torch.randn(B, 128, device=dev); vf = torch.randn(B, 768, device=dev)
    def tail_only(i):
        ip = l2_normalize_native(xm.imu_proj.forward_native(cls, "bf16"))
        vp = l2_normalize_native(xm.video_proj.forward_native(vf, "bf16"))
        o = fus.forward_scores(None, None, 16, precision="bf16", imu_cls=cls, video_feat=vf)
        from crossmodal_imu_video_ood_har_b200.losses import similarity_native
        r = similarity_native(ip, vp, sigmoid=(10.0, -10.0), precision="bf16")
        return o, r
    bench("everything after the encoders (11 launches)", tail_only)
    y = torch.randn(B, 768, device=dev)
    bench("video_proj head + l2norm (3 launches)", lambda i: l2_normalize_native(xm.video_proj.forward_native(y, "bf16")))
    bench("9 x l2norm (9 tiny launches)", lambda i: [l2_normalize_native(y) for _ in range(9)][-1])
