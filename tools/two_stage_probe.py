"""Development aid: the pipelined step with the HBM-bound pooling taken out of the per-lane graphs and issued on
ONE shared stream (pool kernels of different steps never run concurrently), lanes wait on its event.
   [CMHAR_POOL_GRIDY=k] python tools/two_stage_probe.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
N = cm._native
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = cm.default_config()
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to(dev).eval()
xm = cm.CrossModalModel(cfg).to(dev).eval()
fus = cm.LateFusionClassifier(clf.imu_encoder, xm.video_encoder, cfg).to(dev).eval()
B = 256
L = int(os.environ.get("LANES", "8"))
STEPS = 2000
xs = [torch.randn(B, 6, 250, device=dev) for _ in range(L)]
fs = [torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev)).to(torch.bfloat16) for _ in range(L)]
pooled = [torch.empty(B, 512, device=dev) for _ in range(L)]
pipes = [cm.CrossModalOODPipeline(clf, xm, None, frames=16, precision="bf16", fusion=fus) for _ in range(L)]
graphs = []
with torch.no_grad():
    for i in range(L):
        xm.video_encoder.pool_features(fs[i], 16, out=pooled[i])
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): pipes[i].run(xs[i], None, pooled=pooled[i])
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): out = pipes[i].run(xs[i], None, pooled=pooled[i])
        graphs.append((g, out))
    pool_graphs = []
    for i in range(L):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): xm.video_encoder.pool_features(fs[i], 16, out=pooled[i])
        pool_graphs.append(g)
lanes = [torch.cuda.Stream() for _ in range(L)]
spool = torch.cuda.Stream()
ready = [torch.cuda.Event() for _ in range(L)]
done = [torch.cuda.Event() for _ in range(L)]
lib = N.lib()

def run(steps, use_graph_for_pool):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    spool.wait_event(e0)
    for ln in lanes: ln.wait_event(e0)
    with torch.no_grad():
        for i in range(steps):
            k = i % L
            if i >= L: spool.wait_event(done[k])              # pooled[k] is still being read by step i - L
            with torch.cuda.stream(spool):
                if use_graph_for_pool: pool_graphs[k].replay()
                else: xm.video_encoder.pool_features(fs[k], 16, out=pooled[k])
                ready[k].record(spool)
            with torch.cuda.stream(lanes[k]):
                lanes[k].wait_event(ready[k])
                graphs[k][0].replay()
                done[k].record(lanes[k])
    host = (time.perf_counter() - t0) / steps * 1e6
    for ln in lanes + [spool]:
        ev = torch.cuda.Event(); ev.record(ln); torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3, host

for use_graph in (False, True):
    run(200, use_graph)
    us, host = run(STEPS, use_graph)
    print(f"lanes {L} pool-gridy {os.environ.get('CMHAR_POOL_GRIDY', '0'):>3s} pool via {'graph ' if use_graph else 'direct'}: {us:6.2f} us/step   (host issue {host:5.2f} us/step)")
