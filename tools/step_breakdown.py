"""GPU tool (development): CUDA-event time of every launch of the 7-launch bf16 step at the benchmark's batch, each kernel
timed ALONE (back to back on one stream, rotating input sets), next to the single-stream latency of the whole graph.
    python tools/step_breakdown.py [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import crossmodal_imu_video_ood_har_b200 as cm  # noqa: E402
from crossmodal_imu_video_ood_har_b200.losses import similarity_img_native, similarity_img_work  # noqa: E402
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native  # noqa: E402


def timeit(fn, n=200, warm=10):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda:0")
    cfg, clf, xm, fus = bench.build_modules(dev)
    sets = bench.synth_inputs(dev, B, 8, 0)
    T = bench.FRAMES
    feats = torch.randn(4096, 128, device=dev)
    maha = cm.MahalanobisOOD(32, dev, ridge=1e-3).fit(feats, torch.randint(0, 32, (4096,), device=dev), all_reduce=False)
    fus.set_mahalanobis(maha)
    ve = xm.video_encoder
    outs = [dict() for _ in sets]
    rows = {}
    rows["encoder (cls + cls_img)"] = timeit(lambda i: imu_forward_native(clf.imu_encoder, None, None, sets[i % 8][0], want_cls=True, precision="bf16",
                                                                           want_cls_img=True, out=outs[i % 8]))
    enc = outs[0]
    rows["imu projection head (mlp2, K=128)"] = timeit(lambda i: xm.imu_proj.forward_fused(enc["cls_img"], B))
    rows["video pool -> image"] = timeit(lambda i: ve.pool_features(sets[i % 8][1], T, want_img=True, want_rows=False))
    _, pimg = ve.pool_features(sets[0][1], T, want_img=True, want_rows=False)
    lin = ve._packed_projection(dev)
    rows["video projection (linear_tc 512->768, image in/out)"] = timeit(lambda i: lin.forward_img(B, False, x_img=pimg, want_rows=False, want_img=True))
    _, vimg = lin.forward_img(B, False, x_img=pimg, want_rows=False, want_img=True)
    rows["video projection head (mlp2, K=768)"] = timeit(lambda i: xm.video_proj.forward_fused(vimg, B))
    fo = {}
    rows["fusion layer + head + scores (head_tc, pre-layer)"] = timeit(lambda i: fus.forward_scores_img(enc["cls_img"], vimg, B, fo))
    ip, ip_img = xm.imu_proj.forward_fused(enc["cls_img"], B)
    vp, vp_img = xm.video_proj.forward_fused(vimg, B)
    work = similarity_img_work(B, B, dev)
    loss = torch.zeros((), dtype=torch.float64, device=dev)
    rows["similarity + loss (images)"] = timeit(lambda i: similarity_img_native(ip_img, B, vp_img, B, 256, work=work, loss=loss))
    pipe = cm.CrossModalOODPipeline(clf, xm, maha, frames=T, precision="bf16", fusion=fus)
    graphs = [pipe.capture(*s) for s in sets]
    rows["whole step, one stream (graph replay)"] = timeit(lambda i: graphs[i % 8][0].replay(), n=100)
    total = sum(v for k, v in rows.items() if not k.startswith("whole"))
    print(f"batch {B}: per-launch CUDA-event times (us), each kernel alone")
    for k, v in rows.items():
        print(f"  {v:8.2f}  {k}")
    print(f"  {total:8.2f}  sum of the 7 launches")


if __name__ == "__main__":
    main()
