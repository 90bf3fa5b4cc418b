"""Development aid: the pooling kernels alone (flood / co-resident ring with L2 prefetch), 256 and 2 048 clips, checked against torch.
CMHAR_POOL_MODE (1 flood, 2 ring) / CMHAR_POOL_PF (ring stages prefetched into L2) / CMHAR_POOL_STAGE / CMHAR_POOL_CPT / CMHAR_POOL_CTAS_PER_SM."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
N = cm._native; dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tag = " ".join(f"{k[11:]}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("CMHAR_POOL_"))
for B in (256, 2048):
    ns = max(2, 600_000_000 // (B * 262144))
    sets = [torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev)).to(torch.bfloat16) for _ in range(ns)]
    pooled = torch.empty(B, 512, device=dev)
    img = cm.models.operand_image(B, 512, dev)
    def f(i): N.check(N.lib().cmhar_video_pool_img(sets[i % ns].data_ptr(), 1, B, 16, 512, 16, pooled.data_ptr(), img.data_ptr(), N.stream_ptr(dev)))
    f(0)
    want = sets[0].float().view(B, 16, 512, 16).mean(dim=(1, 3))
    err = (pooled - want).abs().max().item()
    for i in range(8): f(i)
    torch.cuda.synchronize(); e0.record()
    for i in range(40): f(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 40 * 1e3
    print(f"[{tag}] B={B}: {us:7.1f} us  {B * 262144 / us / 1e6:6.2f} TB/s  max|err| {err:.2e}  img sum {img.view(torch.int16).long().sum().item()}")
    del sets
