"""Development aid: the co-resident ring pooling kernel alone, 256 clips (67 MB).  CMHAR_POOL_STAGE / CMHAR_POOL_CTAS_PER_SM."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
N = cm._native; dev = torch.device("cuda:0"); B = 256
sets = [torch.relu(torch.randn(B * 16, 512, 4, 4, device=dev)).to(torch.bfloat16) for _ in range(8)]
pooled = torch.empty(B, 512, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def f(i): N.check(N.lib().cmhar_video_pool(sets[i % 8].data_ptr(), 1, B, 16, 512, 16, pooled.data_ptr(), N.stream_ptr(dev)))
for i in range(8): f(i)
torch.cuda.synchronize(); e0.record()
for i in range(40): f(i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 40 * 1e3
print(f"stage {os.environ.get('CMHAR_POOL_STAGE', 'dflt'):>6s} ctas/sm {os.environ.get('CMHAR_POOL_CTAS_PER_SM', '1')}: {us:7.1f} us  {B * 262144 / us / 1e6:6.2f} TB/s")
