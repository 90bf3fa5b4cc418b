"""ncu target: the two big kernels of the configs[1] step at ITS launch size (256 windows / 256 clips), rotating inputs larger than L2:
   ncu --set full -k regex:imu_forward_bf16_kernel|video_pool_kernel -s 8 -c 2 ... python tools/profile_step256.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import crossmodal_imu_video_ood_har_b200 as cm
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native
dev = torch.device("cuda:0")
cfg, clf, xm, fus = bench.build_modules(dev)
sets = bench.synth_inputs(dev, 256, 6, 0)          # 6 x 67 MB > L2
for i in range(6):
    imu, fmap = sets[i]
    imu_forward_native(clf.imu_encoder, None, None, imu, want_cls=True, precision="bf16", want_cls_img=True)
    xm.video_encoder.pool_features(fmap, 16, want_img=True, want_rows=False)
torch.cuda.synchronize()
print("ok")
