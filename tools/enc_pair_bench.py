"""GPU tool: encoder launch time, single-tile kernel (mode 1) vs two-tiles-in-flight kernel (mode 2), over batch sizes.
    python tools/enc_pair_bench.py [batch ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import crossmodal_imu_video_ood_har_b200 as cm  # noqa: E402
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
from crossmodal_imu_video_ood_har_b200.models import imu_forward_native  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024, 2368, 4096, 16384, 65536, 262144]
    dev = torch.device("cuda:0")
    cfg, clf, xm, fus = bench.build_modules(dev)
    N = cm._native
    for nb in sizes:
        xs = [torch.randn(nb, 6, 250, device=dev) for _ in range(max(2, min(8, 200_000_000 // (nb * 6000))))]
        outs = [dict() for _ in xs]
        row = []
        for mode in (1, 2):
            N.check(N.lib().cmhar_debug_set_option(b"enc_kernel", mode))
            fn = lambda i: imu_forward_native(clf.imu_encoder, None, None, xs[i % len(xs)], want_cls=True, precision="bf16", out=outs[i % len(xs)])
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            reps = max(5, min(200, 4_000_000 // nb))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row.append((ms, nb / ms / 1e3, bench.FLOP_ENC * nb / (ms * 1e-3) / 1e12))
        print(f"batch {nb:7d}: single {row[0][0] * 1e3:9.1f} us {row[0][1]:7.2f} M win/s {row[0][2]:6.1f} TFLOP/s | "
              f"pair {row[1][0] * 1e3:9.1f} us {row[1][1]:7.2f} M win/s {row[1][2]:6.1f} TFLOP/s | x{row[0][0] / row[1][0]:.2f}")
    N.check(N.lib().cmhar_debug_set_option(b"enc_kernel", 0))


if __name__ == "__main__":
    main()
