"""GPU tool: the conv / BN / ReLU IMU encoder, CUDA-core fp32 kernel vs tensor-core implicit-GEMM kernel, over batch sizes.
    python tools/bench_conv.py [batch ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm  # noqa: E402

FLOP = 2 * (250 * 32 * 30 + 125 * 64 * 160 + 63 * 128 * 320)      # per window


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 4096, 65536, 262144]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = cm.ConvIMUEncoder(cm.default_config()).to(dev).eval()
    for nb in sizes:
        xs = [torch.randn(nb, 6, 250, device=dev) for _ in range(max(2, min(8, 400_000_000 // (nb * 6000))))]
        out = torch.empty(nb, 128, device=dev)
        row = []
        for prec in ("fp32", "bf16"):
            fn = lambda i: enc.forward_native(xs[i % len(xs)], out=out, precision=prec)
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            reps = max(3, min(200, 4_000_000 // nb)) if prec == "bf16" else max(2, min(50, 400_000 // nb))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row.append((ms, nb / ms / 1e3, FLOP * nb / (ms * 1e-3) / 1e12, nb * 6512 / (ms * 1e-3) / 1e9))
        print(f"batch {nb:7d}: fp32 CUDA cores {row[0][0] * 1e3:9.1f} us {row[0][1]:7.2f} M win/s {row[0][2]:6.1f} TFLOP/s | "
              f"tensor cores {row[1][0] * 1e3:9.1f} us {row[1][1]:7.2f} M win/s {row[1][2]:6.1f} TFLOP/s {row[1][3]:7.1f} GB/s | x{row[0][0] / row[1][0]:.1f}")


if __name__ == "__main__":
    main()
