"""Development aid: pinned host->device copy bandwidth of this box (the e2e metric's PCIe-bound term), one stream
and two streams, 67 MB (one step's feature maps) and 512 MB.   python tools/h2d_probe.py"""
import time, torch
dev = torch.device("cuda:0")
for mb in (67, 512):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("1 stream", "2 streams (halves)"):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10):
                if mode[0] == "1":
                    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
                else:
                    with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
                    with torch.cuda.stream(s2): d[n // 2:].copy_(h[n // 2:], non_blocking=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"H2D {mb:4d} MB, {mode:20s}: {10 * n / dt / 1e9:6.1f} GB/s")
