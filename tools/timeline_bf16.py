"""Development aid: per-warp (code, clock64) timeline of block 0 of the bf16 kernel.
   python tools/timeline_bf16.py [n_windows]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8 * 4
cfg = cm.default_config()
torch.manual_seed(0)
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to("cuda").eval()
x = torch.randn(n, 6, 250, device="cuda")
enc = clf.imu_encoder
blob = enc.packed_blob(x.device, 16)
N = cm._native
CAP = 1024
for rep in range(2):
    tlog = torch.zeros(18 * CAP * 2, dtype=torch.int64, device="cuda")
    N.check(N.lib().cmhar_debug_imu_bf16(blob.data_ptr(), x.data_ptr(), n, x.stride(0), 100, tlog.data_ptr(), None, None, N.stream_ptr(x.device)))
    torch.cuda.synchronize()
t = tlog.view(18, CAP, 2).cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True); np.save("gpurun_out/timeline_bf16.npy", t)
names = {3: "MMA wait HA(patch)", 4: "MMA wait HA(qkv)", 5: "MMA wait QKV smem", 6: "MMA wait P", 8: "MMA wait O", 9: "MMA wait HA(ffn)", 10: "MMA wait HID",
         11: "EPI wait R(patch)", 12: "EPI wait QKV acc", 13: "EPI wait S", 15: "EPI wait O acc", 16: "EPI wait R(outproj)", 17: "EPI wait FFN1 acc", 18: "EPI wait R(ffn2)", 2: "MMA wait Wfull", 1: "LOAD wait Wempty"}
NQ = 4 if os.environ.get("CMHAR_EPI_WARPS") == "16" else 2
for warp, label in ((4 * NQ, "MMA issuer"), (0, "epilogue warp 0")):
    ev = t[warp]; ev = ev[ev[:, 1] > 0]
    if len(ev) == 0: continue
    t0 = ev[0, 1]
    print(f"== {label}: {len(ev)} events; first tile span")
    # print first tile only: until code 3/11 appears the second time
    first = 3 if warp == 4 * NQ else 11
    seen = 0
    agg = {}
    prev_t = t0
    for code, clk in ev:
        if code == first:
            seen += 1
            if seen == 3: break
        if seen == 2:      # second tile (steady state)
            key = int(code)
            dt = clk - prev_t
            agg.setdefault(key, [0, 0]); agg[key][0] += dt; agg[key][1] += 1
        prev_t = clk
    tot = sum(v[0] for v in agg.values())
    print(f"   steady-state tile: {tot} cycles")
    for key, (dt, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        kind = "time WAITING at" if key >= 1000 else "work BEFORE reaching"
        nm = names.get(key % 1000, str(key % 1000))
        print(f"   {dt:9d} cyc {100 * dt / tot:5.1f}%  x{cnt:3d}  {kind} [{nm}]")
