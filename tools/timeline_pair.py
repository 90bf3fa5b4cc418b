"""Development aid: (code, clock64) timeline of block 0 of the two-tiles-in-flight encoder kernel -- the MMA-issuing warp,
the producer lane and epilogue warp 0 of either side -- printed for one steady-state stretch.
   python tools/timeline_pair.py [n_windows] [first_cycle] [last_cycle]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_imu_video_ood_har_b200 as cm
cm._native.enable_dev_env()            # development tool: honour the CMHAR_* A/B switches of the environment
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8 * 4
cfg = cm.default_config()
torch.manual_seed(0)
clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg).to("cuda").eval()
x = torch.randn(n, 6, 250, device="cuda")
blob = clf.imu_encoder.packed_blob(x.device, 16)
N = cm._native
N.check(N.lib().cmhar_debug_set_option(b"enc_kernel", 2))
CAP = 1024
for rep in range(2):
    tlog = torch.zeros(20 * CAP * 2, dtype=torch.int64, device="cuda")
    N.check(N.lib().cmhar_debug_imu_bf16(blob.data_ptr(), x.data_ptr(), n, x.stride(0), 100, tlog.data_ptr(), None, None, N.stream_ptr(x.device)))
    torch.cuda.synchronize()
t = tlog.view(20, CAP, 2).cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/timeline_pair.npy", t)
A = {-1: "patch", 0: "K", 1: "V^T", 2: "Q", 3: "scores", 4: "PV", 5: "outproj"}
E = {11: "h0", 12: "K/V drain", 13: "Q", 14: "softmax", 15: "O", 16: "LN1", 17: "hid", 18: "LN2"}
ev = []
for w in (16, 17):
    for code, clk in t[w][t[w][:, 1] > 0]:
        side, r = divmod(int(code), 1000)
        done = r >= 500
        r = r - 500 if done else r
        name = f"a:{A[r - 10]}" if r < 100 else f"f:{r - 100}"
        ev.append((int(clk), "M", f"side{side} {name} {'issued' if done else 'start'}"))
for w in (18, 19):
    for code, clk in t[w][t[w][:, 1] > 0]:
        ev.append((int(clk), "P", f"chunk {int(code)}"))
for w, side in ((0, 0), (8, 1)):
    for code, clk in t[w][t[w][:, 1] > 0]:
        code = int(code)
        if code >= 200: nm = "published"
        elif code >= 100: nm = f"{E.get(code - 100, code - 100)} acc ready"
        else: nm = f"wait {E.get(code, code)}"
        ev.append((int(clk), f"E{side}", nm))
ev.sort()
t0 = ev[0][0]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
print(f"total span {ev[-1][0] - t0} cycles, {len(ev)} events; showing [{lo}, {hi})")
col = {"M": 0, "P": 34, "E0": 50, "E1": 80}
for clk, tag, nm in ev:
    if lo <= clk - t0 < hi:
        print(f"{clk - t0:8d} {' ' * col[tag]}{tag} {nm}")
