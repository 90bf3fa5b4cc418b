"""Development aid: per-warp (code, clock64) timeline of block 0 of the bf16 kernel, merged over the
MMA-issuer warp and epilogue warp 0 for one steady-state layer.
   [CMHAR_ABLATE=mask] python tools/timeline_bf16.py [n_windows]"""
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
enc = clf.imu_encoder
blob = enc.packed_blob(x.device, 16)
N = cm._native
CAP = 1024
for rep in range(2):
    tlog = torch.zeros(18 * CAP * 2, dtype=torch.int64, device="cuda")
    N.check(N.lib().cmhar_debug_imu_bf16(blob.data_ptr(), x.data_ptr(), n, x.stride(0), 100, tlog.data_ptr(), None, None, N.stream_ptr(x.device)))
    torch.cuda.synchronize()
t = tlog.view(18, CAP, 2).cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True); np.save(f"gpurun_out/timeline_bf16_a{os.environ.get('CMHAR_ABLATE', '0')}.npy", t)
names = {3: "M wait HA(patch)", 4: "M wait HA(qkv)", 5: "M wait QKV smem", 6: "M wait P", 8: "M wait O", 9: "M wait HA(ffn)", 10: "M wait HID",
         11: "E wait R(patch)", 12: "E wait QKV acc", 13: "E wait S", 14: "E P published", 15: "E wait O acc", 150: "E O published", 16: "E wait R(outproj)",
         160: "E LN1 published", 17: "E wait FFN1 acc", 18: "E wait R(ffn2)"}
NQ = 4 if os.environ.get("CMHAR_EPI_WARPS") == "16" else 2
ev = []
for warp, tag in ((4 * NQ, "M"), (0, "E")):
    e = t[warp]; e = e[e[:, 1] > 0]
    ev += [(int(clk), tag, int(code)) for code, clk in e]
ev.sort()
# steady state: second tile of the MMA warp = between the 2nd and 3rd occurrence of code 3
starts = [clk for clk, tag, code in ev if tag == "M" and code == 3]
qk = [clk for clk, tag, code in ev if tag == "M" and code == 4 and clk > starts[1]]
print(f"ablate={os.environ.get('CMHAR_ABLATE', '0')}  tile span {starts[2] - starts[1]} cycles; layer spans {[qk[i + 1] - qk[i] for i in range(3)]}")
lo, hi = qk[1], qk[2]
prev = {"M": lo, "E": lo}
for clk, tag, code in ev:
    if clk < lo or clk > hi: continue
    nm = names.get(code % 1000, str(code % 1000))
    arrived = " (passed)" if code >= 1000 and code % 1000 in names else ""
    pad = "" if tag == "M" else " " * 44
    print(f"{clk - lo:7d} {pad}{tag} +{clk - prev[tag]:5d} {nm}{arrived}")
    prev[tag] = clk
