"""Development aid (test infrastructure): stage-by-stage check of the bf16 tcgen05 encoder against
the float64 oracle.  Run on a GPU box:  timeout 120 python tools/debug_bf16.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_imu_video_ood_har_b200 as cm
from oracle import oracle, weights as W

DEV = "cuda:0"


def stages_oracle(x, sd, dims, prefix="imu_encoder."):
    """h0, h after LN1 of layer 0, h after layer 0 (float64), rows laid out (B, S, 128)."""
    dt = torch.float64
    t = lambda k: torch.from_numpy(np.asarray(sd[prefix + k])).to(dt)
    x = torch.from_numpy(x).to(dt)
    B = x.shape[0]
    d, H = 128, 8
    emb = oracle.patch_embed(x, sd, dims, prefix, dt)
    tok = torch.cat([t("cls_token").expand(B, -1, -1), emb.reshape(B, -1, d)], 1)
    S = dims.seq
    h = tok[:, :S] + t("pos_encoding")[:, :S]
    out = [h.clone()]
    p = "transformer.layers.0."
    qkv = h @ t(p + "self_attn.in_proj_weight").T + t(p + "self_attn.in_proj_bias")
    q, k, v = qkv.split(d, -1)
    sp = lambda z: z.reshape(B, S, H, 16).transpose(1, 2)
    att = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / 4.0, -1)
    a = (att @ sp(v)).transpose(1, 2).reshape(B, S, d)
    a = a @ t(p + "self_attn.out_proj.weight").T + t(p + "self_attn.out_proj.bias")
    h = oracle._layer_norm(h + a, t(p + "norm1.weight"), t(p + "norm1.bias"))
    out.append(h.clone())
    f = torch.relu(h @ t(p + "linear1.weight").T + t(p + "linear1.bias")) @ t(p + "linear2.weight").T + t(p + "linear2.bias")
    h = oracle._layer_norm(h + f, t(p + "norm2.weight"), t(p + "norm2.bias"))
    out.append(h.clone())
    return out


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 250
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 19
    dims = W.Dims(imu_window=L)
    cfg = cm.default_config(imu_window_size=L)
    clf = cm.IMUClassifier(cm.IMUEncoder(cfg), cfg)
    sd = W.classifier_state(11, dims)
    clf.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    clf = clf.to(DEV).eval()
    xs = W.imu_windows(21, n, dims)
    x = torch.from_numpy(xs).to(DEV)
    enc = clf.imu_encoder
    S = dims.seq
    blob = enc.packed_blob(x.device, S)
    N = cm._native
    want = stages_oracle(xs, sd, dims)
    # biases folded into the residual dump: stage 0 holds h0 + b_o(0); stage 1 holds h1 + b_2(0); stage 2 h2 + b_o(1)
    def bo_fold(l):   # b_o + W_o b_v (the value bias is folded into the out-proj bias at pack time)
        pre = f"imu_encoder.transformer.layers.{l}.self_attn."
        return sd[pre + "out_proj.bias"] + sd[pre + "out_proj.weight"] @ sd[pre + "in_proj_bias"][256:]
    add = [bo_fold(0), sd["imu_encoder.transformer.layers.0.linear2.bias"], bo_fold(1)]
    tiles = (n + 7) // 8
    for stage in (0, 1, 2):
        dump = torch.full((tiles * 128, 128), float("nan"), device=DEV)
        cls = torch.empty(n, 128, device=DEV)
        prog = torch.zeros(256 * 16, dtype=torch.int32).pin_memory()
        N.check(N.lib().cmhar_debug_imu_bf16(blob.data_ptr(), x.data_ptr(), n, x.stride(0), stage, dump.data_ptr(), cls.data_ptr(), prog.data_ptr(), N.stream_ptr(x.device)))
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("kernel failed:", str(e).splitlines()[0])
            print("progress codes per warp (block 0):", prog[:16].tolist())
            raise SystemExit(1)
        got = dump.view(tiles * 8, 16, 128)[:n, :S].cpu().numpy() - add[stage]
        w = want[stage].numpy()
        err = np.abs(got - w)
        print(f"stage {stage}: max abs err {err.max():.4e}  (max |want| {np.abs(w).max():.3f})  nan={np.isnan(got).sum()}  "
              f"worst row/col {np.unravel_index(np.nanargmax(err), err.shape)}")
        if err.max() > 0.5 or np.isnan(got).any():
            np.set_printoptions(precision=3, suppress=True, linewidth=200)
            print("got[0,:3,:8]\n", got[0, :3, :8], "\nwant[0,:3,:8]\n", w[0, :3, :8])
            e_rows = err.reshape(-1, 128).max(1)
            print("per-row max err (first 32 rows):", e_rows[:32])
            print("per-col max err:", err.reshape(-1, 128).max(0)[:128])
    wl, wc = oracle.imu_classifier(xs, sd, dims, dtype=torch.float64)
    sc = clf.forward_scores(x, precision="bf16", want_cls=True)
    torch.cuda.synchronize()
    rel = lambda g, t_: float(np.abs(g.cpu().numpy() - t_.numpy()).max() / np.abs(t_.numpy()).max())
    print(f"final: cls rel err {rel(sc['cls'], wc):.4e}   logits rel err {rel(sc['logits'], wl):.4e}   "
          f"pred match {(sc['pred'].cpu().numpy() == oracle.predict(wl)).mean():.3f}")
    _, tok = enc(x) if False else (None, None)


if __name__ == "__main__":
    main()
