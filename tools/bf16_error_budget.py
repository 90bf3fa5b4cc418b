"""Development aid: CPU emulation of the bf16 kernel's rounding points, to attribute the logit
error to each quantisation (test infrastructure; imports the oracle)."""
import itertools, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle, weights as W

def emulate(x, sd, dims, R, prefix="imu_encoder."):
    r = lambda t, key: t.bfloat16().float() if R.get(key, True) else t
    dt = torch.float32
    t = lambda k: torch.from_numpy(np.asarray(sd[prefix + k])).to(dt)
    x = torch.from_numpy(x).to(dt)
    B, d, H = x.shape[0], 128, 8
    N = dims.num_patches
    patches = x[:, 0, :16 * N].reshape(B, N, 16)
    emb = r(patches, "x") @ r(t("patch_embed.projections.0.weight"), "wp").T + t("patch_embed.projections.0.bias")
    tok = torch.cat([t("cls_token").expand(B, -1, -1), emb], 1)
    S = dims.seq
    h = tok[:, :S] + t("pos_encoding")[:, :S]
    for l in range(dims.layers):
        p = f"transformer.layers.{l}."
        hA = r(h, "hA")
        qkv = hA @ r(t(p + "self_attn.in_proj_weight"), "w").T + t(p + "self_attn.in_proj_bias")
        q, k, v = qkv.split(d, -1)
        sp = lambda z: z.reshape(B, S, H, 16).transpose(1, 2)
        att = torch.softmax(sp(r(q, "q")) @ sp(r(k, "k")).transpose(-1, -2) / 4.0, -1)
        a = (r(att, "p") @ sp(r(v, "v"))).transpose(1, 2).reshape(B, S, d)
        a = r(a, "o") @ r(t(p + "self_attn.out_proj.weight"), "w").T + t(p + "self_attn.out_proj.bias")
        h = oracle._layer_norm(h + a, t(p + "norm1.weight"), t(p + "norm1.bias"))
        f = torch.relu(r(h, "hA") @ r(t(p + "linear1.weight"), "w").T + t(p + "linear1.bias"))
        f = r(f, "hid") @ r(t(p + "linear2.weight"), "w").T + t(p + "linear2.bias")
        h = oracle._layer_norm(h + f, t(p + "norm2.weight"), t(p + "norm2.bias"))
    tokens = oracle._layer_norm(h, t("norm.weight"), t("norm.bias"))
    return tokens[:, 0]

dims = W.Dims()
for seed in (11, 13):
    sd = W.classifier_state(seed, dims)
    x = W.imu_windows(21, 256, dims)
    wl, wc = oracle.imu_classifier(x, sd, dims, dtype=torch.float64)
    def report(name, R):
        cls = emulate(x, sd, dims, R)
        lg = oracle.classifier_head(cls, sd, dims)
        e_c = float((cls.double() - wc).abs().max() / wc.abs().max())
        e_l = float((lg.double() - wl).abs().max() / wl.abs().max())
        print(f"seed {seed} {name:28s} cls rel {e_c:.4f}  logits rel {e_l:.4f}  flips {(lg.argmax(1) != wl.argmax(1)).sum().item()}")
    keys = ["x", "wp", "hA", "w", "q", "k", "v", "p", "o", "hid"]
    report("all rounded", {})
    report("none", {k: False for k in keys})
    for k in keys:
        report(f"only {k}", {kk: (kk == k) for kk in keys})
    report("all but x,wp", {"x": False, "wp": False})
    report("all but q,k", {"q": False, "k": False})
    report("all but q,k,x,wp", {"q": False, "k": False, "x": False, "wp": False})
    report("all but hA", {"hA": False})
    report("all but w", {"w": False})
