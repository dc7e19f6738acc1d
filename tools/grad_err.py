"""Development aid: measured relative-L2 gradient error of the tensor-core path per variable (GPU box)."""
import sys, os, glob
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from kcvae_testlib import O, make, frames, eps_for, small_config
import test_reference_goldens as G

def l2(a, b):
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / (np.linalg.norm(b) + 1e-30))

for path in G.GOLD:
    g, cfg, ws, grads, w_after = G.load(path)
    kind = "single" if cfg["model"].get("type") == "KurtosisSingle" else "global"
    m = G.model_class("cuda", kind)(cfg, precision="bf16")
    m.set_weights(ws)
    d, mg = m.loss_and_grads(g["x"], eps=g["eps"])
    print(os.path.basename(path), "tc", m.tc_status(), "max L2", max(l2(a, b) for a, b in zip(mg, grads)), [round(l2(a, b), 4) for a, b in zip(mg, grads)])
for shape in (dict(layers=(32,), enc=8, H=40, W=52, dec=8, latent=8), dict(layers=(32, 5), enc=16, H=224, W=300, dec=32, latent=32)):
    cfg = small_config(**shape)
    m, ws = make(cfg, "cuda", weight_gain=1.3, precision="bf16")
    x, eps = frames(cfg, 3), eps_for(cfg, 3)
    d, grads = m.loss_and_grads(x, eps=eps)
    od, ograds, _, _ = O.loss_and_grads(cfg, ws, x, eps)
    print(shape["layers"], "max L2", max(l2(a, b.numpy()) for a, b in zip(grads, ograds)), [round(l2(a, b.numpy()), 4) for a, b in zip(grads, ograds)])
cfg = O.readme_config()
m, ws = make(cfg, "cuda", precision="bf16")
x, eps = frames(cfg, 16), eps_for(cfg, 16)
d, grads = m.loss_and_grads(x, eps=eps)
od, ograds, _, _ = O.loss_and_grads(cfg, ws, x, eps)
print("readme B=16 max L2", max(l2(a, b.numpy()) for a, b in zip(grads, ograds)), [round(l2(a, b.numpy()), 4) for a, b in zip(grads, ograds)])
print("loss terms rel err", {k: abs(float(d[k]) - float(od[k])) / (abs(float(od[k])) + 1e-12) for k in od})
