"""Development aid: per-tap error of one generic-engine weight gradient (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import test_gpu_gen_engine as T
from kcvae_testlib import O

def conv_s2(Ci, Co, H, W, x3=0, B=3):
    rng = np.random.default_rng(17 * Ci + Co)
    x = rng.random((B, H, W, Ci), dtype=np.float32)
    g = T._rand(rng, B, H // 2, W // 2, Co)
    wt = T._t64(np.zeros((3, 3, Ci, Co))).requires_grad_(True)
    bt = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    y = O.conv2d_s2_same(T._t64(x), wt, bt)
    gw, gb = torch.autograd.grad(y, (wt, bt), T._t64(g))
    dW, db = T.gen_wgrad(T.CONV_S2, 0, 0, x, g, (3, 3, Ci, Co), Co, s_x3=x3)
    gw = gw.numpy()
    print(f"conv_s2 {Ci}->{Co} {H}x{W}: total err {T._err(dW, gw):.4f} bias err {T._err(db, gb.numpy()):.4f}")
    for kh in range(3):
        for kw in range(3):
            e = np.abs(dW[kh, kw] - gw[kh, kw]).max() / np.abs(gw).max()
            blocks = [np.abs(dW[kh, kw, c0:c0 + 16] - gw[kh, kw, c0:c0 + 16]).max() / np.abs(gw).max() for c0 in range(0, Ci, 16)]
            print(f"  tap ({kh},{kw}) err {e:.4f}  per 16-ci block: " + " ".join(f"{b:.3f}" for b in blocks))

if __name__ == "__main__":
    conv_s2(64, 128, 16, 60)
    conv_s2(128, 32, 12, 50)
    conv_s2(32, 5, 24, 50)
