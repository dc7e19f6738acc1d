"""Definitional numpy loops vs the oracle's torch calls, and
the hand-derived latent gradient vs autograd in fp64."""
import numpy as np
import torch

from oracle import kcvae_oracle as O


def _conv_s2_def(x, w, b):
    n, H, W, ci = x.shape
    co = w.shape[3]
    oh, ow = (H + 1) // 2, (W + 1) // 2
    pt = max((oh - 1) * 2 + 3 - H, 0) // 2
    pl = max((ow - 1) * 2 + 3 - W, 0) // 2
    y = np.zeros((n, oh, ow, co))
    for i in range(oh):
        for j in range(ow):
            for kh in range(3):
                for kw in range(3):
                    r, c = 2 * i + kh - pt, 2 * j + kw - pl
                    if 0 <= r < H and 0 <= c < W:
                        y[:, i, j, :] += x[:, r, c, :] @ w[kh, kw]
    return y + b


def _convT_s2_def(x, w, b):
    n, h, wd, ci = x.shape
    co = w.shape[2]
    y = np.zeros((n, 2 * h + 1, 2 * wd + 1, co))
    for i in range(h):
        for j in range(wd):
            for kh in range(3):
                for kw in range(3):
                    y[:, 2 * i + kh, 2 * j + kw, :] += x[:, i, j, :] @ w[kh, kw].T
    return y[:, :2 * h, :2 * wd, :] + b


def _convT_s1_def(x, w, b):
    n, h, wd, ci = x.shape
    co = w.shape[2]
    y = np.zeros((n, h, wd, co))
    for i in range(h):
        for j in range(wd):
            for kh in range(3):
                for kw in range(3):
                    r, c = i + 1 - kh, j + 1 - kw
                    if 0 <= r < h and 0 <= c < wd:
                        y[:, i, j, :] += x[:, r, c, :] @ w[kh, kw].T
    return y + b


def test_layer_semantics_match_definitions():
    rng = np.random.default_rng(0)
    for (H, W) in [(8, 12), (7, 9)]:
        x = rng.standard_normal((2, H, W, 3))
        w = rng.standard_normal((3, 3, 3, 4))
        b = rng.standard_normal(4)
        got = O.conv2d_s2_same(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
        np.testing.assert_allclose(got, _conv_s2_def(x, w, b), atol=1e-12)
    x = rng.standard_normal((2, 5, 6, 4))
    w = rng.standard_normal((3, 3, 3, 4))   # [kh,kw,out,in]
    b = rng.standard_normal(3)
    got = O.conv2dT_s2_same(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    np.testing.assert_allclose(got, _convT_s2_def(x, w, b), atol=1e-12)
    got = O.conv2dT_s1_same(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    np.testing.assert_allclose(got, _convT_s1_def(x, w, b), atol=1e-12)


def test_global_latent_gradient_formula():
    """g_z terms the CUDA latent kernel implements == autograd of the oracle loss."""
    rng = np.random.default_rng(3)
    z = torch.tensor(rng.standard_normal((6, 5)) * 0.7 + 0.3, requires_grad=True)
    t, wk, ws, wl = 3.0, 0.3, 0.2, 0.1
    mu, var = z.mean(), z.var(unbiased=False)
    s = (z - mu) / var.sqrt()
    skew, K = (s ** 3).mean(), (s ** 4).mean()
    loss = wk * (t - K).abs() + ws * skew.abs() + wl * z.abs().mean()
    g, = torch.autograd.grad(loss, z)
    zz = z.detach()
    N = zz.numel()
    sig = zz.var(unbiased=False).sqrt()
    sd = (zz - zz.mean()) / sig
    sk, kk = (sd ** 3).mean(), (sd ** 4).mean()
    man = (wk * (-torch.sign(t - kk)) * 4 / (N * sig) * (sd ** 3 - sk - kk * sd)
           + ws * torch.sign(sk) * 3 / (N * sig) * (sd ** 2 - 1 - sk * sd)
           + wl * torch.sign(zz) / N)
    np.testing.assert_allclose(man.numpy(), g.numpy(), atol=1e-13)


def test_single_latent_gradient_formula():
    rng = np.random.default_rng(4)
    z = torch.tensor(rng.standard_normal((9, 4)) * 0.7 + 0.3, requires_grad=True)
    t, wk, ws, wl = 3.0, 0.3, 0.2, 0.1
    mu = z.mean(0)
    sig = z.std(0, unbiased=False)
    s = (z - mu) / sig
    skew, K = (s ** 3).mean(0), (s ** 4).mean(0)
    loss = wk * ((K - t) ** 2).mean() + ws * (skew ** 2).mean() + wl * (mu ** 2).sum().sqrt()
    g, = torch.autograd.grad(loss, z)
    zz = z.detach()
    B, L = zz.shape
    mu, sig = zz.mean(0), zz.std(0, unbiased=False)
    sd = (zz - mu) / sig
    sk, kk = (sd ** 3).mean(0), (sd ** 4).mean(0)
    l2 = (mu ** 2).sum().sqrt()
    man = (wk * 2 * (kk - t) / L * 4 / (B * sig) * (sd ** 3 - sk - kk * sd)
           + ws * 2 * sk / L * 3 / (B * sig) * (sd ** 2 - 1 - sk * sd)
           + wl * mu / (l2 * B))
    np.testing.assert_allclose(man.numpy(), g.numpy(), atol=1e-13)


def test_adam_first_step_is_lr_sign():
    """TF-Adam identity: after step 1, m/(sqrt(v)) = sign(g) so |dp| ~= lr."""
    opt = O.Adam(1e-3)
    p = [torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64)]
    g = [torch.tensor([0.5, -0.25, 4.0], dtype=torch.float64)]
    opt.apply_gradients(g, p)
    np.testing.assert_allclose(p[0].numpy(), [1 - 1e-3, -2 + 1e-3, 3 - 1e-3], atol=1e-8)
