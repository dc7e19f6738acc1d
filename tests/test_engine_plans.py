"""Host logic of the general tensor-core engine (csrc/tc_gen.cu) WITHOUT a GPU: the planners' output - MMA lists, K slabs,
weight gather tables, accumulator roles, scatter tables - is interpreted with numpy (tests/engine_sim.py) the way the
tcgen05 kernels read it, and the result is compared with the oracle's layer functions (src/abstract_cvae.py:30-33, 81-89
with TF SAME semantics) on bf16-rounded operands.  Plain-bf16 plans must reproduce the fp64 reference of the rounded
operands to rounding noise; hi + lo plans must reproduce the UNROUNDED reference to 2^-16."""
import numpy as np
import pytest
import torch

import engine_sim as S
from kcvae_testlib import O


def _t(a):
    return torch.from_numpy(np.asarray(a, np.float64))


def _rand(rng, *shape, scale=1.0):
    return (rng.standard_normal(shape) * scale).astype(np.float32)


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _conv_plan(kind, in_layout, Ck, Cn, KCk, w_mode, flip, split, Hg, Wg, w_stride=0, ones_col1=0, ones_src=0, w_col0=0):
    return S.ConvPlan(S.dump(0, [kind, in_layout, Ck, Cn, KCk, w_mode, flip, split, w_stride, ones_col1, ones_src, w_col0, Hg, Wg]))


def _wgrad_plan(kind, flip, s_layout, s_KC, u_layout, u_KC, Cs, Cu, w_mode, Hg, Wg):
    return S.WgradPlan(S.dump(1, [kind, flip, s_layout, s_KC, u_layout, u_KC, Cs, Cu, w_mode, Hg, Wg]))


def kc16(c):
    return (c + 15) // 16 * 2


# ------------------------------------------------------------------------------------------ forward-type products
@pytest.mark.parametrize("Ci,Co,H,W,split,x3", [(3, 8, 8, 12, 1, 1), (3, 20, 10, 70, 0, 1), (16, 5, 8, 12, 1, 0), (40, 24, 12, 8, 0, 0)])
def test_conv_s2_plan(Ci, Co, H, W, split, x3):
    rng = np.random.default_rng(Ci + Co)
    x = rng.random((2, H, W, Ci), dtype=np.float32)
    w = _rand(rng, 3, 3, Ci, Co, scale=0.3)
    layout, KC = (S.X3, 1) if x3 else (S.S2D, kc16(Ci))
    plan = _conv_plan(S.CONV_S2, layout, Ci, Co, KC, 0, 0, split, H // 2, W // 2)
    got = S.conv_output(plan, S.run_conv(plan, S.pack_planes(x, layout, KC, split), w, H // 2, W // 2), Co)
    if split:
        want = O.conv2d_s2_same(_t(x), _t(w), torch.zeros(Co, dtype=torch.float64)).numpy()
        assert _rel(got, want) < 5e-5
    else:
        want = O.conv2d_s2_same(_t(S.bf16(x)), _t(S.bf16(w)), torch.zeros(Co, dtype=torch.float64)).numpy()
        assert _rel(got, want) < 1e-12


@pytest.mark.parametrize("Ci,Co,h,w,split", [(16, 5, 5, 7, 1), (5, 32, 6, 33, 0), (24, 130, 4, 6, 0)])
def test_convT_s2_plan(Ci, Co, h, w, split):
    """both row-parity groups, the narrower N of the taps that only reach column parity 0, K padded with a dummy chunk"""
    rng = np.random.default_rng(Ci * 3 + Co)
    x = rng.random((2, h, w, Ci), dtype=np.float32)
    wt = _rand(rng, 3, 3, Co, Ci, scale=0.3)
    KC = kc16(Ci)
    if 2 * ((Co + 15) // 16 * 16) > 256:
        with pytest.raises(RuntimeError, match="256 accumulator columns"):
            _conv_plan(S.CONVT_S2, S.PLAIN, Ci, Co, KC, 1, 0, split, h, w)
        return
    plan = _conv_plan(S.CONVT_S2, S.PLAIN, Ci, Co, KC, 1, 0, split, h, w)
    got = S.conv_output(plan, S.run_conv(plan, S.pack_planes(x, S.PLAIN, KC, split), wt, h, w), Co)
    if split:
        want = O.conv2dT_s2_same(_t(x), _t(wt), torch.zeros(Co, dtype=torch.float64)).numpy()
        assert _rel(got, want) < 5e-5
    else:
        want = O.conv2dT_s2_same(_t(S.bf16(x)), _t(S.bf16(wt)), torch.zeros(Co, dtype=torch.float64)).numpy()
        assert _rel(got, want) < 1e-12


@pytest.mark.parametrize("flip", [1, 0])
def test_conv_s1_plan(flip):
    """output layer forward (flipped taps, weights [kh,kw,out,in]) and its data gradient (un-flipped, K side = the 3-channel gradient)"""
    rng = np.random.default_rng(9 + flip)
    Ci, Co, H, W = 16, 3, 9, 35
    wt = _rand(rng, 3, 3, Co, Ci, scale=0.3)
    if flip:
        x = rng.random((2, H, W, Ci), dtype=np.float32)
        plan = _conv_plan(S.CONV_S1, S.PLAIN, Ci, Co, kc16(Ci), 1, 1, 0, H, W)
        got = S.conv_output(plan, S.run_conv(plan, S.pack_planes(x, S.PLAIN, kc16(Ci), 0), wt, H, W), Co)
        want = O.conv2dT_s1_same(_t(S.bf16(x)), _t(S.bf16(wt)), torch.zeros(Co, dtype=torch.float64)).numpy()
    else:
        g = _rand(rng, 2, H, W, Co)
        plan = _conv_plan(S.CONV_S1, S.PLAIN, Co, Ci, 1, 0, 0, 0, H, W)           # one 8-channel plane: every MMA pairs it with a dummy chunk
        got = S.conv_output(plan, S.run_conv(plan, S.pack_planes(g, S.PLAIN, 1, 0), wt, H, W), Ci)
        xt = torch.zeros(2, H, W, Ci, dtype=torch.float64, requires_grad=True)
        y = O.conv2dT_s1_same(xt, _t(S.bf16(wt)), torch.zeros(Co, dtype=torch.float64))
        (want,) = torch.autograd.grad(y, xt, _t(S.bf16(g)))
        want = want.numpy()
    assert _rel(got, want) < 1e-12


def test_data_gradient_plans():
    """Conv2D s2 data gradient = ConvT-type product with HWIO weights; ConvT s2 data gradient = stride-2 product over S2D planes"""
    rng = np.random.default_rng(21)
    Ci, Co, H, W = 16, 5, 8, 12
    w = _rand(rng, 3, 3, Ci, Co, scale=0.3)
    g = _rand(rng, 2, H // 2, W // 2, Co)
    plan = _conv_plan(S.CONVT_S2, S.PLAIN, Co, Ci, kc16(Co), 1, 0, 0, H // 2, W // 2)
    got = S.conv_output(plan, S.run_conv(plan, S.pack_planes(g, S.PLAIN, kc16(Co), 0), w, H // 2, W // 2), Ci)
    xt = torch.zeros(2, H, W, Ci, dtype=torch.float64, requires_grad=True)
    (want,) = torch.autograd.grad(O.conv2d_s2_same(xt, _t(S.bf16(w)), torch.zeros(Co, dtype=torch.float64)), xt, _t(S.bf16(g)))
    assert _rel(got, want.numpy()) < 1e-12
    wt = _rand(rng, 3, 3, Co, Ci, scale=0.3)
    g2 = _rand(rng, 2, H, W, Co)
    plan = _conv_plan(S.CONV_S2, S.S2D, Co, Ci, kc16(Co), 0, 0, 0, H // 2, W // 2)
    got = S.conv_output(plan, S.run_conv(plan, S.pack_planes(g2, S.S2D, kc16(Co), 0), wt, H // 2, W // 2), Ci)
    xt = torch.zeros(2, H // 2, W // 2, Ci, dtype=torch.float64, requires_grad=True)
    (want,) = torch.autograd.grad(O.conv2dT_s2_same(xt, _t(S.bf16(wt)), torch.zeros(Co, dtype=torch.float64)), xt, _t(S.bf16(g2)))
    assert _rel(got, want.numpy()) < 1e-12


def test_dense_forward_and_gradient_plans():
    """decoder Dense from its long dimension: forward (columns = frames, hi + lo) and weight + bias gradient (a ones column)"""
    rng = np.random.default_rng(5)
    K, N, B = 16, 96, 5
    Wm, z, G = _rand(rng, K, N, scale=0.3), _rand(rng, B, K), _rand(rng, B, N)

    def rows_T(a, split):          # gen_pack_rows_T: [R][N] -> planes [R/8 (x2)][N/32][32][8]
        R = a.shape[0]
        KC = (R + 7) // 8
        hi, lo = S.hi_lo(a)
        out = np.zeros((1, KC * (2 if split else 1), N // 32, 32, 8))
        for r in range(R):
            out[0, r // 8, :, :, r % 8] = hi[r].reshape(N // 32, 32)
            if split:
                out[0, KC + r // 8, :, :, r % 8] = lo[r].reshape(N // 32, 32)
        return out

    plan = _conv_plan(S.DENSE, S.PLAIN, K, B, (K + 7) // 8, 1, 0, 1, N // 32, 32)
    D = S.run_conv(plan, rows_T(Wm, 1), z, N // 32, 32)[0, 0].reshape(N, -1)[:, :B]            # [n][frame]
    assert _rel(D.T, z.astype(np.float64) @ Wm.astype(np.float64)) < 5e-5
    src = np.concatenate([z.ravel(), [1.0]]).astype(np.float32)
    plan = _conv_plan(S.DENSE, S.PLAIN, B, K + 1, (B + 7) // 8, 0, 0, 0, N // 32, 32, w_stride=K, ones_col1=K + 1, ones_src=B * K)
    D = S.run_conv(plan, rows_T(G, 0), src, N // 32, 32)[0, 0].reshape(N, -1)[:, :K + 1]        # [n][k | ones]
    want = np.concatenate([S.bf16(z).T @ S.bf16(G), S.bf16(G).sum(0, keepdims=True)], 0)
    assert _rel(D.T, want) < 1e-12


# ------------------------------------------------------------------------------------------ weight / bias gradients
@pytest.mark.parametrize("Ci,Co,H,W,x3", [(3, 16, 8, 32, 1), (3, 32, 12, 60, 2), (16, 5, 8, 32, 0), (64, 16, 4, 32, 0), (128, 8, 4, 32, 0)])
def test_conv_s2_weight_gradient_plan(Ci, Co, H, W, x3):
    """dense accumulators (small parity blocks) and the pruned (tap, parity) plan of wide layers; the bias gradient against ones"""
    rng = np.random.default_rng(31 + Ci)
    x = rng.random((2, H, W, Ci), dtype=np.float32)
    g = _rand(rng, 2, H // 2, W // 2, Co)
    layout, KC = (S.X27, 1) if x3 == 2 else ((S.X3, 1) if x3 else (S.S2D, kc16(Ci)))      # X27: the 27-value patches, one tap
    plan = _wgrad_plan(S.CONV_S2, 0, layout, KC, S.PLAIN, kc16(Co), Ci, Co, 0, H // 2, W // 2)
    dW, db = S.run_wgrad(plan, S.pack_planes(x, layout, KC, 0), S.pack_planes(g, S.PLAIN, kc16(Co), 0), H // 2, W // 2)
    wt = torch.zeros(3, 3, Ci, Co, dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    gw, gb = torch.autograd.grad(O.conv2d_s2_same(_t(S.bf16(x)), wt, bt), (wt, bt), _t(S.bf16(g)))
    assert _rel(dW.reshape(3, 3, Ci, Co), gw.numpy()) < 1e-12
    assert _rel(db, gb.numpy()) < 1e-12


@pytest.mark.parametrize("Ci,Co,h,w", [(16, 5, 4, 16), (32, 128, 3, 16), (128, 64, 2, 16)])
def test_convT_s2_weight_gradient_plan(Ci, Co, h, w):
    rng = np.random.default_rng(41 + Ci)
    x = rng.random((2, h, w, Ci), dtype=np.float32)
    g = _rand(rng, 2, 2 * h, 2 * w, Co)
    plan = _wgrad_plan(S.CONVT_S2, 0, S.PLAIN, kc16(Ci), S.S2D, kc16(Co), Ci, Co, 1, h, w)
    dW, db = S.run_wgrad(plan, S.pack_planes(x, S.PLAIN, kc16(Ci), 0), S.pack_planes(g, S.S2D, kc16(Co), 0), h, w)
    wt = torch.zeros(3, 3, Co, Ci, dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    gw, gb = torch.autograd.grad(O.conv2dT_s2_same(_t(S.bf16(x)), wt, bt), (wt, bt), _t(S.bf16(g)))
    assert _rel(dW.reshape(3, 3, Co, Ci), gw.numpy()) < 1e-12
    assert _rel(db, gb.numpy()) < 1e-12


def test_output_layer_weight_gradient_plan():
    rng = np.random.default_rng(51)
    Ci, Co, H, W = 64, 3, 5, 30
    x = rng.random((2, H, W, Ci), dtype=np.float32)
    g = _rand(rng, 2, H, W, Co)
    plan = _wgrad_plan(S.CONV_S1, 1, S.PLAIN, kc16(Ci), S.PLAIN, 1, Ci, Co, 1, H, W)
    dW, db = S.run_wgrad(plan, S.pack_planes(x, S.PLAIN, kc16(Ci), 0), S.pack_planes(g, S.PLAIN, 1, 0), H, W)
    wt = torch.zeros(3, 3, Co, Ci, dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    gw, gb = torch.autograd.grad(O.conv2dT_s1_same(_t(S.bf16(x)), wt, bt), (wt, bt), _t(S.bf16(g)))
    assert _rel(dW.reshape(3, 3, Co, Ci), gw.numpy()) < 1e-12
    assert _rel(db, gb.numpy()) < 1e-12


def test_planner_refuses_what_the_kernels_cannot_run():
    with pytest.raises(RuntimeError, match="no tile width"):
        _wgrad_plan(S.CONV_S2, 0, S.S2D, 2, S.PLAIN, 2, 16, 16, 0, 8, 37)              # 37 has no divisor in [8, 31]
    with pytest.raises(RuntimeError, match="S2D / X3 input"):
        _conv_plan(S.CONV_S2, S.PLAIN, 16, 16, 2, 0, 0, 0, 8, 8)


def test_long_k_dense_forward_plan_with_hi_lo_quadrants():
    """encoder Dense (src/abstract_cvae.py:41-44) as a pixel-K product: K = the flattened activation (padded to a multiple of 32
    with a "ones pixel" that carries the bias), both tensors with lo planes behind the hi planes, every output the sum of the
    hi*hi, hi*lo and lo*hi accumulators."""
    rng = np.random.default_rng(61)
    F, E, B = 75, 16, 5
    Fp = (F + 1 + 31) // 32 * 32
    flat, Wm, bias = rng.random((B, F), dtype=np.float32), _rand(rng, F, E, scale=0.3), _rand(rng, E, scale=0.1)
    KCb, KCe = (B + 7) // 8, (E + 7) // 8
    xh, xl = S.hi_lo(flat)
    wh, wl = S.hi_lo(Wm)
    bh, bl = S.hi_lo(bias)
    U = np.zeros((1, 2 * KCb, Fp // 32, 32, 8))           # [hi | lo][i][8 frames]
    Sp = np.zeros((1, 2 * KCe, Fp // 32, 32, 8))          # [hi | lo][i][8 outputs]
    Uf, Sf = U.reshape(1, 2 * KCb, Fp, 8), Sp.reshape(1, 2 * KCe, Fp, 8)
    for b in range(B):
        Uf[0, b // 8, :F, b % 8] = xh[b]
        Uf[0, KCb + b // 8, :F, b % 8] = xl[b]
        Uf[0, b // 8, F, b % 8] = 1.0                      # the ones pixel
    for e in range(E):
        Sf[0, e // 8, :F, e % 8] = wh[:, e]
        Sf[0, KCe + e // 8, :F, e % 8] = wl[:, e]
        Sf[0, e // 8, F, e % 8] = bh[e]
        Sf[0, KCe + e // 8, F, e % 8] = bl[e]
    plan = S.WgradPlan(S.dump(1, [S.DENSE, 0, S.PLAIN, 2 * KCe, S.PLAIN, 2 * KCb, E, B, 1, Fp // 32, 32, 1]))
    out, _ = S.run_wgrad(plan, Sp, U, Fp // 32, 32)
    want = flat.astype(np.float64) @ Wm.astype(np.float64) + bias
    assert _rel(out.reshape(B, E), want) < 5e-5
