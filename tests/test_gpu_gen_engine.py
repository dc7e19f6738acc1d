"""General tensor-core convolution engine (csrc/tc_gen.cu) against the oracle's layer functions, one product at a time,
through the C ABI hooks kcvae_gen_conv_test / kcvae_gen_wgrad_test (-m gpu).

Reference = oracle.kcvae_oracle.conv2d_s2_same / conv2dT_s2_same / conv2dT_s1_same (src/abstract_cvae.py:30-33, 81-89 with
TF SAME semantics) evaluated in fp64 on the CPU; data and weight gradients come from torch autograd on those functions.
Tolerances: split (bf16 hi + lo operands) products are fp32-grade: 2e-5 of max|ref|; plain bf16 operands: 1.2e-2 of
max|ref| (operand rounding 2^-9 each, fp32 accumulation)."""
import ctypes as C
import importlib

import numpy as np
import pytest
import torch

from kcvae_testlib import O

pytestmark = pytest.mark.gpu

CONV_S2, CONVT_S2, CONV_S1 = 0, 1, 2
PRE_NONE, PRE_RELU, PRE_SIGMOID, PRE_BIAS = 0, 1, 2, 3


def _lib():
    return importlib.import_module("trustedai-cl-vae-ad_b200._lib").load()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def gen_conv(kind, w_mode, flip, split, pre, x, w, bias, mask, out_shape, in_x3=0, out_mode=0, mask_mode=0):
    lib = _lib()
    xd, wd = _dev(x), _dev(w)
    bd = _dev(bias) if bias is not None else None
    md = _dev(mask) if mask is not None else None
    out = torch.full(out_shape, float("nan"), dtype=torch.float32, device="cuda")
    B, Hi, Wi, Ck = x.shape
    rc = lib.gen_conv_test(kind, w_mode, flip, split, pre, in_x3, out_mode, mask_mode, _ptr(xd), _ptr(wd), _ptr(bd), _ptr(md), _ptr(out),
                           B, Hi, Wi, Ck, out_shape[-1], None)
    lib.check(rc, None)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def gen_wgrad(kind, w_mode, flip, s, u, w_shape, nbias, s_x3=0):
    lib = _lib()
    sd, ud = _dev(s), _dev(u)
    dW = torch.full((int(np.prod(w_shape)),), float("nan"), dtype=torch.float32, device="cuda")
    db = torch.full((nbias,), float("nan"), dtype=torch.float32, device="cuda")
    B, Hs, Ws, Cs = s.shape
    rc = lib.gen_wgrad_test(kind, w_mode, flip, s_x3, _ptr(sd), _ptr(ud), _ptr(dW), _ptr(db), B, Hs, Ws, Cs, u.shape[-1], None)
    lib.check(rc, None)
    torch.cuda.synchronize()
    return dW.cpu().numpy().reshape(w_shape), db.cpu().numpy()


def _err(got, want):
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


def _tol(split):
    return 2e-5 if split else 1.2e-2


def _rand(rng, *shape, scale=1.0):
    return (rng.standard_normal(shape) * scale).astype(np.float32)


def _t64(a):
    return torch.from_numpy(np.asarray(a, np.float64))


# ------------------------------------------------------------------------------------------ Conv2D k3 s2 (encoder)
@pytest.mark.parametrize("Ci,Co,H,W,split,x3,out_mode", [
    (3, 32, 24, 60, 1, 1, 2),        # README conv0: packed image planes, hi + lo, output stored space-to-depth
    (3, 32, 20, 44, 0, 1, 0),        # inference form (plain bf16)
    (32, 5, 24, 60, 1, 0, 0),        # README conv1: 16 + 16 planes -> K slabs
    (16, 24, 18, 30, 1, 0, 1),
    (64, 128, 16, 64, 0, 0, 2),      # cfg5 widths
    (128, 32, 24, 28, 0, 0, 0),
    (5, 7, 10, 14, 1, 0, 0),
])
def test_conv_s2_forward(Ci, Co, H, W, split, x3, out_mode):
    rng = np.random.default_rng(Ci * 100 + Co)
    B = 2
    x = rng.random((B, H, W, Ci), dtype=np.float32)
    w = _rand(rng, 3, 3, Ci, Co, scale=(2.0 / (9 * Ci)) ** 0.5)
    b = _rand(rng, Co, scale=0.1)
    want = torch.relu(O.conv2d_s2_same(_t64(x), _t64(w), _t64(b))).numpy()
    got = gen_conv(CONV_S2, 0, 0, split, PRE_RELU, x, w, b, None, want.shape, in_x3=x3, out_mode=out_mode)
    assert _err(got, want) < _tol(split), _err(got, want)


# ------------------------------------------------------------------------------------------ Conv2DTranspose k3 s2 (decoder)
@pytest.mark.parametrize("Ci,Co,h,w,split", [
    (32, 5, 12, 31, 1),              # README 32 -> few
    (5, 32, 12, 20, 0),              # README few -> 32 (K padded to 16)
    (64, 32, 9, 33, 0),              # cfg5
    (32, 128, 8, 16, 0),
    (128, 64, 8, 30, 0),             # 256 accumulator columns: one M-tile per buffer
])
def test_convT_s2_forward(Ci, Co, h, w, split):
    rng = np.random.default_rng(Ci * 100 + Co + 1)
    B = 2
    x = rng.random((B, h, w, Ci), dtype=np.float32)
    wt = _rand(rng, 3, 3, Co, Ci, scale=(2.0 / (2.25 * Ci)) ** 0.5)
    b = _rand(rng, Co, scale=0.1)
    want = torch.relu(O.conv2dT_s2_same(_t64(x), _t64(wt), _t64(b))).numpy()
    got = gen_conv(CONVT_S2, 1, 0, split, PRE_RELU, x, wt, b, None, want.shape, out_mode=1)
    assert _err(got, want) < _tol(split), _err(got, want)
    got32 = gen_conv(CONVT_S2, 1, 0, split, PRE_RELU, x, wt, b, None, want.shape, out_mode=0)
    assert _err(got32, want) < _tol(split)


# ------------------------------------------------------------------------------------------ output layer (3x3 s1, flipped) + sigmoid
@pytest.mark.parametrize("Ci,Co,H,W", [(32, 3, 20, 45), (64, 3, 17, 64), (16, 1, 9, 30)])
def test_conv_s1_forward_sigmoid(Ci, Co, H, W):
    rng = np.random.default_rng(Ci + Co)
    x = rng.random((2, H, W, Ci), dtype=np.float32)
    wt = _rand(rng, 3, 3, Co, Ci, scale=(2.0 / (9 * Ci)) ** 0.5)
    b = _rand(rng, Co, scale=0.1)
    want = torch.sigmoid(O.conv2dT_s1_same(_t64(x), _t64(wt), _t64(b))).numpy()
    got = gen_conv(CONV_S1, 1, 1, 0, PRE_SIGMOID, x, wt, b, None, want.shape)
    assert float(np.abs(got - want).max()) < 4e-3


# ------------------------------------------------------------------------------------------ data gradients
@pytest.mark.parametrize("Ci,Co,H,W,mask_mode", [(32, 5, 24, 60, 2), (3, 32, 16, 40, 0), (64, 128, 16, 32, 2), (128, 32, 8, 28, 1)])
def test_conv_s2_data_gradient(Ci, Co, H, W, mask_mode):
    """d/d(input) of Conv2D s2 = Conv2DTranspose-type product over the PLAIN output gradient, masked by the ReLU of the layer
    input (stored space-to-depth in the model: mask_mode 2)."""
    rng = np.random.default_rng(7 * Ci + Co)
    B = 2
    x = rng.random((B, H, W, Ci), dtype=np.float32) - 0.3            # the "activation" whose ReLU mask applies
    w = _rand(rng, 3, 3, Ci, Co, scale=(2.0 / (9 * Ci)) ** 0.5)
    g = _rand(rng, B, H // 2, W // 2, Co)
    xt = _t64(x).requires_grad_(True)
    y = O.conv2d_s2_same(xt, _t64(w), torch.zeros(Co, dtype=torch.float64))
    (gx,) = torch.autograd.grad(y, xt, _t64(g))
    want = (gx * (xt.detach() > 0)).numpy()
    got = gen_conv(CONVT_S2, 1, 0, 0, PRE_NONE, g, w, None, x, want.shape, out_mode=1, mask_mode=mask_mode)
    assert _err(got, want) < _tol(0), _err(got, want)


@pytest.mark.parametrize("Ci,Co,h,w", [(32, 5, 12, 30), (64, 32, 8, 20), (128, 64, 6, 32)])
def test_convT_s2_data_gradient(Ci, Co, h, w):
    """d/d(input) of Conv2DTranspose s2 = stride-2 product over the space-to-depth output gradient."""
    rng = np.random.default_rng(11 * Ci + Co)
    B = 2
    x = rng.random((B, h, w, Ci), dtype=np.float32) - 0.3
    wt = _rand(rng, 3, 3, Co, Ci, scale=(2.0 / (2.25 * Ci)) ** 0.5)
    g = _rand(rng, B, 2 * h, 2 * w, Co)
    xt = _t64(x).requires_grad_(True)
    y = O.conv2dT_s2_same(xt, _t64(wt), torch.zeros(Co, dtype=torch.float64))
    (gx,) = torch.autograd.grad(y, xt, _t64(g))
    want = (gx * (xt.detach() > 0)).numpy()
    got = gen_conv(CONV_S2, 0, 0, 0, PRE_NONE, g, wt, None, x, want.shape, out_mode=0, mask_mode=1)
    assert _err(got, want) < _tol(0), _err(got, want)
    got_s2d = gen_conv(CONV_S2, 0, 0, 0, PRE_NONE, g, wt, None, x, want.shape, out_mode=2 if h % 2 == 0 and w % 2 == 0 else 1, mask_mode=0)
    assert _err(got_s2d, want) < _tol(0)


@pytest.mark.parametrize("Ci,Co,H,W", [(64, 3, 16, 40), (32, 3, 12, 30)])
def test_output_layer_data_gradient(Ci, Co, H, W):
    rng = np.random.default_rng(13 * Ci + Co)
    B = 2
    x = rng.random((B, H, W, Ci), dtype=np.float32) - 0.3
    wt = _rand(rng, 3, 3, Co, Ci, scale=(2.0 / (9 * Ci)) ** 0.5)
    g = _rand(rng, B, H, W, Co)
    xt = _t64(x).requires_grad_(True)
    y = O.conv2dT_s1_same(xt, _t64(wt), torch.zeros(Co, dtype=torch.float64))
    (gx,) = torch.autograd.grad(y, xt, _t64(g))
    want = (gx * (xt.detach() > 0)).numpy()
    got = gen_conv(CONV_S1, 0, 0, 0, PRE_NONE, g, wt, None, x, want.shape, out_mode=2, mask_mode=1)
    assert _err(got, want) < _tol(0), _err(got, want)


# ------------------------------------------------------------------------------------------ weight + bias gradients
@pytest.mark.parametrize("Ci,Co,H,W,x3", [(3, 32, 24, 60, 1), (3, 32, 24, 60, 2), (32, 5, 24, 50, 0), (3, 64, 16, 60, 2), (64, 128, 16, 60, 0), (128, 32, 12, 50, 0)])   # x3 = 2: X27 patch planes
def test_conv_s2_weight_gradient(Ci, Co, H, W, x3):
    rng = np.random.default_rng(17 * Ci + Co)
    B = 3
    x = rng.random((B, H, W, Ci), dtype=np.float32)
    g = _rand(rng, B, H // 2, W // 2, Co)
    wt = _t64(np.zeros((3, 3, Ci, Co))).requires_grad_(True)
    bt = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    y = O.conv2d_s2_same(_t64(x), wt, bt)
    gw, gb = torch.autograd.grad(y, (wt, bt), _t64(g))
    dW, db = gen_wgrad(CONV_S2, 0, 0, x, g, (3, 3, Ci, Co), Co, s_x3=x3)
    assert _err(dW, gw.numpy()) < _tol(0), _err(dW, gw.numpy())
    assert _err(db, gb.numpy()) < _tol(0)


@pytest.mark.parametrize("Ci,Co,h,w", [(32, 5, 12, 25), (64, 32, 8, 30), (32, 128, 8, 30), (128, 64, 6, 30)])
def test_convT_s2_weight_gradient(Ci, Co, h, w):
    rng = np.random.default_rng(19 * Ci + Co)
    B = 3
    x = rng.random((B, h, w, Ci), dtype=np.float32)
    g = _rand(rng, B, 2 * h, 2 * w, Co)
    wt = _t64(np.zeros((3, 3, Co, Ci))).requires_grad_(True)
    bt = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    y = O.conv2dT_s2_same(_t64(x), wt, bt)
    gw, gb = torch.autograd.grad(y, (wt, bt), _t64(g))
    dW, db = gen_wgrad(CONVT_S2, 1, 0, x, g, (3, 3, Co, Ci), Co)
    assert _err(dW, gw.numpy()) < _tol(0), _err(dW, gw.numpy())
    assert _err(db, gb.numpy()) < _tol(0)


@pytest.mark.parametrize("Ci,Co,H,W", [(64, 3, 12, 60), (32, 3, 10, 30)])
def test_output_layer_weight_gradient(Ci, Co, H, W):
    rng = np.random.default_rng(23 * Ci + Co)
    B = 2
    x = rng.random((B, H, W, Ci), dtype=np.float32)
    g = _rand(rng, B, H, W, Co)
    wt = _t64(np.zeros((3, 3, Co, Ci))).requires_grad_(True)
    bt = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    y = O.conv2dT_s1_same(_t64(x), wt, bt)
    gw, gb = torch.autograd.grad(y, (wt, bt), _t64(g))
    dW, db = gen_wgrad(CONV_S1, 1, 1, x, g, (3, 3, Co, Ci), Co)
    assert _err(dW, gw.numpy()) < _tol(0), _err(dW, gw.numpy())
    assert _err(db, gb.numpy()) < _tol(0)


# ------------------------------------------------------------------------------------------ Dense layers on the engine
def gen_dense(mode, split, relu, a, b, bias, out_shape):
    lib = _lib()
    ad, bd = _dev(a), _dev(b)
    biasd = _dev(bias) if bias is not None else None
    out = torch.full(out_shape, float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.gen_dense_test(mode, split, relu, _ptr(ad), _ptr(bd), _ptr(biasd), _ptr(out), a.shape[0], a.shape[1], b.shape[0] if mode != 1 else b.shape[1], None)
    lib.check(rc, None)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("K,N,B,split", [(32, 56 * 75 * 32, 16, 1), (32, 4096, 256, 0), (256, 8192, 37, 1), (8, 640, 3, 0)])
def test_dense_forward(K, N, B, split):
    """Decoder Dense (src/abstract_cvae.py:75-77): relu(z W + b) from the weight matrix's long dimension."""
    rng = np.random.default_rng(K + N + B)
    W = _rand(rng, K, N, scale=(1.0 / K) ** 0.5)
    z = _rand(rng, B, K)
    bias = _rand(rng, N, scale=0.1)
    want = np.maximum(z.astype(np.float64) @ W.astype(np.float64) + bias, 0.0)
    got = gen_dense(0, split, 1, W, z, bias, (B, N))
    assert _err(got, want) < _tol(split), _err(got, want)


@pytest.mark.parametrize("K,N,B", [(32, 56 * 75 * 32, 16), (32, 4096, 256), (192, 8192, 40), (8, 640, 3)])   # K + 1 <= 256 columns per product
def test_dense_weight_and_bias_gradient(K, N, B):
    rng = np.random.default_rng(K * 3 + N + B)
    G = _rand(rng, B, N)
    z = _rand(rng, B, K)
    want = np.concatenate([z.astype(np.float64).T @ G.astype(np.float64), G.astype(np.float64).sum(0, keepdims=True)], 0)   # [K + 1][N]
    got = gen_dense(1, 0, 0, G, z, None, (K + 1, N))
    assert _err(got, want) < _tol(0), _err(got, want)


@pytest.mark.parametrize("K,N,B", [(32, 56 * 75 * 32, 16), (32, 4096, 256), (256, 8192, 40), (8, 640, 3)])
def test_dense_data_gradient(K, N, B):
    rng = np.random.default_rng(K * 5 + N + B)
    G = _rand(rng, B, N)
    W = _rand(rng, K, N, scale=(1.0 / K) ** 0.5)
    want = G.astype(np.float64) @ W.astype(np.float64).T          # [B][K]
    got = gen_dense(2, 0, 0, G, W, None, (B, K))
    assert _err(got, want) < _tol(0), _err(got, want)
