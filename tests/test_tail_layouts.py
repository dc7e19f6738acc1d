"""CPU models of the operand layouts of the specialised decoder-tail kernels (csrc/tc_conv.cu), test infrastructure only.

Three kernels fold convolution taps into the M or N dimension of ONE tcgen05.mma per K step by the way their shared-memory
operands are laid out.  The index algebra of those layouts (which unit a descriptor group reads, which accumulator row /
column is which tap, how the epilogue maps them back to Keras' [kh, kw, co, ci]) is restated here with numpy on one tile
and checked against autograd of the oracle's layer functions, so that it is pinned on machines without a GPU.  The kernels
themselves are checked on the B200 by tests/test_gpu_parity.py (test_tc_backward_parity, test_tc_output_conv_parity).
"""
import numpy as np
import torch

from oracle import kcvae_oracle as O

PW, TW, TR = 32, 30, 32          # halo pitch, tile columns, tile rows (tc_conv.cu)


def _rand(rng, *shape):
    return rng.standard_normal(shape)


def test_out_layer_weight_gradient_copies():
    """tc_out_wgrad_kernel: A = three column-shifted, row-interleaved copies of the dl tile, unit ((R*3 + kw)*32 + c) =
    dl[R - 2][c - 2 + kw]; M-group G = kh*3 + kw of the descriptor starting at halo pixel (r, c) reads unit
    ((r + kh)*3 + kw)*32 + c; B = the halo tile of the activation; D[G*8 + co][ci] is dW[kh, kw, co, ci]."""
    rng = np.random.default_rng(1)
    Cin, Cout = 32, 3
    a = _rand(rng, 1, TR, TW, Cin)                     # one image = one tile
    dl = _rand(rng, 1, TR, TW, Cout)
    w = torch.zeros(3, 3, Cout, Cin, dtype=torch.float64, requires_grad=True)
    y = O.conv2dT_s1_same(torch.from_numpy(a), w, torch.zeros(Cout, dtype=torch.float64))
    (y * torch.from_numpy(dl)).sum().backward()
    want = w.grad.numpy()                              # [kh, kw, co, ci]

    halo = np.zeros((TR + 2, PW, Cin))                 # TMA box at (-1, -1), zero fill outside the image
    halo[1:TR + 1, 1:TW + 1] = a[0]
    copies = np.zeros(((TR + 4) * 3 * PW, 8))          # 16-byte units of 8 channels
    for rho in range(TR):
        for c in range(TW):
            for kw in range(3):
                copies[((rho + 2) * 3 + kw) * PW + c + 2 - kw, :Cout] = dl[0, rho, c]
    D = np.zeros((128, Cin))
    for r in range(TR + 2):                            # K steps: 16 consecutive pixels of a halo row
        for c in range(PW):
            start = r * 3 * PW + c
            for G in range(9):
                D[G * 8:G * 8 + 8] += np.outer(copies[start + G * PW], halo[r, c])
    got = D[:72].reshape(3, 3, 8, Cin)[:, :, :Cout]
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


def test_last_convT_weight_gradient_copies():
    """tc_convT_wgrad_kernel: A = two copies of the activation plane, unit ((R*2 + dh)*32 + c) = a[R - 1][c - dh], M-group
    2*g + dh; B = the space-to-depth gradient tile, N = parity*32 + co; the epilogue keeps (parity, g, dh) combinations
    that are taps: kh = 2 / 0 (row parity 0, g = 0 / 1) or 1 (row parity 1, g = 1), kw = 2*dh (column parity 0) or 1."""
    rng = np.random.default_rng(2)
    TRD, Cin, Cout = 8, 5, 32
    a = _rand(rng, 1, TRD, TW, Cin)
    g_out = _rand(rng, 1, 2 * TRD, 2 * TW, Cout)
    w = torch.zeros(3, 3, Cout, Cin, dtype=torch.float64, requires_grad=True)
    y = O.conv2dT_s2_same(torch.from_numpy(a), w, torch.zeros(Cout, dtype=torch.float64))
    (y * torch.from_numpy(g_out)).sum().backward()
    want = w.grad.numpy()                              # [kh, kw, co, ci]

    GROWS = TRD + 1
    gs2d = np.zeros((4, GROWS, PW, Cout))              # plane block = parity, TMA zero fill past the image
    for pa in range(2):
        for pb in range(2):
            gs2d[pa * 2 + pb, :TRD, :TW] = g_out[0, pa::2, pb::2]
    copies = np.zeros(((TRD + 2) * 2 * PW + 8 * PW, 8))
    for rho in range(TRD):
        for c in range(TW):
            copies[((rho + 1) * 2) * PW + c, :Cin] = a[0, rho, c]
            copies[((rho + 1) * 2 + 1) * PW + c + 1, :Cin] = a[0, rho, c]
    D = np.zeros((64, 4 * Cout))
    for gr in range(GROWS):
        for gc in range(PW):
            start = gr * 2 * PW + gc
            b = gs2d[:, gr, gc].reshape(-1)            # N = parity*32 + co
            for Gm in range(4):
                D[Gm * 8:Gm * 8 + 8] += np.outer(copies[start + Gm * PW], b)
    got = np.zeros_like(want)
    seen = np.zeros((3, 3), int)
    for g in range(2):
        for dh in range(2):
            for par in range(4):
                pa, pb = par >> 1, par & 1
                if not ((pa == 0 or g == 1) and (pb == 0 or dh == 0)):
                    continue
                kh = (2 if g == 0 else 0) if pa == 0 else 1
                kw = 2 * dh if pb == 0 else 1
                seen[kh, kw] += 1
                rows = (g * 2 + dh) * 8
                got[kh, kw] = D[rows:rows + Cin, par * Cout:(par + 1) * Cout].T
    assert (seen == 1).all()                           # every tap exactly once
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


def test_fused_tail_horizontal_taps_in_n():
    """tc_tail_fused_kernel phase B: T[q][kw*8 + co] = sum over kh, ci of halo[q + (2 - kh) rows][ci] W[kh][kw][co][ci] (no
    column shift in the operand); output pixel c of a halo row = T[c][tap 2] + T[c + 1][tap 1] + T[c + 2][tap 0], the two
    lane shuffles of the epilogue."""
    rng = np.random.default_rng(3)
    Cin, Cout = 32, 3
    a = _rand(rng, 1, TR, TW, Cin)
    w = _rand(rng, 3, 3, Cout, Cin)
    want = O.conv2dT_s1_same(torch.from_numpy(a), torch.from_numpy(w), torch.zeros(Cout, dtype=torch.float64)).numpy()[0]

    halo = np.zeros((TR + 2 + 1, PW, Cin))             # + 1 row: the last M-tile's reads stay in range
    halo[1:TR + 1, 1:TW + 1] = a[0]
    flat = halo.reshape(-1, Cin)
    T = np.zeros((TR * PW, 32))
    for q in range(TR * PW):
        for kh in range(3):
            for kw in range(3):
                T[q, kw * 8:kw * 8 + Cout] += w[kh, kw] @ flat[q + (2 - kh) * PW]
    Tr = T.reshape(TR, PW, 32)
    got = np.zeros((TR, TW, Cout))
    for c in range(TW):
        got[:, c] = Tr[:, c, 16:16 + Cout] + Tr[:, c + 1, 8:8 + Cout] + Tr[:, c + 2, 0:Cout]
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)
