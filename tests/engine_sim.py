"""CPU interpreter of the general tensor-core engine's host-built plans (csrc/tc_gen.cu), test infrastructure only.

`kcvae_gen_plan_dump` (needs no GPU) hands over what the kernels execute: the K slabs with their chunk planes, the MMA
list (descriptor low words, instruction descriptor, accumulator column), the weight gather table; for weight gradients
the accumulator roles, their MMA list and the scatter table.  This module runs such a plan with numpy exactly the way
the hardware reads it - shared-memory stages as arrays of 16-byte units, the no-swizzle K-major operand
"unit(row r, chunk c) = start + c * LBO + r" and its MN-major twin "unit(pixel k, 8-channel group g) = start + g * SBO + k" -
so that the planner's index arithmetic (tap shifts, space-to-depth element maps, parity pruning, hi + lo triples, role
packing, scatter) is checked against the oracle's layer functions on machines without a GPU.

Flat int32 layouts (gen_conv_plan_dump / gen_wgrad_plan_dump):
  conv : 1, n_groups, MT, NB, acc_cols, R_in, row0, col0, TW, in_PL, n_slabs, n_mma, type, CHb, a_region, stage_bytes, Cop,
         2 x (slab0, slab_n, a_par), n_slabs x (mma0, mma_n, nplanes, b_src, b_bytes, plane[32]), n_mma x (a_lo, b_lo, idesc, dcol_acc),
         table_len, table...
  wgrad: 2, n_roles, TRr, R_s, row0, col0, TW, s_PL, u_PL, CHs, CHu, s_region, u_region, stage_bytes, ones_off, EW, Cu,
         n_roles x (mma0, mma_n, nS, nU, ncols, s_plane[64], u_plane[64]), n_mma, n_mma x (a_lo, a_hi, b_lo, b_hi, idesc, d_col, b_ones),
         src_len, src...
"""
import ctypes as C
import importlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GP = 32
LO_FLAG = 0x40000000
PLAIN, S2D, X3, X27 = 0, 1, 2, 3
CONV_S2, CONVT_S2, CONV_S1, DENSE = 0, 1, 2, 3


def _cdll():
    build = importlib.import_module("trustedai-cl-vae-ad_b200.build")
    lib = C.CDLL(build.build())
    lib.kcvae_gen_plan_dump.restype = C.c_int64
    lib.kcvae_gen_plan_dump.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
    lib.kcvae_last_error.restype = C.c_char_p
    lib.kcvae_last_error.argtypes = [C.c_void_p]
    return lib


def dump(which, spec):
    lib = _cdll()
    sp = np.asarray(spec, np.int32)
    n = lib.kcvae_gen_plan_dump(which, sp.ctypes.data_as(C.c_void_p), len(sp), None, 0)
    if n < 0:
        raise RuntimeError((lib.kcvae_last_error(None) or b"").decode())
    out = np.zeros(n, np.int32)
    assert lib.kcvae_gen_plan_dump(which, sp.ctypes.data_as(C.c_void_p), len(sp), out.ctypes.data_as(C.c_void_p), n) == n
    return out


def bf16(a):
    """round to bf16 (nearest even), back to float64"""
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.bfloat16).to(torch.float64).numpy()


def hi_lo(a):
    hi = bf16(a)
    return hi, bf16(np.asarray(a, np.float64) - hi)


# ------------------------------------------------------------------------------------------ plane tensors
def planes_count(layout, KC, split):
    return (4 * KC if layout == S2D else (2 if layout == X3 else (4 if layout == X27 else KC))) * (2 if split else 1)


def pack_planes(x, layout, KC, split):
    """x [B, H, W, C] (full resolution) -> [B, PL, Hp, Wp, 8] the way gen_pack_nhwc / gen_pack_x3 lay it out"""
    B, H, W, Cc = x.shape
    hi, lo = hi_lo(x)
    if layout == PLAIN:
        out = np.zeros((B, planes_count(layout, KC, split), H, W, 8))
        for c in range(Cc):
            out[:, c // 8, :, :, c % 8] = hi[..., c]
            if split:
                out[:, KC + c // 8, :, :, c % 8] = lo[..., c]
        return out
    Hp, Wp = H // 2, W // 2
    out = np.zeros((B, planes_count(layout, KC, split), Hp, Wp, 8))
    if layout == X27:                      # element (kh*3 + kw)*3 + c of pixel (i, j) = x[2i + kh][2j + kw][c], zero outside the image
        pad = np.zeros((B, H + 2, W + 2, Cc))
        pad[:, :H, :W] = hi
        for kh in range(3):
            for kw in range(3):
                for c in range(Cc):
                    e = (kh * 3 + kw) * 3 + c
                    out[:, e // 8, :, :, e % 8] = pad[:, kh:kh + H:2, kw:kw + W:2, c]
        return out
    for a in range(2):
        for b in range(2):
            par = a * 2 + b
            for c in range(Cc):
                e = par * 3 + c if layout == X3 else None
                plane = e // 8 if layout == X3 else par * KC + c // 8
                slot = e % 8 if layout == X3 else c % 8
                out[:, plane, :, :, slot] = hi[:, a::2, b::2, c]
                if split:
                    out[:, (2 if layout == X3 else 4 * KC) + plane, :, :, slot] = lo[:, a::2, b::2, c]
    return out


# ------------------------------------------------------------------------------------------ forward-type plans
class ConvPlan:
    def __init__(self, v):
        assert v[0] == 1
        (self.n_groups, self.MT, self.NB, self.acc_cols, self.R_in, self.row0, self.col0, self.TW, self.in_PL, self.n_slabs,
         self.n_mma, self.type, self.CHb, self.a_region, self.stage_bytes, self.Cop) = [int(t) for t in v[1:17]]
        p = 17
        self.groups = [tuple(int(t) for t in v[p + 3 * g:p + 3 * g + 3]) for g in range(2)]
        p += 6
        self.slabs = []
        for _ in range(self.n_slabs):
            mma0, mma_n, npl, b_src, b_bytes = [int(t) for t in v[p:p + 5]]
            self.slabs.append((mma0, mma_n, npl, b_src, b_bytes, [int(t) for t in v[p + 5:p + 5 + npl]]))
            p += 5 + 32
        self.mma = [tuple(int(np.uint32(t)) for t in v[p + 4 * i:p + 4 * i + 4]) for i in range(self.n_mma)]
        p += 4 * self.n_mma
        n = int(v[p])
        self.table = v[p + 1:p + 1 + n].astype(np.int64)


def weight_image(table, w_flat):
    """gen_gather_weights_kernel: bf16 image elements from the fp32 source"""
    w = np.asarray(w_flat, np.float64).ravel()
    hi, lo = hi_lo(w)
    img = np.zeros(len(table))
    ok = table >= 0
    idx = table[ok] & (LO_FLAG - 1)
    img[ok] = np.where((table[ok] & LO_FLAG) != 0, lo[idx], hi[idx])
    return img


def run_conv(plan, planes, w_flat, Hg, Wg):
    """accumulators of every GEMM pixel: returns D [B, n_groups, Hg, Wg, acc_cols] (what the epilogue warps read from TMEM)"""
    B, PL, Hp, Wp, _ = planes.shape
    img = weight_image(plan.table, w_flat)
    TRr = 4 * plan.MT
    out = np.zeros((B, plan.n_groups, Hg, Wg, plan.acc_cols))
    units = plan.stage_bytes // 16
    for n in range(B):
        for ty in range(-(-Hg // TRr)):
            for tx in range(-(-Wg // plan.TW)):
                for g in range(plan.n_groups):
                    slab0, slab_n, _ = plan.groups[g]
                    acc = np.zeros((plan.MT, 128, plan.acc_cols))
                    touched = np.zeros(plan.acc_cols, bool)
                    for sl in range(slab0, slab0 + slab_n):
                        mma0, mma_n, npl, b_src, b_bytes, pls = plan.slabs[sl]
                        stage = np.zeros((units + 64, 8))
                        for pos, pl in enumerate(pls):                      # one TMA box per plane, zero fill outside the tensor
                            for r in range(plan.R_in):
                                y = ty * TRr + plan.row0 + r
                                if not 0 <= y < Hp:
                                    continue
                                x0 = tx * plan.TW + plan.col0
                                lo_, hi_ = max(0, -x0), min(GP, Wp - x0)
                                if hi_ > lo_:
                                    u0 = (pos * plan.CHb) // 16 + r * GP
                                    stage[u0 + lo_:u0 + hi_] = planes[n, pl, y, x0 + lo_:x0 + hi_]
                        stage[plan.a_region // 16:plan.a_region // 16 + b_bytes // 16] = img[b_src // 2:(b_src + b_bytes) // 2].reshape(-1, 8)
                        for i in range(mma0, mma0 + mma_n):
                            a_lo, b_lo, idesc, dcol_acc = plan.mma[i]
                            a_off, a_lbo = a_lo & 0x3FFF, (a_lo >> 16) & 0x3FFF
                            b_off, b_lbo = b_lo & 0x3FFF, (b_lo >> 16) & 0x3FFF
                            N = ((idesc >> 17) & 0x3F) << 3
                            assert ((idesc >> 24) & 0x1F) << 4 == 128 and b_lbo == N
                            dcol, accf = dcol_acc & 0x7FFFFFFF, dcol_acc >> 31
                            if not accf:
                                assert not touched[dcol:dcol + N].any(), "an accumulator column is overwritten after it was accumulated into"
                            else:
                                assert touched[dcol:dcol + N].all(), "accumulate into a column nothing initialised"
                            for mt in range(plan.MT):
                                prod = 0.0
                                for c in range(2):
                                    A = stage[a_off + mt * 128 + c * a_lbo + np.arange(128)]
                                    Bm = stage[b_off + c * b_lbo + np.arange(N)]
                                    prod = prod + A @ Bm.T
                                if accf:
                                    acc[mt, :, dcol:dcol + N] += prod
                                else:
                                    acc[mt, :, dcol:dcol + N] = prod
                            touched[dcol:dcol + N] = True
                    for mt in range(plan.MT):
                        for m in range(128):
                            q = mt * 128 + m
                            gy, gx = ty * TRr + q // GP, tx * plan.TW + q % GP
                            if q % GP < plan.TW and gy < Hg and gx < Wg:
                                out[n, g, gy, gx] = acc[mt, m]
    return out


def conv_output(plan, D, Cn):
    """accumulators -> the layer's pre-activation output [B, Ho, Wo, Cn] (the epilogue's column / parity mapping)"""
    B, G, Hg, Wg, _ = D.shape
    if plan.type == CONVT_S2:
        out = np.zeros((B, 2 * Hg, 2 * Wg, Cn))
        for a in range(2):
            for b in range(2):
                out[:, a::2, b::2, :] = D[:, a, :, :, b * plan.Cop:b * plan.Cop + Cn]
        return out
    return D[:, 0, :, :, :Cn]


# ------------------------------------------------------------------------------------------ weight-gradient plans
class WgradPlan:
    def __init__(self, v):
        assert v[0] == 2
        (self.n_roles, self.TRr, self.R_s, self.row0, self.col0, self.TW, self.s_PL, self.u_PL, self.CHs, self.CHu, self.s_region,
         self.u_region, self.stage_bytes, self.ones_off, self.EW, self.Cu) = [int(t) for t in v[1:17]]
        p = 17
        self.roles = []
        for _ in range(self.n_roles):
            mma0, mma_n, nS, nU, ncols = [int(t) for t in v[p:p + 5]]
            self.roles.append((mma0, mma_n, nS, nU, ncols, [int(t) for t in v[p + 5:p + 5 + nS]], [int(t) for t in v[p + 69:p + 69 + nU]]))
            p += 5 + 128
        n_mma = int(v[p]); p += 1
        self.mma = [tuple(int(np.uint32(t)) for t in v[p + 7 * i:p + 7 * i + 7]) for i in range(n_mma)]
        p += 7 * n_mma
        n = int(v[p])
        self.src = v[p + 1:p + 1 + n].astype(np.int64).reshape(-1, 4)


def run_wgrad(plan, S, U, Hg, Wg):
    """S [B, PLs, Hs, Ws, 8] (shifted operand), U [B, PLu, Hg, Wg, 8] (gradient).  Returns (dW flat [EW], db [Cu])."""
    B = S.shape[0]
    assert Wg % plan.TW == 0
    tmem = np.zeros((plan.n_roles, 128, 512))
    ones = np.ones((2 * plan.CHu // 16, 8))
    for role, (mma0, mma_n, nS, nU, ncols, s_pl, u_pl) in enumerate(plan.roles):
        for n in range(B):
            for ty in range(-(-Hg // plan.TRr)):
                for tx in range(Wg // plan.TW):
                    stage = np.zeros((plan.stage_bytes // 16 + 4096, 8))
                    for pos, pl in enumerate(s_pl):
                        for r in range(plan.R_s):
                            y = ty * plan.TRr + plan.row0 + r
                            if not 0 <= y < S.shape[2]:
                                continue
                            x0 = tx * plan.TW + plan.col0
                            lo_, hi_ = max(0, -x0), min(GP, S.shape[3] - x0)
                            if hi_ > lo_:
                                u0 = (pos * plan.CHs) // 16 + r * GP
                                stage[u0 + lo_:u0 + hi_] = S[n, pl, y, x0 + lo_:x0 + hi_]
                    for pos, pl in enumerate(u_pl):                         # windowed map: only the tile's TW columns are non-zero
                        for r in range(plan.TRr):
                            y = ty * plan.TRr + r
                            if y < Hg:
                                u0 = (plan.s_region + pos * plan.CHu) // 16 + r * GP
                                stage[u0:u0 + plan.TW] = U[n, pl, y, tx * plan.TW:(tx + 1) * plan.TW]
                    for i in range(mma0, mma0 + mma_n):
                        a_lo, a_hi, b_lo, b_hi, idesc, d_col, b_ones = plan.mma[i]
                        a_off, b_off = a_lo & 0x3FFF, b_lo & 0x3FFF
                        assert (a_lo >> 16) & 0x3FFF == 8 and (b_lo >> 16) & 0x3FFF == 8       # 8 pixels per K group
                        a_sbo, b_sbo = a_hi & 0x3FFF, b_hi & 0x3FFF
                        M, N = ((idesc >> 24) & 0x1F) << 4, ((idesc >> 17) & 0x3F) << 3
                        assert (idesc >> 15) & 1 and (idesc >> 16) & 1                          # both operands MN-major
                        K = plan.TRr * GP
                        A = np.concatenate([stage[a_off + g * a_sbo + np.arange(K)] for g in range(M // 8)], axis=1)        # [K, M]
                        Bsrc = ones if b_ones else stage
                        Bm = np.concatenate([Bsrc[b_off + g * b_sbo + np.arange(K)] for g in range(N // 8)], axis=1)        # [K, N]
                        prod = A.T @ Bm                                                                                     # [M, N]
                        lanes = np.arange(M) if M == 128 else (np.arange(M) // 16) * 32 + np.arange(M) % 16
                        tmem[role][lanes[:, None], d_col + np.arange(N)[None, :]] += prod
    flat = tmem.transpose(0, 2, 1).reshape(plan.n_roles * 512 * 128)          # partial index = role*(512*128) + col*128 + lane
    out = np.zeros(plan.src.shape[0])
    for k in range(4):
        ok = plan.src[:, k] >= 0
        out[ok] += flat[plan.src[ok, k]]
    return out[:plan.EW], out[plan.EW:plan.EW + plan.Cu]
