"""Shared helpers for the parity tests.  ``backend`` is 'cuda' (libkcvae.so on a B200, the
product) or 'emu' (tests/emu g++ functional simulation of the same kernel sources: checks
kernel index math / reductions / orchestration on GPU-less machines; never shipped)."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import kcvae_oracle as O  # noqa: E402

pkg = importlib.import_module("trustedai-cl-vae-ad_b200")
_lib = importlib.import_module("trustedai-cl-vae-ad_b200._lib")
_build = importlib.import_module("trustedai-cl-vae-ad_b200.build")

_EMU = None


def emu_binding():
    global _EMU
    if _EMU is None:
        path = _build.build_emu(sanitize=bool(os.environ.get("KCVAE_EMU_ASAN")))   # ASan: also LD_PRELOAD libasan
        _EMU = _lib.Binding(C.CDLL(path), "cpu", path)
    return _EMU


def model_class(backend: str, kind: str = "global"):
    base = pkg.KurtosisGlobalCVAE if kind == "global" else pkg.KurtosisSingleCVAE
    if backend == "emu":
        return type("Emu" + base.__name__, (base,), {"_binding_override": emu_binding()})
    return base


def small_config(kind="global", H=16, W=24, C_=3, layers=(6, 5), enc=7, dec=4, latent=5, w_skew=0.05):
    cfg = {
        "data": {"image_size": [H, W, C_]},
        "loss": {"kurtosis": 3.0, "w_kl_divergence": 0.0, "w_kurtosis": 1e-2, "w_mse": 1.0,
                 "w_skew": w_skew, "w_x_std": 1e-10, "w_z_l1_reg": 1e-2},
        "model": {"decoder_dense_filters": dec, "latent_dimensions": latent, "layers": list(layers)},
        "training": {"batch_size": 4, "beta": 1e-2, "learning_rate": 1e-3, "max_epochs": 1},
    }
    if enc:
        cfg["model"]["encoder_dense_filters"] = enc
    if kind == "single":
        cfg["model"]["type"] = "KurtosisSingle"
    return cfg


def make(cfg, backend, seed=1234, bias_scale=0.05, weight_gain=1.0, **kw):
    kind = "single" if cfg["model"].get("type") == "KurtosisSingle" else "global"
    kw.setdefault("precision", "fp32")     # the library default is the tensor-core path; the tight fp32 bars are asked for
    m = model_class(backend, kind)(cfg, **kw)
    ws = O.glorot_init(cfg, seed, bias_scale=bias_scale)
    if weight_gain != 1.0:
        ws = [w * weight_gain if w.ndim > 1 else w for w in ws]
    m.set_weights(ws)
    return m, ws


def frames(cfg, B, seed=42):
    return O.synthetic_frames(B, cfg, seed)


def eps_for(cfg, B, step=0):
    return O.synthetic_eps(B, cfg, step)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def assert_metrics_close(got: dict, want: dict, rtol=1e-4, atol=1e-6, skip=()):
    assert list(got.keys()) == list(want.keys())
    for k in want:
        if k in skip:
            continue
        g, w = float(got[k]), float(want[k])
        if g == w:      # also covers matching infinities
            continue
        assert abs(g - w) <= atol + rtol * abs(w), f"{k}: got {g} want {w}"
