"""Parity cases for the stages either side of the model, shared by the CPU run (g++ emulation
of the same kernel sources) and the GPU run (libkcvae.so).  Integer / byte outputs must be bit-exact."""
import numpy as np
import torch

from kcvae_testlib import emu_binding, make, pkg, small_config
from oracle import frontend_oracle as FO


def _binding(backend):
    return emu_binding() if backend == "emu" else None


def case_preprocess(backend, in_hw, cfg=None, B=2, seed=3):
    cfg = cfg or small_config()
    H, W, C = cfg["data"]["image_size"]
    m, _ = make(cfg, backend)
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, size=(B, in_hw[0], in_hw[1], C), dtype=np.uint8)
    got = m.preprocess_u8(frames).cpu().numpy()
    want = FO.resize_antialias(frames, H, W)
    assert got.shape == want.shape == (B, H, W, C)
    assert np.array_equal(got, want), f"max |diff| {np.abs(got - want).max()}"
    return m, frames, got


def case_preprocess_errors(backend):
    m, _ = make(small_config(), backend)
    import pytest
    with pytest.raises(ValueError):
        m.preprocess_u8(np.zeros((1, 8, 8, 4), np.uint8))        # wrong channel count
    with pytest.raises(ValueError):
        m.preprocess_u8(np.zeros((1, 8, 8, 3), np.float32))      # not uint8


def case_stream(backend, H=16, W=24, frames=12, seed=5, ma=0.9):
    rng = np.random.default_rng(seed)
    s = pkg.StreamingAnomalyScore(H, W, stream_error_ma=ma, binding=_binding(backend))
    o = FO.StreamScoreOracle(stream_error_ma=ma)
    for t in range(frames):
        err = (rng.random((H, W), dtype=np.float32) ** 2 * 0.3).astype(np.float32)
        if t % 4 == 3:                                           # a planted bright patch
            err[2:5, 3:7] += np.float32(1.5 + 0.1 * t)
        g, w = s.update(err), o.update(err)
        # pixels whose z-of-z sits within float rounding of the threshold may fall either side
        border = int(np.sum(np.abs(w["zz"] - 3.0) < 1e-4))
        assert abs(g["anomaly_count"] - w["anomaly_count"]) <= border, (t, g["anomaly_count"], w["anomaly_count"])
        for k in ("frame_min", "frame_max", "stream_error_min", "stream_error_max"):
            assert g[k] == w[k], (t, k, g[k], w[k])
        for k in ("z_mean", "z_std"):
            assert abs(g[k] - w[k]) <= 1e-5 * max(1.0, abs(w[k])), (t, k, g[k], w[k])
        if border == 0:
            a, b = g["anomaly_score"], w["anomaly_score"]
            assert (np.isnan(a) and np.isnan(b)) or a == b or abs(a - b) <= 1e-5 * max(1.0, abs(b)), (t, a, b)
        assert np.array_equal(g["stream_error_img"].cpu().numpy(), w["stream_error_img"]), t
    s.reset()
    o2 = FO.StreamScoreOracle(stream_error_ma=ma)
    err = rng.random((H, W), dtype=np.float32)
    assert s.update(err)["stream_error_max"] == o2.update(err)["stream_error_max"]


def case_render(backend, B=2, H=16, W=24, seed=9):
    rng = np.random.default_rng(seed)
    norm = rng.random((B, H, W), dtype=np.float32)
    norm.reshape(-1)[:256] = (np.arange(256, dtype=np.float32) / np.float32(255.0))      # every table entry
    norm.reshape(-1)[256:260] = [-0.2, 1.3, 0.5 / 255, 1.5 / 255]                         # saturation, ties
    rec = rng.random((B, H, W, 3), dtype=np.float32)
    got = pkg.render_outputs(norm, rec, binding=_binding(backend))
    want = FO.render_outputs(norm, rec)
    for k in ("err", "heatmap", "overlay", "rec"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    only = pkg.render_outputs(norm_err=norm, binding=_binding(backend))
    assert only["overlay"] is None and np.array_equal(only["heatmap"].cpu().numpy(), want["heatmap"])
