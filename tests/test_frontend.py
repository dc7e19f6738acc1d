"""SURVEY 8f rows 2-4 on CPU: the oracle against OpenCV's own golden vectors and definitional properties, and the
kernel sources (g++ emulation build) against the oracle.  The GPU twins live in tests/test_gpu_frontend.py."""
import os

import numpy as np
import pytest

import frontend_cases as FC
from kcvae_testlib import ROOT
from oracle import frontend_oracle as FO

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "cv2_scorer_outputs.npz"))


def test_oracle_overlay_rounding_matches_cv2_addweighted_exhaustively():
    a = np.repeat(np.arange(256)[:, None], 256, axis=1)
    b = np.repeat(np.arange(256)[None, :], 256, axis=0)
    mine = np.clip(np.rint(0.5 * a + 0.5 * b), 0, 255).astype(np.uint8)
    assert np.array_equal(mine, GOLD["add_weighted"])


def test_jet_table_in_the_kernel_source_is_the_cv2_table():
    inc = open(os.path.join(ROOT, "trustedai-cl-vae-ad_b200", "csrc", "jet_lut.inc")).read()
    import re
    rows = np.array([[int(v) for v in m] for m in re.findall(r"\{(\d+), (\d+), (\d+)\}", inc)], np.uint8)
    assert rows.shape == (256, 3) and np.array_equal(rows, GOLD["jet_lut"])
    try:
        import cv2
    except ImportError:
        return
    live = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET).reshape(256, 3)
    assert np.array_equal(live, GOLD["jet_lut"])


def test_oracle_render_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    norm, rec = rng.random((1, 20, 30), dtype=np.float32), rng.random((1, 20, 30, 3), dtype=np.float32)
    o = FO.render_outputs(norm, rec)
    err = np.round(255. * norm[0]).astype(np.uint8)                         # do_anomaly_detection.py:166
    heat = cv2.applyColorMap(err, cv2.COLORMAP_JET)                         # :167
    over = cv2.addWeighted(heat, 0.5, np.round(255. * rec[0]).astype(np.uint8), 0.5, 0.0)   # :168
    assert np.array_equal(o["err"][0], err) and np.array_equal(o["heatmap"][0], heat) and np.array_equal(o["overlay"][0], over)


def test_resize_oracle_definitional_properties():
    rng = np.random.default_rng(1)
    f = rng.integers(0, 256, size=(1, 8, 12, 3), dtype=np.uint8)
    assert np.array_equal(FO.resize_antialias(f, 8, 12), f.astype(np.float32) / np.float32(255))     # identity
    c = np.full((1, 9, 14, 3), 200, np.uint8)
    for hw in ((4, 7), (18, 28), (5, 5)):                                                             # constants survive
        assert np.allclose(FO.resize_antialias(c, *hw), 200 / 255, atol=2e-7)
    # 2:1 shrink with the widened triangle: interior weights (1, 3, 3, 1) / 8
    g = rng.integers(0, 256, size=(1, 16, 2, 1), dtype=np.uint8)
    out = FO.resize_antialias(g, 8, 2)
    x = g.astype(np.float64)[0, :, :, 0] / 255
    want = (x[1] + 3 * x[2] + 3 * x[3] + x[4]) / 8
    assert np.allclose(out[0, 1, :, 0], want, atol=1e-6)
    # 1:2 enlargement is plain half-pixel bilinear: 0.75 / 0.25 blends
    u = FO.resize_antialias(g, 32, 2)
    assert np.allclose(u[0, 3, :, 0], 0.75 * x[1] + 0.25 * x[2], atol=1e-6)


@pytest.mark.parametrize("in_hw", [(16, 24), (37, 53), (8, 12), (16, 50), (33, 24)])
def test_emu_preprocess(in_hw):
    FC.case_preprocess("emu", in_hw)


def test_emu_preprocess_errors():
    FC.case_preprocess_errors("emu")


def test_emu_stream_score():
    FC.case_stream("emu")
    FC.case_stream("emu", H=7, W=9, frames=5, ma=0.99)


def test_emu_render_outputs():
    FC.case_render("emu")


def test_device_data_queue_follows_the_camera_tools_dataqueue():
    """camera_streamer_qt.py:61-81 semantics (restated: capacity copies of the first sample, append advances the
    index first) and the :1342 stacking with the replay buffer, through one continual-learning step."""
    from kcvae_testlib import make, pkg, small_config
    cfg = small_config()
    H, W, C = cfg["data"]["image_size"]
    rng = np.random.default_rng(4)
    frames = rng.random((6, H, W, C), dtype=np.float32)
    replay = rng.random((2, H, W, C), dtype=np.float32)
    q = pkg.DeviceDataQueue(frames[0], 3, replay_buffer=replay)
    ref = [frames[0].copy() for _ in range(3)]
    idx = 0
    for f in frames[1:]:
        q.append(f)
        idx = (idx + 1) % 3
        ref[idx] = f
        assert q._idx == idx and np.array_equal(q.get().numpy(), f)
    assert np.array_equal(q.to_numpy(), np.array(ref))
    assert np.array_equal(q.stacked().numpy(), np.vstack((np.array(ref), replay)))
    m, _ = make(cfg, "emu")
    m.compile(optimizer=pkg.Adam(learning_rate=1e-3))
    loss, r_img = m.train_step_and_run(q.stacked())            # :1345
    assert r_img[q._idx].shape == (H, W, C) and np.isfinite(float(loss["loss"]))   # :1347
