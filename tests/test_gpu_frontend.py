"""SURVEY 8f rows 2-4 on the B200: libkcvae.so against the oracle (bit-exact bytes), the uint8 host entry points
against the fp32 ones, and the README shape."""
import numpy as np
import pytest
import torch

import frontend_cases as FC
from kcvae_testlib import O, make, pkg
from oracle import frontend_oracle as FO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("in_hw", [(16, 24), (37, 53), (8, 12), (16, 50)])
def test_preprocess(in_hw):
    FC.case_preprocess("cuda", in_hw)


def test_preprocess_errors():
    FC.case_preprocess_errors("cuda")


def test_stream_score():
    FC.case_stream("cuda")
    FC.case_stream("cuda", H=7, W=9, frames=5, ma=0.99)


def test_render_outputs():
    FC.case_render("cuda")


def test_readme_shape_camera_frames_end_to_end():
    """480x640 camera frames -> antialiased resize -> score -> streaming score -> rendered outputs."""
    cfg = O.readme_config()
    m, _ = make(cfg, "cuda", precision="bf16")
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, size=(3, 480, 640, 3), dtype=np.uint8)
    x = m.preprocess_u8(frames)
    want = FO.resize_antialias(frames[:1], 224, 300)
    assert np.array_equal(x[:1].cpu().numpy(), want)
    r = m.score(x, return_err=True, return_rec=True)
    # the uint8 host entry point gives the same scores as scoring the preprocessed frames
    sc = m.score_host_u8(torch.from_numpy(frames))
    assert np.array_equal(sc.numpy(), r["score"].cpu().numpy())
    s = pkg.StreamingAnomalyScore(224, 300)
    o = FO.StreamScoreOracle()
    for i in range(3):
        e = r["err"][i]
        g, w = s.update(e), o.update(e.cpu().numpy())
        assert np.array_equal(g["stream_error_img"].cpu().numpy(), w["stream_error_img"])
        assert abs(g["anomaly_count"] - w["anomaly_count"]) <= int(np.sum(np.abs(w["zz"] - 3.0) < 1e-4))
    norm = (r["err"] - r["err"].min()) / (r["err"].max() - r["err"].min())
    got = pkg.render_outputs(norm, r["rec"])
    wantr = FO.render_outputs(norm.cpu().numpy(), r["rec"].cpu().numpy())
    for k in ("err", "heatmap", "overlay", "rec"):
        assert np.array_equal(got[k].cpu().numpy(), wantr[k]), k


def test_u8_host_training_matches_float_host_training():
    cfg = O.readme_config()
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, size=(4, 224, 300, 3), dtype=np.uint8)
    eps = torch.from_numpy(O.synthetic_eps(4, cfg))
    outs = []
    for mode in ("u8", "f32"):
        m, _ = make(cfg, "cuda", precision="bf16")
        m.compile(optimizer=pkg.Adam(learning_rate=1e-3))
        if mode == "u8":
            outs.append(m.train_step_host_u8(torch.from_numpy(frames), eps).numpy().copy())
        else:
            x = torch.from_numpy(frames.astype(np.float32) / np.float32(255))
            outs.append(m.train_step_host(x, eps).numpy().copy())
    assert np.array_equal(outs[0], outs[1])


def test_u8_prefetch_pipeline_gives_the_same_scores():
    cfg = O.readme_config()
    m, _ = make(cfg, "cuda", precision="bf16")
    rng = np.random.default_rng(13)
    batches = [torch.from_numpy(rng.integers(0, 256, size=(3, 224, 300, 3), dtype=np.uint8)).pin_memory() for _ in range(5)]
    want = [m.score_host_u8(b).numpy().copy() for b in batches]
    got = []
    m.prefetch_host_u8(batches[0])
    for i, b in enumerate(batches):
        if i + 1 < len(batches):
            m.prefetch_host_u8(batches[i + 1])
        got.append(m.score_host_u8(b).numpy().copy())
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # a prefetch that is never consumed does not poison later calls (different pointer, different size)
    m.prefetch_host_u8(batches[0])
    other = torch.from_numpy(rng.integers(0, 256, size=(2, 100, 120, 3), dtype=np.uint8))
    sc = m.score_host_u8(other).numpy()
    ref = m.score(m.preprocess_u8(other), return_err=False)["score"].cpu().numpy()
    assert np.array_equal(sc, ref)
