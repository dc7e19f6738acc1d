"""Kernel-logic parity on CPU: the CUDA sources compiled by g++ against tests/emu/cuda_emu.h
(a functional simulator of blocks/threads/__syncthreads/shuffles) vs the oracle.  This does
NOT replace the GPU parity tests (tests/test_gpu_parity.py, -m gpu); it finds indexing and
orchestration bugs without spending GPU time and runs in the driver's CPU-only gate."""
import pytest

import parity_cases as PC
from kcvae_testlib import small_config

BACKEND = "emu"


def test_structure():
    PC.case_structure(BACKEND)


def test_forward_and_layers():
    PC.case_forward(BACKEND)
    PC.case_layers(BACKEND)


@pytest.mark.parametrize("kind", ["global", "single"])
@pytest.mark.parametrize("training", [True, False])
def test_loss(kind, training):
    PC.case_loss(BACKEND, kind, training=training)


@pytest.mark.parametrize("kind", ["global", "single"])
def test_grads(kind):
    PC.case_grads(BACKEND, kind)


@pytest.mark.parametrize("kind", ["global", "single"])
def test_train_steps(kind):
    PC.case_train_steps(BACKEND, kind)


def test_score():
    PC.case_score(BACKEND)


@pytest.mark.parametrize("shape", [
    dict(layers=(4,), enc=0, H=8, W=12),                 # one layer, no encoder Dense
    dict(layers=(4, 3, 5), enc=6, H=16, W=24),           # three layers
    dict(layers=(33,), enc=3, H=6, W=10, dec=9, latent=3),  # channel counts that are not tile-friendly
    dict(layers=(2, 2), enc=2, H=12, W=20, C_=1, latent=1),  # single channel, latent 1
    dict(layers=(32, 5), enc=6, H=16, W=24, dec=32, latent=6),  # README channel pattern: specialised fp32 kernels
    dict(layers=(32, 8), enc=0, H=8, W=20, dec=4, latent=4),    # few<->many edge cases (Co = 8, Ci = 4)
])
def test_topologies(shape):
    cfg = small_config(**shape)
    PC.case_forward(BACKEND, cfg, B=2)
    PC.case_grads(BACKEND, cfg=cfg, B=3)


def test_driver_replay(tmp_path):
    PC.case_driver_replay(BACKEND, tmp_path)


def test_error_behaviour():
    import numpy as np
    from kcvae_testlib import make, model_class, frames
    bad = small_config()
    bad["model"]["layers"] = [4] * 6
    with pytest.raises(RuntimeError, match="Collapse"):
        make(bad, BACKEND)
    cfg = small_config()
    del cfg["loss"]["w_x_std"]
    with pytest.raises(KeyError):
        model_class(BACKEND, "global")(cfg)
    m, _ = make(small_config(), BACKEND)
    with pytest.raises(ValueError):
        m.call(np.zeros((2, 8, 8, 3), np.float32))
    with pytest.raises(RuntimeError):
        m.train_step(frames(small_config(), 2))          # not compiled
    odd = small_config(H=18, W=24)                        # not divisible by 2^L: loss cannot broadcast
    mo, _ = make(odd, BACKEND)
    mo.encode(frames(odd, 1))                             # encode alone works, like the reference
    with pytest.raises(Exception, match="divisible"):
        mo.compute_loss(frames(odd, 1))


@pytest.mark.parametrize("hw", [(15, 21), (18, 26), (7, 9)])
def test_encoder_odd_sizes_specialised_kernels(hw):
    """SAME padding with odd sizes (pad_before = 1) through the specialised conv kernels: the
    loss cannot broadcast for such sizes (like the reference), but encode() must match."""
    import numpy as np
    import torch
    from kcvae_testlib import O, frames, make
    cfg = small_config(layers=(32, 5), enc=5, H=hw[0], W=hw[1], dec=8, latent=3)
    m, ws = make(cfg, BACKEND)
    x = frames(cfg, 2)
    mean, lv = m.encode(x)
    om, olv, _ = O.encoder_forward(O.topology(cfg), [torch.tensor(w) for w in ws], torch.tensor(x))
    np.testing.assert_allclose(mean.numpy(), om.numpy(), atol=3e-5)
    np.testing.assert_allclose(lv.numpy(), olv.numpy(), atol=3e-5)
