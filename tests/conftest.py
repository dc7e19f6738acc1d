import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(autouse=True)
def _dense_engine_at_every_batch(monkeypatch):
    """The library routes the decoder Dense layer to the tensor-core engine from 64 frames up (below that its 17 MB weight
    matrix bounds it either way).  The parity tests run at 1..16 frames: lower the threshold so they drive that path too."""
    monkeypatch.setenv("KCVAE_GEN_DENSE_MIN_BATCH", "1")
    # the engine's Dense forward and its encoder-Dense products are implemented and verified but measured slower than the
    # CUDA-core kernels: off in the product, ON in the tests so that they stay correct
    monkeypatch.setenv("KCVAE_GEN_DENSE_FWD", "1")
    monkeypatch.setenv("KCVAE_GEN_EDENSE", "1")
