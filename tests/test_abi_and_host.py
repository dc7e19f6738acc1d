"""CPU-only checks: the C-ABI library exports every symbol include/kcvae.h declares (no
compute calls without a GPU), the product loader refuses to run without its CUDA library /
a GPU (no silent fallback), host-side factory logic."""
import ctypes
import importlib
import os
import re

import pytest

from kcvae_testlib import ROOT, pkg, small_config

_lib = importlib.import_module("trustedai-cl-vae-ad_b200._lib")


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "kcvae.h")).read()
    return sorted(set(re.findall(r"\b(kcvae_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_are_exported_by_the_cuda_library():
    build = importlib.import_module("trustedai-cl-vae-ad_b200.build")
    path = build.build()                      # nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/kcvae.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(syms)      # every declared entry point has a ctypes signature, and vice versa
    lib.kcvae_abi_version.restype = ctypes.c_int
    assert lib.kcvae_abi_version() == 1


def test_product_loader_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _lib._BINDING = None
    with pytest.raises(RuntimeError, match="no CPU fallback|There is no CPU fallback"):
        pkg.load_model_from_config(small_config())


def test_factory_type_dispatch_without_gpu():
    assert pkg.import_vae_based_on_type(None) is pkg.KurtosisGlobalCVAE
    assert pkg.import_vae_based_on_type("KurtosisGlobal") is pkg.KurtosisGlobalCVAE
    assert pkg.import_vae_based_on_type("KurtosisSingle") is pkg.KurtosisSingleCVAE
    with pytest.raises(NotImplementedError):
        pkg.import_vae_based_on_type("KLGaussian")
    with pytest.raises(Exception):
        pkg.import_vae_based_on_type("Nope")


def test_config_roundtrip(tmp_path):
    cfg = small_config()
    p = tmp_path / "config.yml"
    pkg.save_config(cfg, str(p))
    assert pkg.load_config(str(p)) == cfg
