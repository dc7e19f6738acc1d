"""ctypes binding of libkcvae.so (include/kcvae.h).

The product path is strict: :func:`load` opens only the in-tree CUDA library and raises if
it is missing or no GPU is usable - there is no CPU fallback.  ``Binding`` itself is
runtime-agnostic so the kernel-logic tests can bind the g++ emulation build
(tests/emu) explicitly; nothing in this package does that."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libkcvae.so")
MAX_LAYERS = 8
NUM_METRICS = 16

OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_COLLAPSE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
METRICS_FULL, METRICS_LOSS_ONLY = 0, 1
PREC_FP32, PREC_BF16_TC = 0, 1


class KcvaeConfig(C.Structure):
    _fields_ = [
        ("image_h", C.c_int32), ("image_w", C.c_int32), ("image_c", C.c_int32),
        ("n_layers", C.c_int32), ("layers", C.c_int32 * MAX_LAYERS),
        ("encoder_dense_filters", C.c_int32), ("decoder_dense_filters", C.c_int32),
        ("latent_dimensions", C.c_int32), ("model_type", C.c_int32),
        ("kurtosis_target", C.c_float), ("w_mse", C.c_float), ("w_kurtosis", C.c_float),
        ("w_skew", C.c_float), ("w_kl_divergence", C.c_float), ("w_z_l1_reg", C.c_float),
        ("w_x_std", C.c_float), ("beta", C.c_float), ("learning_rate", C.c_float),
        ("max_batch", C.c_int32), ("precision", C.c_int32),
    ]


_P = C.c_void_p
_SIGS = {
    "kcvae_abi_version": (C.c_int, []),
    "kcvae_create": (C.c_int, [C.POINTER(KcvaeConfig), C.c_int, C.POINTER(_P)]),
    "kcvae_destroy": (C.c_int, [_P]),
    "kcvae_last_error": (C.c_char_p, [_P]),
    "kcvae_num_variables": (C.c_int, [_P]),
    "kcvae_param_count": (C.c_int64, [_P]),
    "kcvae_variable_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kcvae_set_weights": (C.c_int, [_P, _P, C.c_int64]),
    "kcvae_get_weights": (C.c_int, [_P, _P, C.c_int64]),
    "kcvae_get_grads": (C.c_int, [_P, _P, C.c_int64]),
    "kcvae_weights_device": (_P, [_P]),
    "kcvae_grads_device": (_P, [_P]),
    "kcvae_init_glorot": (C.c_int, [_P, C.c_uint64, _P]),
    "kcvae_adam_reset": (C.c_int, [_P]),
    "kcvae_set_adam_state": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64]),
    "kcvae_get_adam_state": (C.c_int, [_P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "kcvae_set_learning_rate": (C.c_int, [_P, C.c_float]),
    "kcvae_set_beta": (C.c_int, [_P, C.c_float]),
    "kcvae_set_train_image_noise": (C.c_int, [_P, C.c_int]),
    "kcvae_set_loss_weights": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "kcvae_seed": (C.c_int, [_P, C.c_uint64]),
    "kcvae_comm_unique_id": (C.c_int, [_P]),
    "kcvae_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "kcvae_comm_world": (C.c_int, [_P]),
    "kcvae_broadcast_weights": (C.c_int, [_P, C.c_int, _P]),
    "kcvae_encode": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "kcvae_reparameterize": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "kcvae_decode": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "kcvae_forward": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "kcvae_loss": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P]),
    "kcvae_train_step": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, C.c_int, _P]),
    "kcvae_loss_and_grads": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, C.c_int, _P]),
    "kcvae_score": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "kcvae_normalize_scores": (C.c_int, [_P, _P, _P, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                         C.c_float, _P, _P, _P, _P]),
    "kcvae_prefetch_host": (C.c_int, [_P, _P, C.c_int]),
    "kcvae_train_step_host": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, C.c_int, _P]),
    "kcvae_score_host": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "kcvae_preprocess_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "kcvae_prefetch_host_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "kcvae_score_host_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "kcvae_train_step_host_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P]),
    "kcvae_stream_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "kcvae_stream_destroy": (C.c_int, [_P]),
    "kcvae_stream_reset": (C.c_int, [_P]),
    "kcvae_stream_last_error": (C.c_char_p, [_P]),
    "kcvae_stream_update": (C.c_int, [_P, _P, C.c_double, _P, _P, _P]),
    "kcvae_render_outputs": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "kcvae_launch_count": (C.c_int64, [_P]),
    "kcvae_debug_activation": (C.c_int64, [_P, C.c_int, _P, C.c_int64]),
    "kcvae_tc_status": (C.c_int, [_P]),
    "kcvae_profile_enable": (C.c_int, [C.c_int]),
    "kcvae_profile_report": (C.c_int64, [C.c_char_p, C.c_int64]),
    "kcvae_gen_conv_test": (C.c_int, [C.c_int] * 8 + [_P] * 5 + [C.c_int] * 5 + [_P]),
    "kcvae_gen_wgrad_test": (C.c_int, [C.c_int] * 4 + [_P] * 4 + [C.c_int] * 5 + [_P]),
    "kcvae_gen_dense_test": (C.c_int, [C.c_int] * 3 + [_P] * 4 + [C.c_int] * 3 + [_P]),
    "kcvae_gen_plan_dump": (C.c_int64, [C.c_int, _P, C.c_int, _P, C.c_int64]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


class KcvaeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kcvae error {code}: {msg}")
        self.code = code
        self.message = msg


class Binding:
    """Typed access to one loaded copy of the library.  ``device_type`` is the torch device
    type of the buffers the library expects ('cuda' for libkcvae.so)."""

    def __init__(self, cdll: C.CDLL, device_type: str = "cuda", path: str = ""):
        self.cdll = cdll
        self.device_type = device_type
        self.path = path
        for name, (res, args) in _SIGS.items():
            fn = getattr(cdll, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[len("kcvae_"):], fn)
        if self.abi_version() != 1:
            raise RuntimeError(f"libkcvae ABI {self.abi_version()} != 1")

    def check(self, rc: int, handle=None):
        if rc < 0:
            msg = self.last_error(handle)
            raise KcvaeError(rc, msg.decode() if msg else "unknown")
        return rc


_BINDING: Optional[Binding] = None


def load() -> Binding:
    """Load the CUDA library.  Raises (never falls back) when it cannot run on a GPU."""
    global _BINDING
    if _BINDING is not None:
        return _BINDING
    import torch
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    if not torch.cuda.is_available():
        raise RuntimeError("libkcvae.so needs a CUDA device (B200, sm_100a); none is visible. "
                           "There is no CPU fallback.")
    _BINDING = Binding(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL), "cuda", LIB_PATH)
    return _BINDING
