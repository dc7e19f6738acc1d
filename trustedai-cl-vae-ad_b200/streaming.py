"""The stages either side of the model in the reference's tools, on the GPU:

* :class:`StreamingAnomalyScore` - the per-frame anomaly score of the camera tool
  (camera_streamer_qt.py:1364-1408): per-pixel EMA moments of the error map, z-of-z threshold count,
  EMA-normalised uint8 error image.  Attribute names follow the tool's own state variables.
* :func:`render_outputs` - uint8 error image, JET heat map, 0.5/0.5 overlay, uint8 reconstruction
  (do_anomaly_detection.py:166-170, output_reconstructions.py:68-83, camera_streamer_qt.py:1417-1418).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .model import _ptr, _wrap


def _stream_ptr(binding, device):
    if binding.device_type == "cuda":
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return C.c_void_p(0)


class StreamingAnomalyScore:
    def __init__(self, height: int, width: int, device: Optional[int] = None, stream_error_ma: float = 0.99,
                 anomaly_score_ma_weight: float = 0.9, binding: Optional[_lib.Binding] = None):
        self._lib = binding or _lib.load()
        if self._lib.device_type == "cuda":
            idx = torch.cuda.current_device() if device is None else int(device)
            self.device = torch.device("cuda", idx)
        else:
            idx, self.device = 0, torch.device("cpu")
        self.H, self.W = int(height), int(width)
        self._h = C.c_void_p()
        rc = self._lib.stream_create(self.H, self.W, idx, C.byref(self._h))
        if rc < 0:
            raise _lib.KcvaeError(rc, (self._lib.stream_last_error(None) or b"").decode())
        self.stream_error_ma = float(stream_error_ma)                  # :213
        self.anomaly_score_ma_weight = float(anomaly_score_ma_weight)  # :219
        self.reset()

    def reset(self):
        self._lib.stream_reset(self._h)
        self.stream_error_min = 0.0     # :211-212
        self.stream_error_max = 0.0
        self.anomaly_score = 0.0
        self.anomaly_score_ma = 0.0     # :218
        self.anomaly_count = 0.0
        self.stream_error_img = None

    def update(self, err, return_image: bool = True) -> dict:
        """One frame.  ``err`` [H,W] = sum_c (x - x_rec)^2 (``model.score(x)['err'][i]``)."""
        e = err if isinstance(err, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(err, dtype=np.float32))
        e = e.as_subclass(torch.Tensor).to(self.device, torch.float32).contiguous()
        if tuple(e.shape) != (self.H, self.W):
            raise ValueError(f"expected an error map of shape [{self.H}, {self.W}], got {tuple(e.shape)}")
        img = torch.empty(self.H, self.W, dtype=torch.uint8, device=self.device) if return_image else None
        out = (C.c_float * 8)()
        rc = self._lib.stream_update(self._h, _ptr(e), self.stream_error_ma, _ptr(img), out, _stream_ptr(self._lib, self.device))
        if rc < 0:
            raise _lib.KcvaeError(rc, (self._lib.stream_last_error(self._h) or b"").decode())
        self.anomaly_count, self.anomaly_score = float(out[0]), float(out[1])
        self.stream_error_min, self.stream_error_max = float(out[4]), float(out[5])
        as_ma = self.anomaly_score_ma_weight                                           # :1404-1408
        ma = as_ma * self.anomaly_score_ma + (1.0 - as_ma) * self.anomaly_score
        if not math.isnan(ma):
            self.anomaly_score_ma = ma
        self.stream_error_img = img
        return {"anomaly_count": self.anomaly_count, "anomaly_score": self.anomaly_score,
                "anomaly_score_ma": self.anomaly_score_ma, "frame_min": float(out[2]), "frame_max": float(out[3]),
                "stream_error_min": self.stream_error_min, "stream_error_max": self.stream_error_max,
                "z_mean": float(out[6]), "z_std": float(out[7]), "stream_error_img": img}

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.stream_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class DeviceDataQueue:
    """The camera tool's ``DataQueue`` (camera_streamer_qt.py:61-81) kept on the GPU together with the replay
    buffer it is stacked with before every continual-learning step (:1341-1345): one preallocated
    ``[capacity + n_replay, H, W, C]`` device tensor, so ``train_step_and_run(queue.stacked())`` uploads nothing
    but the newest frame.  Same semantics: initialised with ``capacity`` copies of the first sample, ``append``
    advances ``_idx`` first and overwrites that slot, ``to_numpy`` returns the slots in storage order."""

    def __init__(self, data_sample, capacity: int, replay_buffer=None, device=None):
        assert capacity > 0
        first = torch.as_tensor(np.asarray(data_sample.cpu() if isinstance(data_sample, torch.Tensor) else data_sample),
                                dtype=torch.float32)
        n_rep = 0 if replay_buffer is None else int(replay_buffer.shape[0])
        self._buf = torch.empty((capacity + n_rep,) + tuple(first.shape), dtype=torch.float32, device=device)
        self._buf[:capacity] = first.to(self._buf.device)
        if n_rep:
            self._buf[capacity:] = torch.as_tensor(np.asarray(replay_buffer), dtype=torch.float32).to(self._buf.device)
        self._idx = 0
        self._capacity = capacity

    def append(self, x):
        self._increment()
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=torch.float32)
        self._buf[self._idx].copy_(t.as_subclass(torch.Tensor), non_blocking=True)

    def _increment(self):
        self._idx = (self._idx + 1) % self._capacity

    def get(self):
        return self._buf[self._idx]

    def stacked(self):
        """np.vstack((inf_buffer.to_numpy(), replay_buffer)) of :1342, as a device view (no copy)."""
        return self._buf

    def to_numpy(self):
        return self._buf[:self._capacity].cpu().numpy()


def render_outputs(norm_err=None, rec=None, device=None, binding: Optional[_lib.Binding] = None) -> dict:
    """``norm_err`` [B,H,W] in [0,1] and / or ``rec`` [B,H,W,C] in [0,1] -> uint8 tensors:
    ``err`` = round(255 norm_err), ``heatmap`` = cv2.applyColorMap(err, COLORMAP_JET) (OpenCV's channel order),
    ``rec`` = round(255 rec), ``overlay`` = cv2.addWeighted(heatmap, .5, rec, .5, 0)."""
    lib = binding or _lib.load()
    dev = torch.device("cpu") if lib.device_type != "cuda" else torch.device("cuda", torch.cuda.current_device() if device is None else int(device))

    def prep(a):
        if a is None:
            return None
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return t.as_subclass(torch.Tensor).to(dev, torch.float32).contiguous()

    ne, rc_ = prep(norm_err), prep(rec)
    if ne is None and rc_ is None:
        raise ValueError("render_outputs needs norm_err and / or rec")
    if ne is not None:
        B, H, W = ne.shape
        Cc = rc_.shape[3] if rc_ is not None else 3
    else:
        B, H, W, Cc = rc_.shape
    u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
    err_u8 = u8(B, H, W) if ne is not None else None
    heat = u8(B, H, W, Cc) if ne is not None else None
    rec_u8 = u8(B, H, W, Cc) if rc_ is not None else None
    over = u8(B, H, W, Cc) if (ne is not None and rc_ is not None) else None
    rc = lib.render_outputs(_ptr(ne), _ptr(rc_), B, H, W, Cc, _ptr(err_u8), _ptr(heat), _ptr(over), _ptr(rec_u8), _stream_ptr(lib, dev))
    if rc < 0:
        raise _lib.KcvaeError(rc, "render_outputs: invalid arguments or launch failure")
    return {"err": err_u8, "heatmap": heat, "overlay": over, "rec": rec_u8}
