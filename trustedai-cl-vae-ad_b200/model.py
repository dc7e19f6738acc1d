"""Python-facing mirror of the reference model classes (src/abstract_cvae.py,
src/kurtosis_global_cvae.py, src/kurtosis_single_cvae.py) over the kcvae C ABI.

Same constructor (the config.yml dict), attributes and method names / argument order /
defaults as the reference, so train.py / do_anomaly_detection.py-style drivers run
unchanged against the object ``load_model_from_config`` returns.  All arithmetic happens
in libkcvae.so; torch is used only as the container for device buffers and streams."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .optimizers import Adam

GLOBAL_KEYS = ["loss", "mse", "z_l1", "var_loss", "skew_loss", "z_kurtosis_loss", "z_kurtosis",
               "r_min", "r_max", "cross_entropy", "kl_div", "x_std_loss"]   # src/kurtosis_global_cvae.py:93-106
SINGLE_KEYS = ["loss", "mse", "z_l1", "z_l2", "skew_loss", "z_kurtosis_loss", "z_kurtosis",
               "r_min", "r_max", "x_std_loss"]                              # src/kurtosis_single_cvae.py:62-73


class KTensor(torch.Tensor):
    """Device tensor returned by the model.  Behaves like the tf eager tensors callers of the
    reference expect: ``.numpy()`` works wherever the data lives, indexing and elementwise
    math return tensors, ``__dlpack__`` / ``__cuda_array_interface__`` come from torch."""

    def numpy(self, *a, **k):  # noqa: D401
        return self.detach().cpu().as_subclass(torch.Tensor).numpy(*a, **k)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a


def _wrap(t: torch.Tensor) -> KTensor:
    return t.as_subclass(KTensor)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# ------------------------------------------------------------------------------------------
# structural mirrors of the Keras objects the reference exposes (tests/test_kurtosis_*:72-148)
# ------------------------------------------------------------------------------------------
class Variable:
    """One trainable variable, living inside the library's flat fp32 parameter vector."""

    def __init__(self, model: "AbstractCVAE", index: int, name: str, shape: Sequence[int]):
        self._model, self.index, self.name, self.shape = model, index, name, tuple(int(s) for s in shape)
        self.dtype = np.float32
        self.trainable = True

    def numpy(self) -> np.ndarray:
        return self._model.get_weights()[self.index]

    def assign(self, value) -> "Variable":
        ws = self._model.get_weights()
        ws[self.index] = np.asarray(value, np.float32).reshape(self.shape)
        self._model.set_weights(ws)
        return self

    def __array__(self, dtype=None, copy=None):
        return self.numpy() if dtype is None else self.numpy().astype(dtype)

    def __repr__(self):
        return f"<kcvae.Variable '{self.name}' shape={self.shape} dtype=float32>"


class Layer:
    def __init__(self, kind: str, name: str, input_shape, output_shape, variables: List[Variable], **attrs):
        self.kind, self.name = kind, name
        self.input_shape, self.output_shape = tuple(input_shape), tuple(output_shape)
        self.variables = variables
        self.trainable_weights = variables
        self.trainable_variables = variables
        self.weights = variables
        for k, v in attrs.items():
            setattr(self, k, v)

    def get_weights(self):
        return [v.numpy() for v in self.variables]

    def count_params(self):
        return int(sum(np.prod(v.shape) for v in self.variables))

    def __repr__(self):
        return f"<kcvae.{self.kind} '{self.name}' {self.input_shape}->{self.output_shape}>"


class Sequential:
    """Stands where ``tf.keras.Sequential`` stands in the reference (``model.encoder`` /
    ``model.decoder``): ``.layers``, ``.summary()``, ``.save(path)`` and call."""

    def __init__(self, model: "AbstractCVAE", name: str, layers: List[Layer], runner):
        self._model, self.name, self.layers, self._runner = model, name, layers, runner

    @property
    def variables(self):
        return [v for l in self.layers for v in l.variables]

    trainable_weights = trainable_variables = weights = variables

    def count_params(self):
        return sum(l.count_params() for l in self.layers)

    def __call__(self, x, training=False):
        return self._runner(x)

    def get_weights(self):
        return [v.numpy() for v in self.variables]

    def set_weights(self, arrays):
        ws = self._model.get_weights()
        vs = self.variables
        assert len(arrays) == len(vs)
        for v, a in zip(vs, arrays):
            ws[v.index] = np.asarray(a, np.float32).reshape(v.shape)
        self._model.set_weights(ws)

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"')
        print_fn("_" * 65)
        print_fn(f"{'Layer (type)':<29}{'Output Shape':<26}{'Param #':<10}")
        print_fn("=" * 65)
        for l in self.layers:
            print_fn(f"{(l.name + ' (' + l.kind + ')'):<29}{str(l.output_shape):<26}{l.count_params():<10}")
        print_fn("=" * 65)
        print_fn(f"Total params: {self.count_params():,}")
        print_fn("_" * 65)

    def save(self, path: str, tf_variables: bool = True):
        """``encoder.save(dir)`` / ``decoder.save(dir)`` (train.py:127-128).  Writes ``<dir>/weights.npz`` keyed
        by Keras variable order + ``layers.json`` and, unless ``tf_variables=False``, the same variables as a
        TensorBundle under ``<dir>/variables/`` with the keys a Keras SavedModel uses (tf_bundle.py), readable with
        ``tf.train.load_checkpoint`` on the TensorFlow side.  No ``saved_model.pb`` is written."""
        os.makedirs(path, exist_ok=True)
        ws = self.get_weights()
        np.savez(os.path.join(path, "weights.npz"), **{f"{i:02d}": w for i, w in enumerate(ws)})
        spec = [{"kind": l.kind, "name": l.name, "input_shape": list(l.input_shape),
                 "output_shape": list(l.output_shape)} for l in self.layers]
        with open(os.path.join(path, "layers.json"), "w") as f:
            json.dump({"name": self.name, "layers": spec, "variables": [v.name for v in self.variables]}, f, indent=1)
        if tf_variables:
            from .tf_bundle import keras_weights_to_bundle
            keras_weights_to_bundle(os.path.join(path, "variables", "variables"), ws)

    def load(self, path: str):
        """Either this runtime's ``weights.npz`` or the ``variables/`` TensorBundle of a Keras SavedModel written by
        the reference (``vae.encoder.save(dir)``): variables are matched by layer order, kernel before bias."""
        npz = os.path.join(path, "weights.npz")
        if os.path.exists(npz):
            z = np.load(npz)
            self.set_weights([z[k] for k in sorted(z.files)])
            return
        from .tf_bundle import keras_weights_from_bundle
        ws = keras_weights_from_bundle(os.path.join(path, "variables", "variables"))
        vs = self.variables
        if len(ws) != len(vs) or any(tuple(w.shape) != v.shape for w, v in zip(ws, vs)):
            raise ValueError(f"{path}: SavedModel variables {[tuple(w.shape) for w in ws]} do not match this "
                             f"{self.name} {[v.shape for v in vs]} (config.yml differs from the checkpoint?)")
        self.set_weights(ws)


class History:
    def __init__(self):
        self.history: Dict[str, list] = {}
        self.epoch: List[int] = []


class Callback:
    """Minimal stand-in for ``tf.keras.callbacks.Callback`` (train.py:40-47)."""
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): ...
    def on_train_end(self, logs=None): ...
    def on_epoch_begin(self, epoch, logs=None): ...
    def on_epoch_end(self, epoch, logs=None): ...
    def on_train_batch_end(self, batch, logs=None): ...


class BetaAnnealingCallback(Callback):
    """train.py:40-47: ``model.beta *= rate`` at every epoch end."""

    def __init__(self, rate=0.98):
        self.rate = rate

    def on_epoch_end(self, epoch, logs=None):
        self.model.beta *= self.rate


# ------------------------------------------------------------------------------------------
class AbstractCVAE:
    """src/abstract_cvae.py:7-178 over libkcvae.so."""

    MODEL_TYPE = 0
    _binding_override: Optional[_lib.Binding] = None   # tests/emu only

    def __init__(self, config, device: Optional[int] = None, precision: Optional[str] = None,
                 metrics: str = "full"):
        self.config = config
        self.beta = float(config["training"]["beta"])                       # :14
        self.encoder_input_shape = config["data"]["image_size"]             # :15
        self.latent_size = config["model"]["latent_dimensions"]             # :16
        self._read_loss_config(config["loss"])
        self._lib = self._binding_override or _lib.load()
        self._dev_type = self._lib.device_type
        if self._dev_type == "cuda":
            self._device_index = torch.cuda.current_device() if device is None else int(device)
            self.device = torch.device("cuda", self._device_index)
        else:
            self._device_index, self.device = 0, torch.device("cpu")
        self.metric_tier = _lib.METRICS_FULL if metrics == "full" else _lib.METRICS_LOSS_ONLY
        self.train_image_noise = False        # opt-in README behaviour (SURVEY Note A)
        cfg = self._c_config(precision)
        self._h = C.c_void_p()
        rc = self._lib.create(C.byref(cfg), self._device_index, C.byref(self._h))
        if rc < 0:
            msg = (self._lib.last_error(None) or b"").decode()
            if rc == _lib.ERR_COLLAPSE:
                raise RuntimeError(msg)                                      # :65-68
            raise _lib.KcvaeError(rc, msg)
        self._nvars = self._lib.num_variables(self._h)
        self._nparams = int(self._lib.param_count(self._h))
        self._var_shapes = []
        for i in range(self._nvars):
            rank, dims, off = C.c_int32(), (C.c_int64 * 4)(), C.c_int64()
            self._lib.check(self._lib.variable_info(self._h, i, C.byref(rank), dims, C.byref(off)), self._h)
            self._var_shapes.append(tuple(dims[k] for k in range(rank.value)))
        self.optimizer: Optional[Adam] = None
        self._metrics_buf = torch.zeros(_lib.NUM_METRICS, dtype=torch.float32, device=self.device)
        self._dist_world = 1
        self.encoder = self._build_encoder()                                # :18
        self.decoder = self._build_decoder()                                # :19
        seed = int(config.get("seed", 0)) if isinstance(config, dict) else 0
        self._lib.check(self._lib.init_glorot(self._h, seed or 0x5EED, self._stream()), self._h)
        self.stop_training = False

    # -- config -------------------------------------------------------------------------
    def _read_loss_config(self, loss_config):
        raise NotImplementedError

    def _c_config(self, precision) -> _lib.KcvaeConfig:
        m, d, t = self.config["model"], self.config["data"], self.config["training"]
        cfg = _lib.KcvaeConfig()
        cfg.image_h, cfg.image_w, cfg.image_c = [int(v) for v in d["image_size"]]
        layers = [int(f) for f in m["layers"]]
        if len(layers) > _lib.MAX_LAYERS:
            # the reference fails later with the collapse error for any realistic image size
            h0 = int(float(cfg.image_h) / float(2 ** len(layers)))
            raise RuntimeError(f"Error: Build Decoder: Width Collapse: Too many layers, check configuration file: {cfg.image_h} -> {h0}: {len(layers)} Layers")
        cfg.n_layers = len(layers)
        for i, f in enumerate(layers):
            cfg.layers[i] = f
        edf = m.get("encoder_dense_filters")
        cfg.encoder_dense_filters = int(edf) if edf else 0
        cfg.decoder_dense_filters = int(m["decoder_dense_filters"])
        cfg.latent_dimensions = int(m["latent_dimensions"])
        cfg.model_type = self.MODEL_TYPE
        cfg.kurtosis_target, cfg.w_mse, cfg.w_kurtosis = self.kurtosis_target, self.w_mse, self.w_kurtosis
        cfg.w_skew, cfg.w_z_l1_reg = self.w_skew, self.w_z_l1_reg
        cfg.w_kl_divergence = getattr(self, "w_kl_divergence", 0.0)
        cfg.w_x_std = getattr(self, "w_x_std", 0.0)
        cfg.beta = self.beta
        cfg.learning_rate = float(t.get("learning_rate", 1e-3))
        cfg.max_batch = 0
        # default: the tcgen05 tensor-core path (bf16 operands - hi + lo pairs where the loss needs fp32-grade products -
        # fp32 accumulation; holds north_star's 1e-3 / 1e-2 bars).  precision="fp32": CUDA-core kernels, tighter bars.
        prec = precision or m.get("precision", "bf16")
        cfg.precision = _lib.PREC_FP32 if str(prec).lower() in ("fp32", "f32", "float32") else _lib.PREC_BF16_TC
        return cfg

    # -- topology mirrors (src/abstract_cvae.py:22-92) -----------------------------------------
    def _build_encoder(self) -> Sequential:
        H, W, Cc = [int(v) for v in self.config["data"]["image_size"]]
        layers, vi = [], 0
        h, w, c = H, W, Cc
        for i, f in enumerate(self.config["model"]["layers"]):
            oh, ow = (h + 1) // 2, (w + 1) // 2
            vs = [Variable(self, vi, f"conv2d_{i}/kernel:0", self._var_shapes[vi]),
                  Variable(self, vi + 1, f"conv2d_{i}/bias:0", self._var_shapes[vi + 1])]
            layers.append(Layer("Conv2D", f"conv2d_{i}", (None, h, w, c), (None, oh, ow, int(f)), vs,
                                filters=int(f), kernel_size=(3, 3), strides=(2, 2), padding="same", activation="relu"))
            h, w, c, vi = oh, ow, int(f), vi + 2
        flat = h * w * c
        layers.append(Layer("Flatten", "flatten", (None, h, w, c), (None, flat), []))
        edf = self.config["model"].get("encoder_dense_filters")
        k = flat
        if edf:
            vs = [Variable(self, vi, "dense/kernel:0", self._var_shapes[vi]), Variable(self, vi + 1, "dense/bias:0", self._var_shapes[vi + 1])]
            layers.append(Layer("Dense", "dense", (None, k), (None, int(edf)), vs, units=int(edf), activation="linear"))
            k, vi = int(edf), vi + 2
        L2 = 2 * int(self.latent_size)
        vs = [Variable(self, vi, "dense_1/kernel:0", self._var_shapes[vi]), Variable(self, vi + 1, "dense_1/bias:0", self._var_shapes[vi + 1])]
        layers.append(Layer("Dense", "dense_1", (None, k), (None, L2), vs, units=L2, activation="linear"))
        self._n_enc_vars = vi + 2
        return Sequential(self, "encoder", layers, self._run_encoder_seq)

    def _build_decoder(self) -> Sequential:
        H, W, Cc = [int(v) for v in self.config["data"]["image_size"]]
        fl = [int(f) for f in self.config["model"]["layers"]]
        n = len(fl)
        h0, w0 = int(float(H) / float(2 ** n)), int(float(W) / float(2 ** n))
        ddf = int(self.config["model"]["decoder_dense_filters"])
        vi = self._n_enc_vars
        units = h0 * w0 * ddf
        layers = []
        vs = [Variable(self, vi, "dense_2/kernel:0", self._var_shapes[vi]), Variable(self, vi + 1, "dense_2/bias:0", self._var_shapes[vi + 1])]
        layers.append(Layer("Dense", "dense_2", (None, int(self.latent_size)), (None, units), vs, units=units, activation="relu"))
        layers.append(Layer("Reshape", "reshape", (None, units), (None, h0, w0, ddf), [], target_shape=(h0, w0, ddf)))
        vi += 2
        h, w, c = h0, w0, ddf
        for i, f in enumerate(reversed(fl)):
            vs = [Variable(self, vi, f"conv2d_transpose_{i}/kernel:0", self._var_shapes[vi]),
                  Variable(self, vi + 1, f"conv2d_transpose_{i}/bias:0", self._var_shapes[vi + 1])]
            layers.append(Layer("Conv2DTranspose", f"conv2d_transpose_{i}", (None, h, w, c), (None, 2 * h, 2 * w, f), vs,
                                filters=f, kernel_size=(3, 3), strides=(2, 2), padding="same", activation="relu"))
            h, w, c, vi = 2 * h, 2 * w, f, vi + 2
        vs = [Variable(self, vi, f"conv2d_transpose_{n}/kernel:0", self._var_shapes[vi]),
              Variable(self, vi + 1, f"conv2d_transpose_{n}/bias:0", self._var_shapes[vi + 1])]
        layers.append(Layer("Conv2DTranspose", f"conv2d_transpose_{n}", (None, h, w, c), (None, h, w, Cc), vs,
                            filters=Cc, kernel_size=(3, 3), strides=(1, 1), padding="same", activation="linear"))
        return Sequential(self, "decoder", layers, lambda z: self.decode(z, False))

    # -- plumbing -----------------------------------------------------------------------
    def _stream(self):
        if self._dev_type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    def _to_dev(self, x, shape_tail=None) -> torch.Tensor:
        """numpy / torch / DLPack / __cuda_array_interface__ -> contiguous fp32 tensor on
        the model's device (SURVEY 8b tensor interop)."""
        if isinstance(x, torch.Tensor):
            t = x.as_subclass(torch.Tensor)
        elif isinstance(x, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        elif hasattr(x, "__cuda_array_interface__"):
            t = torch.as_tensor(x, device=self.device)
        elif hasattr(x, "__dlpack__"):
            t = torch.from_dlpack(x)
        else:
            t = torch.as_tensor(np.asarray(x, dtype=np.float32))
        t = t.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        if shape_tail is not None and tuple(t.shape[1:]) != tuple(shape_tail):
            raise ValueError(f"expected input of shape [B, {', '.join(map(str, shape_tail))}], got {tuple(t.shape)}")
        return t

    def _empty(self, *shape) -> torch.Tensor:
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def _img_shape(self):
        return tuple(int(v) for v in self.config["data"]["image_size"])

    def _dict(self, metrics: torch.Tensor) -> Dict[str, KTensor]:
        m = _wrap(metrics.clone())
        return {k: m[i] for i, k in enumerate(self.METRIC_KEYS)}

    def _push_hparams(self):
        if self.optimizer is not None:
            self._lib.set_learning_rate(self._h, float(_as_float(self.optimizer.learning_rate)))
        self._lib.set_beta(self._h, float(self.beta))
        self._lib.set_train_image_noise(self._h, int(bool(self.train_image_noise)))
        self._lib.set_loss_weights(self._h, self.kurtosis_target, self.w_mse, self.w_kurtosis, self.w_skew, self.w_z_l1_reg)

    def close(self):
        """Release the library handle (device memory, streams, the NCCL communicator) NOW.  The object graph has reference
        cycles (variables point back at the model), so ``del model`` alone leaves destruction to the cyclic garbage
        collector at an arbitrary later time - under data parallel that would tear a communicator down while the peers
        are inside another collective.  Call it on every rank at the same point (bench.py does, between barriers)."""
        if getattr(self, "_h", None) and self._h.value:
            self._lib.destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ---------------------------------------------------------------------------
    @property
    def trainable_weights(self) -> List[Variable]:
        return self.encoder.variables + self.decoder.variables

    trainable_variables = weights = variables = trainable_weights

    def count_params(self) -> int:
        return self._nparams

    def get_weights(self) -> List[np.ndarray]:
        flat = np.empty(self._nparams, np.float32)
        self._lib.check(self._lib.get_weights(self._h, flat.ctypes.data_as(C.c_void_p), self._nparams), self._h)
        return self._split(flat)

    def set_weights(self, arrays: Sequence[np.ndarray]):
        assert len(arrays) == self._nvars, f"expected {self._nvars} arrays"
        parts = []
        for a, s in zip(arrays, self._var_shapes):
            a = np.asarray(a, np.float32)
            if tuple(a.shape) != s:
                raise ValueError(f"weight shape {a.shape} != {s}")
            parts.append(a.ravel())
        flat = np.ascontiguousarray(np.concatenate(parts))
        self._lib.check(self._lib.set_weights(self._h, flat.ctypes.data_as(C.c_void_p), self._nparams), self._h)

    def get_gradients(self) -> List[np.ndarray]:
        """dL/dw of the last train_step / loss_and_grads, Keras variable order."""
        flat = np.empty(self._nparams, np.float32)
        self._lib.check(self._lib.get_grads(self._h, flat.ctypes.data_as(C.c_void_p), self._nparams), self._h)
        return self._split(flat)

    def _split(self, flat):
        out, o = [], 0
        for s in self._var_shapes:
            n = int(np.prod(s))
            out.append(flat[o:o + n].reshape(s).copy())
            o += n
        return out

    def get_optimizer_state(self):
        m, v, t = np.empty(self._nparams, np.float32), np.empty(self._nparams, np.float32), C.c_int64()
        self._lib.check(self._lib.get_adam_state(self._h, m.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                                 self._nparams, C.byref(t)), self._h)
        return self._split(m), self._split(v), int(t.value)

    def set_optimizer_state(self, m, v, t: int):
        fm = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32).ravel() for a in m]))
        fv = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32).ravel() for a in v]))
        self._lib.check(self._lib.set_adam_state(self._h, fm.ctypes.data_as(C.c_void_p), fv.ctypes.data_as(C.c_void_p),
                                                 self._nparams, int(t)), self._h)

    def load_model(self, model_path):
        """src/abstract_cvae.py:95-106: ``<dir>/encoder`` + ``<dir>/decoder``."""
        assert os.path.exists(model_path)
        assert os.path.isdir(model_path)
        encoder_path = os.path.join(model_path, "encoder")
        assert os.path.exists(encoder_path)
        decoder_path = os.path.join(model_path, "decoder")
        assert os.path.exists(decoder_path)
        self.encoder.load(encoder_path)
        self.decoder.load(decoder_path)
        opt = os.path.join(model_path, "optimizer.npz")
        if os.path.exists(opt):
            z = np.load(opt)
            n = self._nvars
            self.set_optimizer_state([z[f"m{i:02d}"] for i in range(n)], [z[f"v{i:02d}"] for i in range(n)], int(z["t"]))

    def save_optimizer(self, model_path):
        """Adam state, which the reference never checkpoints."""
        m, v, t = self.get_optimizer_state()
        d = {f"m{i:02d}": a for i, a in enumerate(m)}
        d.update({f"v{i:02d}": a for i, a in enumerate(v)})
        np.savez(os.path.join(model_path, "optimizer.npz"), t=np.int64(t), **d)

    # -- Keras-style training surface ------------------------------------------------------
    def compile(self, optimizer=None, **_):
        """``vae.compile(optimizer=tf.keras.optimizers.Adam(lr))`` (train.py:99-101).  Any
        object with a ``learning_rate`` attribute is accepted; the update rule is Keras Adam."""
        self.optimizer = optimizer if optimizer is not None else Adam(float(self.config["training"]["learning_rate"]))
        self._lib.check(self._lib.adam_reset(self._h), self._h)

    def seed(self, seed: int):
        self._lib.seed(self._h, int(seed))

    def distribute(self):
        """Attach this replica to the default torch.distributed group: data-parallel training
        with NCCL all-reduce of gradients and of the latent/image moment sums."""
        from .dist import attach
        attach(self)
        return self

    # -- forward (src/abstract_cvae.py:109-149) -----------------------------------------------
    def sample(self, eps=None):
        if eps is None:
            eps = torch.randn(100, int(self.latent_size), device=self.device)
        return self.decode(eps, apply_sigmoid=True)

    def encode(self, x, training=False, noise=None):
        xt = self._to_dev(x, self._img_shape())
        B = xt.shape[0]
        mean, logvar = self._empty(B, self.latent_size), self._empty(B, self.latent_size)
        nz = self._to_dev(noise, self._img_shape()) if noise is not None else None
        self._push_hparams()
        self._lib.check(self._lib.encode(self._h, _ptr(xt), B, int(bool(training)), _ptr(nz), _ptr(mean), _ptr(logvar),
                                         self._stream()), self._h)
        return _wrap(mean), _wrap(logvar)

    def _run_encoder_seq(self, x):
        mean, logvar = self.encode(x, False)
        return _wrap(torch.cat([mean, logvar], dim=1))

    def reparameterize(self, mean, logvar, training=False, eps=None):
        m, lv = self._to_dev(mean), self._to_dev(logvar)
        B = m.shape[0]
        z = self._empty(B, self.latent_size)
        e = self._to_dev(eps) if eps is not None else None
        self._lib.check(self._lib.reparameterize(self._h, _ptr(m), _ptr(lv), B, int(bool(training)), _ptr(e), _ptr(z),
                                                 self._stream()), self._h)
        return _wrap(z)

    def decode(self, z, apply_sigmoid=False):
        zt = self._to_dev(z, (int(self.latent_size),))
        B = zt.shape[0]
        H, W, Cc = self._img_shape()
        n = len(self.config["model"]["layers"])
        oh, ow = int(float(H) / float(2 ** n)) * 2 ** n, int(float(W) / float(2 ** n)) * 2 ** n
        out = self._empty(B, oh, ow, Cc)
        self._lib.check(self._lib.decode(self._h, _ptr(zt), B, int(bool(apply_sigmoid)), _ptr(out), self._stream()), self._h)
        return _wrap(out)

    def call_detailed(self, x, training=False, eps=None):
        xt = self._to_dev(x, self._img_shape())
        B = xt.shape[0]
        xh = self._empty(*xt.shape)
        z, mean, logvar = (self._empty(B, self.latent_size) for _ in range(3))
        e = self._to_dev(eps) if eps is not None else None
        self._lib.check(self._lib.forward(self._h, _ptr(xt), B, int(bool(training)), _ptr(e), _ptr(xh), _ptr(z), _ptr(mean),
                                          _ptr(logvar), self._stream()), self._h)
        return _wrap(xh), _wrap(z), _wrap(mean), _wrap(logvar)

    def call(self, x, training=False, eps=None):
        xt = self._to_dev(x, self._img_shape())
        xh = self._empty(*xt.shape)
        e = self._to_dev(eps) if eps is not None else None
        self._lib.check(self._lib.forward(self._h, _ptr(xt), xt.shape[0], int(bool(training)), _ptr(e), _ptr(xh), None, None,
                                          None, self._stream()), self._h)
        return _wrap(xh)

    __call__ = call

    def predict(self, x, batch_size=None, verbose=0):
        if isinstance(x, (np.ndarray, torch.Tensor)):
            return self.call(x, False).numpy()
        return np.concatenate([self.call(b, False).numpy() for b in x], axis=0)

    # -- loss / steps (src/abstract_cvae.py:151-178) --------------------------------------------
    def compute_loss(self, x, training=False, return_inf=False, eps=None):
        xt = self._to_dev(x, self._img_shape())
        xh = self._empty(*xt.shape) if return_inf else None
        e = self._to_dev(eps) if eps is not None else None
        self._push_hparams()
        self._lib.check(self._lib.loss(self._h, _ptr(xt), xt.shape[0], int(bool(training)), _ptr(e), _ptr(self._metrics_buf),
                                       _ptr(xh), self.metric_tier, self._stream()), self._h)
        d = self._dict(self._metrics_buf)
        return (d, _wrap(xh)) if return_inf else d

    def _step(self, x, eps, want_xhat, update=True, noise=None):
        if self.optimizer is None and update:
            raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer=...)`.")
        xt = self._to_dev(x, self._img_shape())
        xh = self._empty(*xt.shape) if want_xhat else None
        e = self._to_dev(eps) if eps is not None else None
        nz = None
        if noise is not None:
            nz = self._to_dev(noise, self._img_shape())
        self._push_hparams()          # carries train_image_noise: the library draws N(0, beta^2) with its own Philox stream
        if update:
            rc = self._lib.train_step(self._h, _ptr(xt), xt.shape[0], _ptr(e), _ptr(nz), _ptr(self._metrics_buf), _ptr(xh),
                                      self.metric_tier, self._stream())
        else:
            rc = self._lib.loss_and_grads(self._h, _ptr(xt), xt.shape[0], _ptr(e), _ptr(self._metrics_buf), _ptr(xh),
                                          self.metric_tier, self._stream())
        self._lib.check(rc, self._h)
        if update and self.optimizer is not None and hasattr(self.optimizer, "iterations"):
            self.optimizer.iterations += 1
        d = self._dict(self._metrics_buf)
        return (d, _wrap(xh)) if want_xhat else d

    def train_step(self, x, eps=None, noise=None):
        return self._step(x, eps, False, noise=noise)

    def test_step(self, x):
        return self.compute_loss(x, training=False)

    def train_step_and_run(self, x, eps=None):
        return self._step(x, eps, True)

    def loss_and_grads(self, x, eps=None):
        """tape.gradient(loss['loss'], trainable_weights) without the optimizer update."""
        d = self._step(x, eps, False, update=False)
        return d, self.get_gradients()

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
            shuffle=True, steps_per_epoch=None, **_):
        """Minimal ``tf.keras.Model.fit`` (train.py:123): per batch ``train_step``, per epoch
        ``test_step`` over validation data, callback hooks (BetaAnnealingCallback, :40-47).
        Like Keras with a dict-returning custom step, logs hold the last batch's values."""
        callbacks = list(callbacks or [])
        hist = History()
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            else:
                cb.model = self
        _call(callbacks, "on_train_begin")
        self.stop_training = False
        for epoch in range(int(epochs)):
            _call(callbacks, "on_epoch_begin", epoch)
            logs = {}
            for bi, batch in enumerate(_batches(x, batch_size, shuffle)):
                if steps_per_epoch is not None and bi >= steps_per_epoch:
                    break
                d = self.train_step(batch)
                logs = d
                _call(callbacks, "on_train_batch_end", bi, d)
            logs = {k: float(v) for k, v in logs.items()}
            if validation_data is not None:
                v = {}
                for batch in _batches(validation_data, batch_size, False):
                    v = self.test_step(batch)
                logs.update({"val_" + k: float(t) for k, t in v.items()})
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.6g}" for k, v in logs.items()))
            _call(callbacks, "on_epoch_end", epoch, logs)
            if self._dev_type == "cuda":
                self.tc_status()          # raises if a bounded tcgen05 barrier wait expired during this epoch
            if self.stop_training:
                break
        _call(callbacks, "on_train_end")
        return hist

    # -- anomaly scoring primitive (do_anomaly_detection.py:61-62, 86-90) ----------------------
    def score(self, x, return_err=True, return_rec=False):
        """err = sum_c (x - call(x))^2 [B,H,W], per-frame sum [B], per-frame (min,max) [B,2]."""
        xt = self._to_dev(x, self._img_shape())
        B, H, W, _ = xt.shape
        err = self._empty(B, H, W) if return_err else None
        sc, mm = self._empty(B), self._empty(B, 2)
        xh = self._empty(*xt.shape) if return_rec else None
        self._lib.check(self._lib.score(self._h, _ptr(xt), B, _ptr(err), _ptr(sc), _ptr(mm), _ptr(xh), self._stream()), self._h)
        return {"err": _wrap(err) if err is not None else None, "score": _wrap(sc), "err_minmax": _wrap(mm),
                "rec": _wrap(xh) if xh is not None else None}

    # -- host-buffer entry points: H2D / D2H inside the call (bench.py e2e) ---------------------
    def train_step_host(self, x_host: torch.Tensor, eps_host: Optional[torch.Tensor] = None,
                        metrics_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One train_step on HOST buffers (ideally pinned): copies x (and eps) to the GPU, runs
        the step, copies the metrics vector back and synchronises.  Returns float32[16]."""
        if self.optimizer is None:
            raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer=...)`.")
        assert x_host.device.type == "cpu" and x_host.dtype == torch.float32 and x_host.is_contiguous()
        out = metrics_host if metrics_host is not None else torch.empty(_lib.NUM_METRICS, dtype=torch.float32)
        self._push_hparams()
        self._lib.check(self._lib.train_step_host(self._h, _ptr(x_host), x_host.shape[0], _ptr(eps_host), _ptr(out), None,
                                                  self.metric_tier, self._stream()), self._h)
        return out

    def prefetch_host(self, x_host: torch.Tensor):
        """Start copying the NEXT step's (pinned) host frames while the current step runs."""
        assert x_host.device.type == "cpu" and x_host.dtype == torch.float32 and x_host.is_contiguous()
        self._lib.check(self._lib.prefetch_host(self._h, _ptr(x_host), x_host.shape[0]), self._h)

    def score_host(self, x_host: torch.Tensor, score_host: Optional[torch.Tensor] = None,
                   err_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert x_host.device.type == "cpu" and x_host.dtype == torch.float32 and x_host.is_contiguous()
        out = score_host if score_host is not None else torch.empty(x_host.shape[0], dtype=torch.float32)
        self._lib.check(self._lib.score_host(self._h, _ptr(x_host), x_host.shape[0], _ptr(err_host), _ptr(out),
                                             self._stream()), self._h)
        return out

    # -- uint8 front end (src/data_loader.py:10-20, camera_streamer_qt.py:1296) ------------------
    def _u8_dev(self, frames) -> torch.Tensor:
        t = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != self._img_shape()[2]:
            raise ValueError(f"expected uint8 frames [B, h, w, {self._img_shape()[2]}], got {t.dtype} {tuple(t.shape)}")
        return t.to(self.device, non_blocking=True).contiguous()

    def preprocess_u8(self, frames):
        """uint8 NHWC frames of any size -> the model's fp32 input: ``tf.image.resize(frames / 255.,
        image_size[:2], antialias=True)`` (only the cast when the size already matches)."""
        t = self._u8_dev(frames)
        H, W, Cc = self._img_shape()
        out = self._empty(t.shape[0], H, W, Cc)
        self._lib.check(self._lib.preprocess_u8(self._h, _ptr(t), t.shape[0], t.shape[1], t.shape[2], _ptr(out),
                                                self._stream()), self._h)
        return _wrap(out)

    def prefetch_host_u8(self, frames_host: torch.Tensor):
        """Start copying the NEXT call's (pinned) uint8 host frames while the current call computes."""
        assert frames_host.device.type == "cpu" and frames_host.dtype == torch.uint8 and frames_host.is_contiguous()
        B, ih, iw, _ = frames_host.shape
        self._lib.check(self._lib.prefetch_host_u8(self._h, _ptr(frames_host), B, ih, iw), self._h)

    def score_host_u8(self, frames_host: torch.Tensor, score_host: Optional[torch.Tensor] = None,
                      err_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """score_host fed with uint8 host frames [B,h,w,C]: a quarter of the H2D bytes, cast / resize on the GPU."""
        assert frames_host.device.type == "cpu" and frames_host.dtype == torch.uint8 and frames_host.is_contiguous()
        B, ih, iw, _ = frames_host.shape
        out = score_host if score_host is not None else torch.empty(B, dtype=torch.float32)
        self._lib.check(self._lib.score_host_u8(self._h, _ptr(frames_host), B, ih, iw, _ptr(err_host), _ptr(out),
                                                self._stream()), self._h)
        return out

    def train_step_host_u8(self, frames_host: torch.Tensor, eps_host: Optional[torch.Tensor] = None,
                           metrics_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.optimizer is None:
            raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer=...)`.")
        assert frames_host.device.type == "cpu" and frames_host.dtype == torch.uint8 and frames_host.is_contiguous()
        B, ih, iw, _ = frames_host.shape
        out = metrics_host if metrics_host is not None else torch.empty(_lib.NUM_METRICS, dtype=torch.float32)
        self._push_hparams()
        self._lib.check(self._lib.train_step_host_u8(self._h, _ptr(frames_host), B, ih, iw, _ptr(eps_host), _ptr(out), None,
                                                     self.metric_tier, self._stream()), self._h)
        return out

    def tc_status(self) -> int:
        """1 if the tcgen05 kernels are active, 0 if only the fp32 path runs; raises on a pipeline error."""
        return self._lib.check(self._lib.tc_status(self._h), self._h)

    def profile(self, on: bool):
        self._lib.profile_enable(int(bool(on)))

    def profile_report(self) -> Dict[str, tuple]:
        """{'<layer tag>/<kernel>': (launcher calls, total ms)} since profiling was enabled."""
        n = int(self._lib.profile_report(None, 0))
        buf = C.create_string_buffer(n + 1)
        self._lib.profile_report(buf, n + 1)
        out = {}
        for line in buf.value.decode().splitlines():
            k, c, ms = line.rsplit(" ", 2)
            out[k] = (int(c), float(ms))
        return out

    def launch_count(self) -> int:
        return int(self._lib.launch_count(self._h))

    def debug_activation(self, which: int) -> np.ndarray:
        n = int(self._lib.debug_activation(self._h, which, None, 0))
        if n < 0:
            self._lib.check(n, self._h)
        out = np.empty(n, np.float32)
        self._lib.check(int(self._lib.debug_activation(self._h, which, out.ctypes.data_as(C.c_void_p), n)), self._h)
        return out


def _as_float(v):
    if hasattr(v, "numpy"):
        return float(v.numpy())
    return float(v)


def _call(cbs, name, *a):
    for cb in cbs:
        fn = getattr(cb, name, None)
        if fn is not None:
            fn(*a)


def _batches(data, batch_size, shuffle):
    if data is None:
        return
    if isinstance(data, (np.ndarray, torch.Tensor)):
        n = data.shape[0]
        bs = int(batch_size or 32)
        idx = np.random.permutation(n) if shuffle else np.arange(n)
        for i in range(0, n, bs):
            j = idx[i:i + bs]
            yield data[j] if isinstance(data, np.ndarray) else data[torch.as_tensor(j)]
    else:
        for b in data:   # tf.data.Dataset / list of batches: already batched (src/raite_loader.py:31,53)
            yield b


class KurtosisGlobalCVAE(AbstractCVAE):
    """src/kurtosis_global_cvae.py:9-110."""
    MODEL_TYPE = 0
    METRIC_KEYS = GLOBAL_KEYS

    def _read_loss_config(self, loss_config):
        self.kurtosis_target = float(loss_config["kurtosis"])            # :15-21 (KeyError if absent)
        self.w_mse = float(loss_config["w_mse"])
        self.w_kurtosis = float(loss_config["w_kurtosis"])
        self.w_skew = float(loss_config["w_skew"])
        self.w_kl_divergence = float(loss_config["w_kl_divergence"])
        self.w_z_l1_reg = float(loss_config["w_z_l1_reg"])
        self.w_x_std = float(loss_config["w_x_std"])

    def compute_loss_new(self, x, training=False, return_inf=False):
        return self.compute_loss(x, training, return_inf)


class KurtosisSingleCVAE(AbstractCVAE):
    """src/kurtosis_single_cvae.py:9-77."""
    MODEL_TYPE = 1
    METRIC_KEYS = SINGLE_KEYS

    def _read_loss_config(self, loss_config):
        self.kurtosis_target = float(loss_config["kurtosis"])            # :15-19
        self.w_mse = float(loss_config["w_mse"])
        self.w_kurtosis = float(loss_config["w_kurtosis"])
        self.w_skew = float(loss_config["w_skew"])
        self.w_z_l1_reg = float(loss_config["w_z_l1_reg"])
