"""A minimal torch-backed stand-in for the slice of TensorFlow 2.x / Keras that the reference's hot
path touches, so that the reference's OWN Python (src/abstract_cvae.py, src/kurtosis_*_cvae.py,
do_anomaly_detection.py under /root/reference) can be imported and executed in this container,
where TensorFlow itself cannot be installed.

TEST INFRASTRUCTURE ONLY - used by tests/golden/make_reference_goldens.py to generate the
fixtures in this directory.  What this pins: the reference's composition (topology wiring, loss
algebra, train_step, scoring loops) executed line by line.  What it cannot pin: TensorFlow's op
semantics themselves, which are restated here from the TF documentation (SAME padding,
conv2d_transpose as the input-gradient of conv2d, population std, divide_no_nan, Keras Adam) -
deliberately with a different formulation than oracle/kcvae_oracle.py (tap-by-tap strided slices
and scatter-adds instead of torch conv calls), so the two restatements check each other.
"""
import math
import sys
import types

import numpy as np
import torch

_noise_queue = []          # tensors handed out by tf.random.normal, in call order


def push_noise(*tensors):
    _noise_queue.extend(tensors)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x))


# ------------------------------------------------------------------------------------ ops
def _reduce(fn):
    def f(x, axis=None, keepdims=False):
        x = _t(x)
        if axis is None:
            return fn(x)
        return fn(x, dim=axis, keepdim=keepdims)
    return f


def reduce_min(x, axis=None):
    x = _t(x)
    return x.min() if axis is None else x.min(dim=axis).values


def reduce_max(x, axis=None):
    x = _t(x)
    return x.max() if axis is None else x.max(dim=axis).values


def reduce_variance(x, axis=None):
    x = _t(x)
    if axis is None:
        return ((x - x.mean()) ** 2).mean()
    return ((x - x.mean(dim=axis, keepdim=True)) ** 2).mean(dim=axis)


def reduce_std(x, axis=None):
    return torch.sqrt(reduce_variance(x, axis))


def divide_no_nan(a, b):
    a, b = _t(a), _t(b)
    safe = torch.where(b == 0, torch.ones_like(b), b)
    return torch.where(b == 0, torch.zeros_like(a / safe), a / safe)


def random_normal(shape, mean=0.0, stddev=1.0):
    shape = tuple(int(s) for s in shape)
    if _noise_queue:
        n = _noise_queue.pop(0)
        assert tuple(n.shape) == shape, (tuple(n.shape), shape)
        return mean + stddev * n
    return mean + stddev * torch.randn(shape)


def split(x, num_or_size_splits, axis=0):
    return list(torch.chunk(_t(x), num_or_size_splits, dim=axis))


class GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, target, sources):
        return list(torch.autograd.grad(target, list(sources), allow_unused=True))


# --------------------------------------------------------------------------------- layers
def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def _act(name):
    if name is None:
        return lambda v: v
    if name == "relu":
        return torch.relu
    raise NotImplementedError(name)


class _Layer:
    built = False

    @property
    def variables(self):
        return [v for v in (getattr(self, "kernel", None), getattr(self, "bias", None)) if v is not None]

    trainable_weights = variables


def _glorot(shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape) * 2 - 1) * lim).requires_grad_(True)


class InputSpec:
    def __init__(self, shape):
        self.shape = tuple(shape)


def Input(shape):
    return InputSpec(shape)


class Conv2D(_Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", activation=None):
        assert padding == "same"
        self.filters, self.k = filters, kernel_size
        self.s = strides[0] if isinstance(strides, (tuple, list)) else strides
        self.act = _act(activation)

    def build(self, shape):              # shape = (H, W, C)
        h, w, c = shape
        k = self.k
        self.kernel = _glorot((k, k, c, self.filters), k * k * c, k * k * self.filters)   # HWIO
        self.bias = torch.zeros(self.filters, requires_grad=True)
        self.input_shape = (None,) + tuple(shape)
        return (-(-h // self.s), -(-w // self.s), self.filters)

    def __call__(self, x):
        n, h, w, c = x.shape
        ho, pt, pb = _same_pad(h, self.k, self.s)
        wo, pl, pr = _same_pad(w, self.k, self.s)
        xp = torch.zeros((n, h + pt + pb, w + pl + pr, c), dtype=x.dtype)
        xp[:, pt:pt + h, pl:pl + w, :] = x
        y = torch.zeros((n, ho, wo, self.filters), dtype=x.dtype) + self.bias
        for kh in range(self.k):
            for kw in range(self.k):
                patch = xp[:, kh:kh + self.s * (ho - 1) + 1:self.s, kw:kw + self.s * (wo - 1) + 1:self.s, :]
                y = y + patch @ self.kernel[kh, kw]
        return self.act(y)


class Conv2DTranspose(_Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None):
        assert padding == "same"
        self.filters, self.k = filters, kernel_size
        self.s = strides[0] if isinstance(strides, (tuple, list)) else strides
        self.act = _act(activation)

    def build(self, shape):
        h, w, c = shape
        k = self.k
        # Keras Conv2DTranspose kernel: (kh, kw, out, in); fans use shape[-2], shape[-1] as in/out
        self.kernel = _glorot((k, k, self.filters, c), k * k * self.filters, k * k * c)
        self.bias = torch.zeros(self.filters, requires_grad=True)
        self.input_shape = (None,) + tuple(shape)
        return (h * self.s, w * self.s, self.filters)

    def __call__(self, x):
        # conv2d_transpose = gradient of conv2d wrt its input: with the forward conv mapping an
        # (s*H, s*W) image to (H, W) under SAME padding, out[s*i + kh - pad_before] += x[i] * W[kh]
        n, h, w, c = x.shape
        s, k = self.s, self.k
        oh, ow = h * s, w * s
        _, pt, _ = _same_pad(oh, k, s)
        _, pl, _ = _same_pad(ow, k, s)
        canvas = torch.zeros((n, s * (h - 1) + k, s * (w - 1) + k, self.filters), dtype=x.dtype)
        for kh in range(k):
            for kw in range(k):
                contrib = x @ self.kernel[kh, kw].transpose(0, 1)          # [n,h,w,in] @ [in,out]
                idx_h = slice(kh, kh + s * (h - 1) + 1, s)
                idx_w = slice(kw, kw + s * (w - 1) + 1, s)
                pad = torch.zeros_like(canvas)
                pad[:, idx_h, idx_w, :] = contrib
                canvas = canvas + pad
        y = canvas[:, pt:pt + oh, pl:pl + ow, :] + self.bias
        return self.act(y)


class Flatten(_Layer):
    def build(self, shape):
        return (int(np.prod(shape)),)

    def __call__(self, x):
        return x.reshape(x.shape[0], -1)


class Dense(_Layer):
    def __init__(self, units, activation=None):
        self.units = units
        self.act = _act(activation)

    def build(self, shape):
        (k,) = shape
        self.kernel = _glorot((k, self.units), k, self.units)
        self.bias = torch.zeros(self.units, requires_grad=True)
        self.input_shape = (None, k)
        return (self.units,)

    def __call__(self, x):
        return self.act(x @ self.kernel + self.bias)


class Reshape(_Layer):
    def __init__(self, target_shape):
        self.target_shape = tuple(target_shape)

    def build(self, shape):
        return self.target_shape

    def __call__(self, x):
        return x.reshape((x.shape[0],) + self.target_shape)


class Sequential:
    def __init__(self, layers, name=None):
        self.name = name
        assert isinstance(layers[0], InputSpec)
        shape = layers[0].shape
        self.layers = list(layers[1:])
        for l in self.layers:
            shape = l.build(tuple(shape))
        self.output_shape = (None,) + tuple(shape)

    def __call__(self, x, training=False):
        x = _t(x)
        for l in self.layers:
            x = l(x)
        return x

    @property
    def trainable_weights(self):
        return [v for l in self.layers for v in l.variables]

    def summary(self):
        pass


class Model:
    def __init__(self):
        self.optimizer = None

    def __call__(self, *a, **k):
        return self.call(*a, **k)

    @property
    def trainable_weights(self):
        return self.encoder.trainable_weights + self.decoder.trainable_weights

    trainable_variables = trainable_weights

    def compile(self, optimizer=None, **k):
        self.optimizer = optimizer


class Adam:
    """Keras optimizer_v2 Adam (TF < 2.11): epsilon 1e-7, lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
    p -= lr_t * m / (sqrt(v) + eps)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations, self.m, self.v = 0, {}, {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        lr_t = float(self.learning_rate) * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for g, p in grads_and_vars:
                if g is None:
                    continue
                m = self.m.setdefault(id(p), torch.zeros_like(p))
                v = self.v.setdefault(id(p), torch.zeros_like(p))
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr_t * m / (torch.sqrt(v) + self.eps))


def install():
    """Register the shim as `tensorflow` (and stubs for the plotting / dataset modules that
    do_anomaly_detection.py imports at module scope) in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "shim (torch-backed, tests/golden/tf_shim.py)"
    tf.Tensor = torch.Tensor
    tf.function = lambda f=None, **k: f if f is not None else (lambda g: g)
    tf.convert_to_tensor = _t
    tf.zeros = lambda shape, dtype=None: torch.zeros(tuple(int(s) for s in shape))
    tf.shape = lambda x: tuple(x.shape)
    tf.split = split
    tf.concat = lambda xs, axis=0: torch.cat([_t(x) for x in xs], dim=axis)
    tf.sigmoid = lambda x: torch.sigmoid(_t(x))
    tf.exp = lambda x: torch.exp(_t(x))
    tf.abs = lambda x: torch.abs(_t(x))
    tf.pow = lambda x, p: _t(x) ** p
    tf.reduce_sum = _reduce(torch.sum)
    tf.reduce_mean = _reduce(torch.mean)
    tf.reduce_min, tf.reduce_max = reduce_min, reduce_max
    tf.GradientTape = GradientTape
    tf.math = types.SimpleNamespace(
        log=lambda x: torch.log(_t(x) if not isinstance(x, float) else torch.tensor(x)),
        reduce_std=reduce_std, reduce_variance=reduce_variance, divide_no_nan=divide_no_nan,
        reduce_mean=_reduce(torch.mean), reduce_sum=_reduce(torch.sum), pow=lambda x, p: _t(x) ** p,
        sqrt=lambda x: torch.sqrt(_t(x)))
    tf.random = types.SimpleNamespace(normal=random_normal)
    tf.config = types.SimpleNamespace(list_physical_devices=lambda kind=None: [],
                                      experimental=types.SimpleNamespace(set_memory_growth=lambda *a: None))
    keras = types.SimpleNamespace(
        Model=Model, Sequential=Sequential,
        layers=types.SimpleNamespace(Input=Input, Conv2D=Conv2D, Conv2DTranspose=Conv2DTranspose, Flatten=Flatten,
                                     Dense=Dense, Reshape=Reshape),
        optimizers=types.SimpleNamespace(Adam=Adam),
        callbacks=types.SimpleNamespace(Callback=object),
        models=types.SimpleNamespace(load_model=None))
    tf.keras = keras
    sys.modules["tensorflow"] = tf
    from unittest import mock
    for name in ("cv2", "PIL", "PIL.Image", "matplotlib", "matplotlib.pyplot", "tensorflow_datasets", "tqdm"):
        if name == "tqdm":
            m = types.ModuleType("tqdm")
            m.tqdm = lambda it, **k: it
            sys.modules[name] = m
        elif name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    return tf
