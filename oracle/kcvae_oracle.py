"""CPU oracle for the KurtosisCVAE hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU (torch-CPU / numpy, fp32 with an fp64 mode), the
algorithm the reference runs through TensorFlow/Keras.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline - never as part of the shipped path (the product path is
``trustedai-cl-vae-ad_b200`` -> ``libkcvae.so`` and fails loudly without it).

Parity pin status
-----------------
The arithmetic of the reference lives in TensorFlow<2.11 + Keras
(``env.yml:19,22``), a third-party dependency that is neither vendored under
/root/reference nor installable in this image.  The oracle is therefore pinned
against every *weight-independent* golden value the reference's own tests hold
(``tests/test_kurtosis_global_cvae.py:155-168``,
``tests/test_kurtosis_single_cvae.py:155-166``; see
``tests/test_oracle_golden.py``) and against definitional numpy loops of the
TF layer semantics (``tests/test_oracle_layers.py``).  The network-dependent
golden numbers (z_l1, kl_div, r_min/r_max, Single's z_kurtosis ...) depend on
TF's seeded Glorot draws and are **parity unpinned**.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

GLOBAL_KEYS = ["loss", "mse", "z_l1", "var_loss", "skew_loss", "z_kurtosis_loss",
               "z_kurtosis", "r_min", "r_max", "cross_entropy", "kl_div", "x_std_loss"]
SINGLE_KEYS = ["loss", "mse", "z_l1", "z_l2", "skew_loss", "z_kurtosis_loss",
               "z_kurtosis", "r_min", "r_max", "x_std_loss"]


# --------------------------------------------------------------------------
# topology  (src/abstract_cvae.py:22-92)
# --------------------------------------------------------------------------
@dataclass
class Topology:
    H: int
    W: int
    C: int
    layers: List[int]
    enc_dense: int          # 0 when encoder_dense_filters is absent/falsy
    dec_dense: int
    latent: int
    enc_hw: List[Tuple[int, int]]   # spatial size after each encoder conv (ceil)
    dec_h0: int
    dec_w0: int

    @property
    def flat(self) -> int:
        h, w = self.enc_hw[-1] if self.layers else (self.H, self.W)
        c = self.layers[-1] if self.layers else self.C
        return h * w * c


def topology(config: dict) -> Topology:
    """Shapes exactly as ``_build_encoder`` / ``_build_decoder`` derive them
    (src/abstract_cvae.py:30-45, 53-71).  Encoder halves with ceil (SAME, s2),
    the decoder starts from int(H / 2^L) and doubles L times."""
    H, W, C = [int(v) for v in config["data"]["image_size"]]
    layers = [int(f) for f in config["model"]["layers"]]
    enc_dense = config["model"].get("encoder_dense_filters")
    enc_dense = int(enc_dense) if enc_dense else 0
    dec_dense = int(config["model"]["decoder_dense_filters"])
    latent = int(config["model"]["latent_dimensions"])
    hw, h, w = [], H, W
    for _ in layers:
        h, w = (h + 1) // 2, (w + 1) // 2
        hw.append((h, w))
    n = len(layers)
    h0 = int(float(H) / float(2 ** n))
    w0 = int(float(W) / float(2 ** n))
    if h0 == 0:
        raise RuntimeError(f"Error: Build Decoder: Width Collapse: Too many layers, check configuration file: {H} -> {h0}: {n} Layers")
    if w0 == 0:
        raise RuntimeError(f"Error: Build Decoder: Height Collapse: Too many layers, check configuration file: {W} -> {w0}: {n} Layers")
    return Topology(H, W, C, layers, enc_dense, dec_dense, latent, hw, h0, w0)


def variable_shapes(config: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """``trainable_weights`` order and Keras layouts: Conv2D kernel HWIO,
    Conv2DTranspose kernel [kh,kw,out,in], Dense kernel [in,out]
    (SURVEY 8a row 1)."""
    t = topology(config)
    out: List[Tuple[str, Tuple[int, ...]]] = []
    cin = t.C
    for i, f in enumerate(t.layers):
        out.append((f"encoder/conv2d_{i}/kernel", (3, 3, cin, f)))
        out.append((f"encoder/conv2d_{i}/bias", (f,)))
        cin = f
    k = t.flat
    if t.enc_dense:
        out.append(("encoder/dense/kernel", (k, t.enc_dense)))
        out.append(("encoder/dense/bias", (t.enc_dense,)))
        k = t.enc_dense
    out.append(("encoder/dense_head/kernel", (k, 2 * t.latent)))
    out.append(("encoder/dense_head/bias", (2 * t.latent,)))
    units = t.dec_h0 * t.dec_w0 * t.dec_dense
    out.append(("decoder/dense/kernel", (t.latent, units)))
    out.append(("decoder/dense/bias", (units,)))
    cin = t.dec_dense
    for i, f in enumerate(reversed(t.layers)):
        out.append((f"decoder/conv2d_transpose_{i}/kernel", (3, 3, f, cin)))
        out.append((f"decoder/conv2d_transpose_{i}/bias", (f,)))
        cin = f
    out.append(("decoder/conv2d_transpose_out/kernel", (3, 3, t.C, cin)))
    out.append(("decoder/conv2d_transpose_out/bias", (t.C,)))
    return out


def glorot_init(config: dict, seed: int = 1234, bias_scale: float = 0.0) -> List[np.ndarray]:
    """Keras defaults: glorot_uniform kernels, zero biases.  Fans
    follow Keras' ``_compute_fans``: rank-4 -> shape[-2]*kh*kw, shape[-1]*kh*kw.
    ``bias_scale``>0 gives small random biases so bias paths are exercised in
    parity tests (not a reference behaviour)."""
    rng = np.random.default_rng(seed)
    ws = []
    for _, shp in variable_shapes(config):
        if len(shp) == 1:
            if bias_scale:
                ws.append((rng.standard_normal(shp) * bias_scale).astype(np.float32))
            else:
                ws.append(np.zeros(shp, np.float32))
            continue
        if len(shp) == 4:
            rf = shp[0] * shp[1]
            fan_in, fan_out = shp[2] * rf, shp[3] * rf
        else:
            fan_in, fan_out = shp
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        ws.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
    return ws


# --------------------------------------------------------------------------
# TF layer semantics
# --------------------------------------------------------------------------
def _same_pad(n_in: int, stride: int, k: int = 3) -> Tuple[int, int]:
    out = -(-n_in // stride)
    total = max((out - 1) * stride + k - n_in, 0)
    return total // 2, total - total // 2


def conv2d_s2_same(x: torch.Tensor, w_hwio: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Keras Conv2D(k3, s2, 'same') on NHWC (src/abstract_cvae.py:32); A1/A2:
    even sizes pad bottom/right only."""
    n, h, wd, c = x.shape
    pt, pb = _same_pad(h, 2)
    pl, pr = _same_pad(wd, 2)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), b, stride=2)
    return y.permute(0, 2, 3, 1)


def conv2dT_s2_same(x: torch.Tensor, w_hwoi: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Keras Conv2DTranspose(k3, s2, 'same') (src/abstract_cvae.py:83); A3:
    y[2i+kh, 2j+kw, co] += x[i,j,ci] W[kh,kw,co,ci], crop the END to 2*in."""
    n, h, wd, c = x.shape
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w_hwoi.permute(3, 2, 0, 1), b, stride=2)
    y = y[:, :, : 2 * h, : 2 * wd]
    return y.permute(0, 2, 3, 1)


def conv2dT_s1_same(x: torch.Tensor, w_hwoi: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Keras Conv2DTranspose(k3, s1, 'same') (src/abstract_cvae.py:88); A4."""
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w_hwoi.permute(3, 2, 0, 1), b, stride=1, padding=1)
    return y.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------
# forward (src/abstract_cvae.py:115-149)
# --------------------------------------------------------------------------
def _t(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def encoder_forward(t: Topology, ws: Sequence[torch.Tensor], x: torch.Tensor, keep: Optional[list] = None):
    """``self.encoder(x)`` then ``tf.split(.., 2, axis=1)`` (:120-121)."""
    i = 0
    a = x
    for _ in t.layers:
        a = torch.relu(conv2d_s2_same(a, ws[i], ws[i + 1]))
        i += 2
        if keep is not None:
            keep.append(a)
    a = a.reshape(a.shape[0], -1)          # Flatten of NHWC (A5)
    if t.enc_dense:
        a = a @ ws[i] + ws[i + 1]          # linear (:44)
        i += 2
        if keep is not None:
            keep.append(a)
    a = a @ ws[i] + ws[i + 1]              # linear (:45)
    i += 2
    L = t.latent
    return a[:, :L], a[:, L:], i


def decoder_forward(t: Topology, ws: Sequence[torch.Tensor], z: torch.Tensor, i: int, keep: Optional[list] = None):
    """``self.decoder(z)`` (:74-89) - returns logits NHWC."""
    a = torch.relu(z @ ws[i] + ws[i + 1])
    i += 2
    a = a.reshape(z.shape[0], t.dec_h0, t.dec_w0, t.dec_dense)
    if keep is not None:
        keep.append(a)
    for _ in t.layers:
        a = torch.relu(conv2dT_s2_same(a, ws[i], ws[i + 1]))
        i += 2
        if keep is not None:
            keep.append(a)
    return conv2dT_s1_same(a, ws[i], ws[i + 1])


def n_encoder_vars(t: Topology) -> int:
    return 2 * len(t.layers) + (2 if t.enc_dense else 0) + 2


def call_detailed(config: dict, weights, x, eps=None, img_noise=None, dtype=torch.float32,
                  keep: Optional[list] = None):
    """``call_detailed(x, training)`` (:139-144): encode (never noisy unless
    the caller injects ``img_noise``, Note A) -> z = mean + 0.5*logvar + eps
    (:124-129) -> sigmoid(decoder(z)) (:131-137).  ``eps=None`` <=> training
    False (eps = 0)."""
    t = topology(config)
    ws = [_t(w, dtype) for w in weights]
    xt = _t(x, dtype)
    if img_noise is not None:
        xt = xt + _t(img_noise, dtype)
    mean, logvar, i = encoder_forward(t, ws, xt, keep)
    z = mean + logvar * 0.5
    if eps is not None:
        z = z + _t(eps, dtype)
    logits = decoder_forward(t, ws, z, i, keep)
    return torch.sigmoid(logits), z, mean, logvar


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------
def _divide_no_nan(a, b):
    return torch.where(b == 0, torch.zeros_like(a), a / torch.where(b == 0, torch.ones_like(b), b))


def loss_global(x, x_hat, z, mean, logvar, lc: dict) -> Dict[str, torch.Tensor]:
    """``KurtosisGlobalCVAE.compute_loss_new`` (src/kurtosis_global_cvae.py:40-110)."""
    x_logit = torch.log(torch.exp(x) / torch.sum(torch.exp(x)))                 # :46
    ce = -torch.mean(x_hat * x_logit)                                           # :47
    mse = torch.mean((x - x_hat) ** 2)                                          # :50
    z_mean = torch.mean(z)                                                      # :57
    z_var = torch.var(z, unbiased=False)                                        # :59
    z_std = torch.sqrt(z_var)                                                   # :58
    zs = _divide_no_nan(z - z_mean, z_std)                                      # :60
    z_skew = torch.mean(zs ** 3)                                                # :61
    z_kurt = torch.mean(zs ** 4)                                                # :62
    x_std = torch.std(x, dim=0, unbiased=False)                                 # :64
    xh_std = torch.std(x_hat, dim=0, unbiased=False)                            # :65
    x_std_loss = torch.mean((x_std - xh_std) ** 2)                              # :66
    var_loss = torch.abs(1.0 - z_var)                                           # :73
    skew_loss = torch.abs(z_skew)                                               # :75
    kurt_loss = torch.abs(float(lc["kurtosis"]) - z_kurt)                       # :77
    kl = 0.5 * torch.sum(torch.abs(1.0 + logvar ** 2 - mean ** 2 - torch.exp(logvar ** 2)))  # :36-38
    z_l1 = torch.mean(torch.abs(z))                                             # :83
    loss = (float(lc["w_mse"]) * mse + float(lc["w_kurtosis"]) * kurt_loss
            + float(lc["w_skew"]) * skew_loss + float(lc["w_z_l1_reg"]) * z_l1)  # :91
    return {"loss": loss, "mse": mse, "z_l1": z_l1, "var_loss": var_loss, "skew_loss": skew_loss,
            "z_kurtosis_loss": kurt_loss, "z_kurtosis": z_kurt, "r_min": torch.min(x_hat),
            "r_max": torch.max(x_hat), "cross_entropy": ce, "kl_div": kl, "x_std_loss": x_std_loss}


def loss_single(x, x_hat, z, lc: dict) -> Dict[str, torch.Tensor]:
    """``KurtosisSingleCVAE.compute_loss`` (src/kurtosis_single_cvae.py:25-77):
    per-latent-dimension moments across the batch axis."""
    mse = torch.mean((x - x_hat) ** 2)                                          # :31
    x_std = torch.std(x, dim=0, unbiased=False)                                 # :34
    xh_std = torch.std(x_hat, dim=0, unbiased=False)
    x_std_loss = torch.mean((x_std - xh_std) ** 2)                              # :36
    z_meu = torch.mean(z, dim=0)                                                # :39
    z_std = torch.std(z, dim=0, unbiased=False)                                 # :40
    zs = _divide_no_nan(z - z_meu, z_std.expand_as(z))                          # :41
    z_skew = torch.mean(zs ** 3, dim=0)                                         # :43
    z_kurt = torch.mean(zs ** 4, dim=0)                                         # :44
    kurt_loss = torch.mean((z_kurt - float(lc["kurtosis"])) ** 2)               # :47
    skew_loss = torch.mean(z_skew ** 2)                                         # :48
    z_l2 = torch.sqrt(torch.sum(z_meu ** 2))                                    # :51
    z_l1 = torch.mean(torch.abs(z))                                             # :54
    loss = (float(lc["w_mse"]) * mse + float(lc["w_kurtosis"]) * kurt_loss
            + float(lc["w_skew"]) * skew_loss + float(lc["w_z_l1_reg"]) * z_l2)  # :56-60
    return {"loss": loss, "mse": mse, "z_l1": z_l1, "z_l2": z_l2, "skew_loss": skew_loss,
            "z_kurtosis_loss": kurt_loss, "z_kurtosis": torch.sqrt(torch.mean(z_kurt ** 2)),
            "r_min": torch.min(x_hat), "r_max": torch.max(x_hat), "x_std_loss": x_std_loss}


def model_type(config: dict) -> str:
    """``import_vae_based_on_type`` (src/load_model.py:9-31)."""
    ty = config["model"].get("type")
    if ty is None:
        return "global"
    avail = ["KLGaussian", "KurtosisGlobal", "KurtosisSingle"]
    if ty not in avail:
        raise Exception(f"Error, type {ty} not found in available types: {avail}")
    if ty.lower() == "klgaussian":
        raise NotImplementedError("KLGaussian not yet implemented")
    return "global" if ty.lower() == "kurtosisglobal" else "single"


def compute_loss(config: dict, weights, x, eps=None, img_noise=None, dtype=torch.float32):
    """``compute_loss(x, training, return_inf=True)``; returns (dict, x_hat, z, mean, logvar)."""
    x_hat, z, mean, logvar = call_detailed(config, weights, x, eps, img_noise, dtype)
    xt = _t(x, dtype)
    if model_type(config) == "global":
        d = loss_global(xt, x_hat, z, mean, logvar, config["loss"])
    else:
        d = loss_single(xt, x_hat, z, config["loss"])
    return d, x_hat, z, mean, logvar


def loss_and_grads(config: dict, weights, x, eps=None, img_noise=None, dtype=torch.float32):
    """``tape.gradient(loss['loss'], trainable_weights)`` (src/abstract_cvae.py:156-160)."""
    ws = [_t(w, dtype).clone().requires_grad_(True) for w in weights]
    d, x_hat, z, mean, logvar = compute_loss(config, ws, x, eps, img_noise, dtype)
    grads = torch.autograd.grad(d["loss"], ws, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(w) for g, w in zip(grads, ws)]
    return ({k: v.detach() for k, v in d.items()}, [g.detach() for g in grads],
            x_hat.detach(), z.detach())


# --------------------------------------------------------------------------
# optimizer: tf.keras.optimizers.Adam (TF<2.11 optimizer_v2), train.py:99-101
# --------------------------------------------------------------------------
class Adam:
    """Keras optimizer_v2 Adam (third-party; formula per its published
    ``_resource_apply_dense``): t = iterations+1;
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    p -= lr_t * m / (sqrt(v) + eps), eps=1e-7."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.b1, self.b2, self.eps = float(learning_rate), beta_1, beta_2, epsilon
        self.iterations = 0
        self.m: Optional[List[torch.Tensor]] = None
        self.v: Optional[List[torch.Tensor]] = None

    def apply_gradients(self, grads, params: List[torch.Tensor]):
        if self.m is None:
            self.m = [torch.zeros_like(p) for p in params]
            self.v = [torch.zeros_like(p) for p in params]
        self.iterations += 1
        t = self.iterations
        lr_t = self.learning_rate * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            p.sub_(lr_t * m / (torch.sqrt(v) + self.eps))


class OracleModel:
    """Stateful CPU model used as the checker and as the timed CPU baseline:
    mirrors ``train_step`` / ``test_step`` / ``call`` of the reference."""

    def __init__(self, config: dict, weights=None, dtype=torch.float32, seed: int = 1234):
        self.config = config
        self.dtype = dtype
        ws = weights if weights is not None else glorot_init(config, seed)
        self.weights = [_t(w, dtype).clone() for w in ws]
        self.optimizer = Adam(float(config["training"]["learning_rate"]))

    def train_step(self, x, eps):
        d, grads, x_hat, _ = loss_and_grads(self.config, self.weights, x, eps, None, self.dtype)
        self.optimizer.apply_gradients(grads, self.weights)
        return d, x_hat

    def test_step(self, x):
        with torch.no_grad():
            return compute_loss(self.config, self.weights, x, None, None, self.dtype)[0]

    def call(self, x):
        with torch.no_grad():
            return call_detailed(self.config, self.weights, x, None, None, self.dtype)[0]


# --------------------------------------------------------------------------
# anomaly scoring (do_anomaly_detection.py:57-117)
# --------------------------------------------------------------------------
def error_map(x: torch.Tensor, x_rec: torch.Tensor) -> torch.Tensor:
    """``tf.reduce_sum(tf.pow(batch - x_rec, 2), axis=3)`` (:62, :88)."""
    return torch.sum((x - x_rec) ** 2, dim=3)


def get_data_scale(model: OracleModel, batches) -> Dict[str, torch.Tensor]:
    """``get_data_scale`` (:57-79)."""
    errs = []
    for b in batches:
        xt = _t(b, model.dtype)
        errs.append(error_map(xt, model.call(xt)))
    err_vec = torch.cat(errs, 0)
    red = err_vec.sum(dim=2).sum(dim=1)
    meu = red.mean()
    sigma = torch.std(red, unbiased=False)
    return {"meu": meu, "sigma": sigma, "min": err_vec.min(), "max": err_vec.max(),
            "z_scores": (red - meu) / sigma}


def evaluate_anomalies(model: OracleModel, batches, scale, threshold: float = 3.0):
    """``evaluate_anomalies`` (:82-117)."""
    recs, errs, zs, norms = [], [], [], []
    for b in batches:
        xt = _t(b, model.dtype)
        rec = model.call(xt)
        err = error_map(xt, rec)
        red = err.sum(dim=2).sum(dim=1)
        zs.append((red - scale["meu"]) / scale["sigma"])
        norms.append((err - scale["min"]) / (scale["max"] - scale["min"]))
        recs.append(rec)
        errs.append(err)
    z = torch.cat(zs, 0)
    return {"rec": torch.cat(recs, 0).numpy(), "errs": torch.cat(errs, 0).numpy(),
            "z_scores": z.numpy(), "norm_errs": torch.cat(norms, 0).numpy(),
            "anomalies": z.numpy() > threshold}


# --------------------------------------------------------------------------
# configs used by tests / bench (README.md:52-85; tests/test_kurtosis_global_cvae.py:30-56)
# --------------------------------------------------------------------------
def readme_config(model_type_: Optional[str] = None) -> dict:
    cfg = {
        "data": {"dataset": "synthetic", "image_size": [224, 300, 3],
                 "train_split": "train_labels.json", "val_split": "val_labels.json"},
        "loss": {"kurtosis": 3.0, "w_kl_divergence": 0.0, "w_kurtosis": 1e-3, "w_mse": 1.0,
                 "w_skew": 0.0, "w_x_std": 1e-10, "w_z_l1_reg": 1e-3},
        "model": {"encoder_dense_filters": 16, "decoder_dense_filters": 32,
                  "latent_dimensions": 32, "layers": [32, 5]},
        "training": {"batch_size": 16, "beta": 1e-6, "learning_rate": 1e-4, "max_epochs": 1000},
    }
    if model_type_:
        cfg["model"]["type"] = model_type_
    return cfg


def unit_test_config(model_type_: Optional[str] = None) -> dict:
    cfg = {
        "data": {"image_size": [224, 300, 3]},
        "loss": {"kurtosis": 3.0, "w_kl_divergence": 0.0, "w_kurtosis": 1e-3, "w_mse": 1.0,
                 "w_skew": 0.0, "w_x_std": 1e-10, "w_z_l1_reg": 1e-3},
        "model": {"decoder_dense_filters": 4, "encoder_dense_filters": 4,
                  "latent_dimensions": 2, "layers": [5, 5]},
        "training": {"batch_size": 16, "beta": 1e-6, "learning_rate": 1e-4, "max_epochs": 10},
    }
    if model_type_:
        cfg["model"]["type"] = model_type_
    return cfg


def scaled_config() -> dict:
    """BASELINE.json config 5 instance pinned by SURVEY 8d."""
    cfg = readme_config()
    cfg["data"]["image_size"] = [448, 600, 3]
    cfg["model"].update({"layers": [64, 128, 32], "encoder_dense_filters": 64,
                         "decoder_dense_filters": 64, "latent_dimensions": 256})
    return cfg


def synthetic_frames(batch: int, config: dict, seed: int = 42) -> np.ndarray:
    """SURVEY 8d synthetic inputs: uniform [0,1) float32 NHWC."""
    H, W, C = config["data"]["image_size"]
    return np.random.default_rng(seed).random((batch, H, W, C), dtype=np.float32)


def synthetic_eps(batch: int, config: dict, step: int = 0) -> np.ndarray:
    L = int(config["model"]["latent_dimensions"])
    return np.random.default_rng(7 + step).standard_normal((batch, L)).astype(np.float32)
