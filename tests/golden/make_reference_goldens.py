#!/usr/bin/env python3
"""Generate tests/golden/reference_*.npz by EXECUTING THE REFERENCE'S OWN PYTHON
(/root/reference/src/*.py and do_anomaly_detection.py) over the torch-backed TensorFlow shim in
tf_shim.py.  Run in the build container (where /root/reference exists):

    python tests/golden/make_reference_goldens.py

Nothing here is read at test time except the .npz files it writes; the GPU box has no
/root/reference.  Each fixture holds the config, the injected weights (Keras trainable_weights
order and layouts), inputs, noise draws and what the reference code returned:
  call_detailed outputs, the compute_loss dict (training=True with the injected eps), the
  gradients tape.gradient produced, the weights after three train_step calls with
  tf.keras.optimizers.Adam, and get_data_scale / evaluate_anomalies results on a 3-batch dataset.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tf_shim  # noqa: E402

tf = tf_shim.install()
REF = os.environ.get("KCVAE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from src.load_model import load_model_from_config  # noqa: E402  (the reference's factory)
import do_anomaly_detection as ref_ad  # noqa: E402  (the reference's scoring loops)


def config(kind, H, W, C, layers, enc, dec, latent):
    cfg = {
        "data": {"image_size": [H, W, C]},
        "loss": {"kurtosis": 3.0, "w_kl_divergence": 0.0, "w_kurtosis": 1e-2, "w_mse": 1.0, "w_skew": 0.05,
                 "w_x_std": 1e-10, "w_z_l1_reg": 1e-2},
        "model": {"decoder_dense_filters": dec, "latent_dimensions": latent, "layers": list(layers)},
        "training": {"batch_size": 4, "beta": 1e-2, "learning_rate": 1e-3, "max_epochs": 1},
    }
    if enc:
        cfg["model"]["encoder_dense_filters"] = enc
    if kind == "single":
        cfg["model"]["type"] = "KurtosisSingle"
    return cfg


def set_weights(model, rng, gain):
    ws = []
    for v in model.trainable_weights:
        shape = tuple(v.shape)
        if len(shape) > 1:
            rf = int(np.prod(shape[:-2]))
            lim = gain * np.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
            a = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            a = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        with torch.no_grad():
            v.copy_(torch.from_numpy(a))
        ws.append(a)
    return ws


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def make(name, kind, B, **shape):
    rng = np.random.default_rng(20240 + len(name))
    torch.manual_seed(0)
    cfg = config(kind, **shape)
    model = load_model_from_config(cfg)
    assert type(model).__name__ == ("KurtosisSingleCVAE" if kind == "single" else "KurtosisGlobalCVAE")
    ws = set_weights(model, rng, gain=1.6)
    L = shape["latent"]
    x = rng.random((B, shape["H"], shape["W"], shape["C"]), dtype=np.float32)
    eps = rng.standard_normal((B, L)).astype(np.float32)
    out = {"config_json": np.array(json.dumps(cfg)), "x": x, "eps": eps, "n_weights": np.array(len(ws))}
    for i, w in enumerate(ws):
        out[f"w{i}"] = w
    xt = torch.from_numpy(x)

    # call_detailed(x, training=True): the reference draws eps with tf.random.normal -> injected
    tf_shim.push_noise(torch.from_numpy(eps))
    with torch.no_grad():
        xh, z, mean, logvar = model.call_detailed(xt, training=True)
    out.update(xhat=npy(xh), z=npy(z), mean=npy(mean), logvar=npy(logvar))
    with torch.no_grad():
        out["xhat_inference"] = npy(model.call(xt, training=False))
        d0 = model.compute_loss(xt, training=False)
    out["loss_keys"] = np.array(list(d0.keys()))
    out["loss_inference"] = np.array([float(v) for v in d0.values()], np.float64)

    # compute_loss(training=True) + tape.gradient, exactly the body of train_step
    tf_shim.push_noise(torch.from_numpy(eps))
    with tf.GradientTape() as tape:
        d = model.compute_loss(xt, training=True)
    grads = tape.gradient(d["loss"], model.trainable_weights)
    out["loss_train"] = np.array([float(v) for v in d.values()], np.float64)
    for i, g in enumerate(grads):
        out[f"g{i}"] = npy(g)

    # three train_step calls with Keras Adam (train.py:99-101), fresh noise per step
    model.compile(optimizer=tf.keras.optimizers.Adam(learning_rate=float(cfg["training"]["learning_rate"])))
    step_eps, step_loss = [], []
    for s in range(3):
        e = rng.standard_normal((B, L)).astype(np.float32)
        step_eps.append(e)
        tf_shim.push_noise(torch.from_numpy(e))
        ds = model.train_step(xt)
        step_loss.append([float(v) for v in ds.values()])
    out["step_eps"] = np.stack(step_eps)
    out["step_loss"] = np.array(step_loss, np.float64)
    for i, v in enumerate(model.trainable_weights):
        out[f"w_after{i}"] = npy(v)

    # scoring loops of do_anomaly_detection.py on a 3-batch dataset (weights after training)
    batches = [torch.from_numpy(rng.random((B, shape["H"], shape["W"], shape["C"]), dtype=np.float32)) for _ in range(3)]
    batches[1][0, 2:6, 3:9, :] = 1.0          # a planted bright patch: one clear anomaly
    data = {"train": batches}
    with torch.no_grad():
        scale = ref_ad.get_data_scale(model, cfg, data)
        ev = ref_ad.evaluate_anomalies(model, cfg, data, scale, 1.0)
    out["score_frames"] = np.stack([npy(b) for b in batches])
    for k in ("meu", "sigma", "min", "max", "z_scores"):
        out[f"scale_{k}"] = npy(scale[k])
    for k in ("rec", "errs", "z_scores", "norm_errs", "anomalies"):
        out[f"eval_{k}"] = npy(ev[k])
    path = os.path.join(HERE, f"reference_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, loss {dict(zip(d.keys(), out['loss_train'].round(6)))}")


if __name__ == "__main__":
    make("global_small", "global", 4, H=16, W=24, C=3, layers=(6, 5), enc=7, dec=4, latent=5)
    make("single_small", "single", 6, H=16, W=24, C=3, layers=(6, 5), enc=7, dec=4, latent=5)
    make("global_readme_channels", "global", 3, H=16, W=24, C=3, layers=(32, 5), enc=16, dec=32, latent=32)
    make("global_noenc", "global", 2, H=8, W=12, C=3, layers=(4,), enc=0, dec=3, latent=4)
