"""Fixtures produced by running the REFERENCE'S OWN PYTHON (/root/reference/src/*.py,
do_anomaly_detection.py) over the torch-backed TensorFlow shim (tests/golden/tf_shim.py,
generator tests/golden/make_reference_goldens.py).  They pin the reference's composition -
topology wiring, loss algebra, tape.gradient, the train_step/Adam sequence, the scoring loops -
for (1) the oracle (CPU, every run), (2) the kernel sources through the functional emulator (CPU)
and (3) the CUDA path through the C ABI (-m gpu).  Tolerances: north_star's 1e-3 relative on loss
terms and 1e-2 max-abs on reconstructions are the contract; the fp32 paths are held much tighter."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from kcvae_testlib import O, assert_metrics_close, model_class, pkg, rel_err

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "reference_*.npz")))
IDS = [os.path.basename(p)[len("reference_"):-4] for p in GOLD]


def load(path):
    g = np.load(path, allow_pickle=False)
    cfg = json.loads(str(g["config_json"]))
    n = int(g["n_weights"])
    return g, cfg, [g[f"w{i}"] for i in range(n)], [g[f"g{i}"] for i in range(n)], [g[f"w_after{i}"] for i in range(n)]


def as_dict(g, key):
    return dict(zip([str(k) for k in g["loss_keys"]], g[key].tolist()))


def test_fixtures_present():
    assert len(GOLD) >= 4, "tests/golden/reference_*.npz missing: run tests/golden/make_reference_goldens.py"


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_oracle_matches_reference_run(path):
    g, cfg, ws, grads, w_after = load(path)
    x, eps = g["x"], g["eps"]
    # variable order / shapes = Keras trainable_weights of the reference model
    assert [tuple(w.shape) for w in ws] == [tuple(s) for _, s in O.variable_shapes(cfg)]
    xh, z, mean, lv = O.call_detailed(cfg, ws, x, eps)
    np.testing.assert_allclose(mean.numpy(), g["mean"], atol=2e-5)
    np.testing.assert_allclose(lv.numpy(), g["logvar"], atol=2e-5)
    np.testing.assert_allclose(z.numpy(), g["z"], atol=3e-5)
    np.testing.assert_allclose(xh.numpy(), g["xhat"], atol=1e-5)
    np.testing.assert_allclose(O.call_detailed(cfg, ws, x, None)[0].numpy(), g["xhat_inference"], atol=1e-5)
    d0 = O.compute_loss(cfg, ws, x, None)[0]
    assert_metrics_close({k: float(v) for k, v in d0.items()}, as_dict(g, "loss_inference"), rtol=1e-4, atol=1e-6)
    d, ograds, _, _ = O.loss_and_grads(cfg, ws, x, eps)
    assert_metrics_close({k: float(v) for k, v in d.items()}, as_dict(g, "loss_train"), rtol=1e-4, atol=1e-6)
    for i, (og, rg) in enumerate(zip(ograds, grads)):
        assert rel_err(og.numpy(), rg) < 2e-4, f"gradient of variable {i}"
    om = O.OracleModel(cfg, ws)
    for s in range(3):
        ds, _ = om.train_step(x, g["step_eps"][s])
        assert_metrics_close({k: float(v) for k, v in ds.items()}, dict(zip(as_dict(g, "loss_train").keys(), g["step_loss"][s])),
                             rtol=5e-4, atol=2e-6)
    lr = float(cfg["training"]["learning_rate"])
    for i, (w, rw) in enumerate(zip(om.weights, w_after)):
        diff = np.abs(w.numpy() - rw)
        assert np.mean(diff) < 0.02 * lr, f"variable {i} after 3 Adam steps: mean |dw| = {np.mean(diff)}"
    # scoring loops with the trained weights
    omt = O.OracleModel(cfg, w_after)
    batches = [b for b in g["score_frames"]]
    sc = O.get_data_scale(omt, batches)
    for k in ("meu", "sigma", "min", "max"):
        assert abs(float(sc[k]) - float(g[f"scale_{k}"])) <= 1e-6 + 1e-4 * abs(float(g[f"scale_{k}"])), k
    np.testing.assert_allclose(sc["z_scores"].numpy(), g["scale_z_scores"], atol=2e-3)
    ev = O.evaluate_anomalies(omt, batches, sc, 1.0)
    np.testing.assert_allclose(ev["rec"], g["eval_rec"], atol=1e-5)
    np.testing.assert_allclose(ev["errs"], g["eval_errs"], atol=1e-5)
    np.testing.assert_allclose(ev["norm_errs"], g["eval_norm_errs"], atol=1e-4)
    np.testing.assert_array_equal(ev["anomalies"], g["eval_anomalies"])
    np.testing.assert_array_equal(np.argsort(-ev["z_scores"], kind="stable"), np.argsort(-g["eval_z_scores"], kind="stable"))


def metrics_close(got, want, rtol, atol, a_rec):
    """loss terms at the relative bar; r_min / r_max are reconstruction VALUES and get the (absolute)
    reconstruction bar"""
    assert_metrics_close(got, want, rtol=rtol, atol=atol, skip=("r_min", "r_max"))
    for k in ("r_min", "r_max"):
        assert abs(float(got[k]) - float(want[k])) <= a_rec, f"{k}: got {float(got[k])} want {float(want[k])}"


def check_backend(backend, path, precision=None):
    g, cfg, ws, grads, w_after = load(path)
    x, eps = g["x"], g["eps"]
    kind = "single" if cfg["model"].get("type") == "KurtosisSingle" else "global"
    kw = {"precision": precision} if precision else {}
    m = model_class(backend, kind)(cfg, **kw)
    m.set_weights(ws)
    tc = precision == "bf16" and m.tc_status() == 1     # tensor-core kernels active for this topology
    a_rec, r_loss, r_grad = (4e-3, 1e-3, 5e-2) if tc else (3e-5, 2e-4, 2e-4)
    xh, z, mean, lv = m.call_detailed(x, training=True, eps=eps)
    np.testing.assert_allclose(mean.numpy(), g["mean"], atol=2e-5)
    np.testing.assert_allclose(z.numpy(), g["z"], atol=3e-5)
    np.testing.assert_allclose(xh.numpy(), g["xhat"], atol=a_rec)                       # contract: 1e-2
    np.testing.assert_allclose(m.call(x, False).numpy(), g["xhat_inference"], atol=a_rec)
    metrics_close(m.compute_loss(x, training=False), as_dict(g, "loss_inference"), r_loss, 1e-6, a_rec)
    d, mg = m.loss_and_grads(x, eps=eps)
    metrics_close(d, as_dict(g, "loss_train"), r_loss, 1e-6, a_rec)                     # contract: 1e-3
    for i, (a, b) in enumerate(zip(mg, grads)):
        if tc:   # bf16 operands: relative L2 bar (as tests/test_gpu_parity.py), max-abs only loosely
            l2 = float(np.linalg.norm(np.asarray(a, np.float64) - b) / (np.linalg.norm(b) + 1e-30))
            assert l2 < r_grad and rel_err(a, b) < 3 * r_grad, f"gradient of variable {i}: L2 {l2}, max {rel_err(a, b)}"
        else:
            assert rel_err(a, b) < r_grad, f"gradient of variable {i}: {rel_err(a, b)}"
    m.compile(optimizer=pkg.Adam(learning_rate=float(cfg["training"]["learning_rate"])))
    keys = list(as_dict(g, "loss_train").keys())
    for s in range(3):
        ds = m.train_step(x, eps=g["step_eps"][s])
        # from the second step on the two sides no longer hold identical weights (bf16 gradients move them by
        # O(lr * 2^-9)), and |target - kurtosis| is a difference of nearly equal numbers: absolute slack there
        metrics_close(ds, dict(zip(keys, g["step_loss"][s])), 5e-3 if tc else 5e-4, 3e-4 if tc else 2e-6, a_rec)
    lr = float(cfg["training"]["learning_rate"])
    for i, (w, rw) in enumerate(zip(m.get_weights(), w_after)):
        assert np.mean(np.abs(w - rw)) < (0.1 if tc else 0.02) * lr, f"variable {i} after 3 Adam steps"
    m.set_weights(w_after)
    batches = [b for b in g["score_frames"]]
    sc = pkg.get_data_scale(m, cfg, {"train": batches})
    for k in ("meu", "sigma", "min", "max"):
        assert abs(float(sc[k]) - float(g[f"scale_{k}"])) <= 1e-5 + max(r_loss, 2e-4) * abs(float(g[f"scale_{k}"])), k
    ev = pkg.evaluate_anomalies(m, cfg, {"train": batches}, {k: g[f"scale_{k}"] for k in ("meu", "sigma", "min", "max")}, 1.0)
    np.testing.assert_allclose(ev["rec"], g["eval_rec"], atol=a_rec)
    np.testing.assert_allclose(ev["errs"], g["eval_errs"], atol=10 * a_rec)
    np.testing.assert_array_equal(ev["anomalies"], g["eval_anomalies"])
    # identical ranking (north_star)
    np.testing.assert_array_equal(np.argsort(-ev["z_scores"], kind="stable"), np.argsort(-g["eval_z_scores"], kind="stable"))


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_emulated_kernels_match_reference_run(path):
    check_backend("emu", path)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_cuda_matches_reference_run(path, precision):
    check_backend("cuda", path, precision)
