"""Anomaly scoring loops of do_anomaly_detection.py:57-117 on the CUDA scorer.

Same function names, argument order and returned keys as the reference; the per-batch
work (forward, per-pixel error, per-frame sum, per-frame min/max) is one kcvae_score call,
the set-level statistics are a handful of scalars."""
import ctypes as C

import numpy as np
import torch

from .model import _ptr, _wrap


def _iter(data):
    return data['train'] if isinstance(data, dict) else data


def set_statistics(err_reduced: torch.Tensor, emin: torch.Tensor, emax: torch.Tensor, distributed: bool = False):
    """meu, sigma (population), min, max over the WHOLE frame set (do_anomaly_detection.py:63-71).  With the frames
    sharded over ranks the set-level numbers come from one all-reduce of [sum s, sum s^2, n] (fp64)
    and one MIN all-reduce of [min, -max]: every rank gets the statistics of the unsharded set."""
    s64 = err_reduced.double()
    acc = torch.stack([s64.sum(), (s64 * s64).sum(), torch.tensor(float(s64.numel()), dtype=torch.float64, device=s64.device)])
    mm = torch.stack([emin.double(), -emax.double()])
    if distributed:
        import torch.distributed as dist
        if dist.get_backend() != "nccl":
            acc, mm = acc.cpu(), mm.cpu()
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        dist.all_reduce(mm, op=dist.ReduceOp.MIN)
        acc, mm = acc.to(s64.device), mm.to(s64.device)
    n = acc[2]
    meu = acc[0] / n
    # two-pass form on the local shard around the GLOBAL mean keeps fp32-grade scores from cancelling
    dev = ((s64 - meu) ** 2).sum()
    if distributed:
        import torch.distributed as dist
        d = dev.reshape(1).cpu() if dist.get_backend() != "nccl" else dev.reshape(1)
        dist.all_reduce(d, op=dist.ReduceOp.SUM)
        dev = d.to(s64.device)[0]
    sigma = torch.sqrt(dev / n)
    return meu.float(), sigma.float(), mm[0].float(), (-mm[1]).float()


def get_data_scale(model, config: dict, data, distributed=None):
    """do_anomaly_detection.py:57-79.  Unlike the reference it does not keep every error map
    in memory: min/max come from the per-frame (min,max) pairs the kernel emits.
    distributed (default: the model was attached with distribute()): `data` is this rank's shard of the frame set;
    meu / sigma / min / max are those of the whole set, z_scores are this rank's frames."""
    scores, mins, maxs = [], [], []
    for batch in _iter(data):
        r = model.score(batch, return_err=False)
        scores.append(r['score'])
        mins.append(r['err_minmax'][:, 0])
        maxs.append(r['err_minmax'][:, 1])
    err_reduced = torch.cat(scores)
    if distributed is None:
        distributed = getattr(model, "_dist_world", 1) > 1
    meu, sigma, emin, emax = set_statistics(err_reduced, torch.cat(mins).min(), torch.cat(maxs).max(), bool(distributed))
    return {
        'meu': _wrap(meu), 'sigma': _wrap(sigma),            # tf.math.reduce_std: population
        'min': _wrap(emin), 'max': _wrap(emax),
        'z_scores': _wrap((err_reduced - meu) / sigma),
    }


def evaluate_anomalies(model, config: dict, data, data_scale: dict, anomaly_threshold: float, keep_rec: bool = True):
    """do_anomaly_detection.py:82-117."""
    meu, sigma = float(data_scale['meu']), float(data_scale['sigma'])
    emin, emax = float(data_scale['min']), float(data_scale['max'])
    recs, errs, zs, norms, flags = [], [], [], [], []
    lib, h = model._lib, model._h
    for batch in _iter(data):
        r = model.score(batch, return_err=True, return_rec=keep_rec)
        B = r['score'].shape[0]
        norm = torch.empty_like(r['err'])
        z = torch.empty_like(r['score'])
        fl = torch.empty(B, dtype=torch.uint8, device=z.device)
        lib.check(lib.normalize_scores(h, _ptr(r['err']), _ptr(r['score']), B, meu, sigma, emin, emax,
                                       float(anomaly_threshold), _ptr(norm), _ptr(z), _ptr(fl), model._stream()), h)
        if keep_rec:
            recs.append(r['rec'])
        errs.append(r['err']); zs.append(z); norms.append(norm); flags.append(fl)
    cat = lambda v: torch.cat(v, 0).cpu().numpy()
    return {
        'rec': cat(recs) if keep_rec else None,
        'errs': cat(errs),
        'z_scores': cat(zs),
        'norm_errs': cat(norms),
        'anomalies': cat(flags).astype(bool),
    }


def output_anomalies(evaluation_data, anomaly_results: dict, data_scale: dict, output_path=None, anomaly_threshold: float = 3.0,
                     filenames=None, model=None):
    """do_anomaly_detection.py:118-198 without the plotting (and without the stray ``exit()`` at :157 that makes
    the reference stop before it writes anything): uint8 error image, JET heat map, overlay and reconstruction
    per frame (one kernel per batch) and the descending-z ranking.  Returns the arrays and the ranked
    ``(name, z_score)`` rows; writes PNGs + ``anomaly_list.csv`` under ``output_path`` when given."""
    import csv
    import os
    from .streaming import render_outputs
    binding = model._lib if model is not None else None
    r = render_outputs(anomaly_results['norm_errs'], anomaly_results['rec'], binding=binding)
    out = {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}
    n = out['err'].shape[0]
    names = list(filenames) if filenames is not None else [f'{i:06d}.png' for i in range(n)]
    order = rank_anomalies(anomaly_results['z_scores'])
    out['ranking'] = [(names[i], float(anomaly_results['z_scores'][i])) for i in order]
    if output_path is not None:
        from PIL import Image
        for sub in ('err', 'heatmap', 'overlay', 'rec'):
            os.makedirs(os.path.join(output_path, sub), exist_ok=True)
        for i in range(n):
            Image.fromarray(out['err'][i], mode='L').save(os.path.join(output_path, 'err', names[i]))
            for sub in ('heatmap', 'overlay', 'rec'):
                Image.fromarray(out[sub][i], mode='RGB').save(os.path.join(output_path, sub, names[i]))
        with open(os.path.join(output_path, 'anomaly_list.csv'), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['orig_filepath', 'z_score'])
            w.writerows(out['ranking'])
    return out


def rank_anomalies(z_scores: np.ndarray) -> np.ndarray:
    """Descending z-score order (do_anomaly_detection.py:190, dead code after exit() at :157)."""
    return np.argsort(-np.asarray(z_scores), kind='stable')
