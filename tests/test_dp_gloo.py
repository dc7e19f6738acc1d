"""Data-parallel host logic on CPU: world_size 2 over gloo.  Each rank runs the emulated
kernels on its shard; the library's all-reduce calls (latent/image moment sums, per-position
batch moments, min/max, flat gradient) are routed through torch.distributed.  The result must
equal the unsharded oracle on the full batch - batch-global kurtosis/skew included."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kcvae_testlib import O, emu_binding, eps_for, frames, make, rel_err, small_config


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


_CB_TYPE = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int, C.c_int)


def _allreduce_cb(buf, count, is_double, op):
    ctype = C.c_double if is_double else C.c_float
    arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(ctype)), shape=(count,))
    t = torch.from_numpy(arr)
    dist.all_reduce(t, op={0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}[op])


def _worker(rank, world, port, kind, tier, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = emu_binding()
        cb = _CB_TYPE(_allreduce_cb)
        lib.cdll.kcvae_emu_set_allreduce(cb)
        cfg = small_config(kind)
        Bl = 3
        m, ws = make(cfg, "emu", weight_gain=1.6, metrics=tier)
        m.distribute()
        x, eps = frames(cfg, Bl * world), eps_for(cfg, Bl * world)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        d, grads = m.loss_and_grads(x[sl], eps=eps[sl])
        q.put((rank, {k: float(v) for k, v in d.items()}, [g.copy() for g in grads]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,tier", [("global", "full"), ("single", "full"), ("global", "loss_only")])
def test_dp2_equals_unsharded_oracle(kind, tier):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, tier, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    cfg = small_config(kind)
    ws = O.glorot_init(cfg, 1234, bias_scale=0.05)
    ws = [w * 1.6 if w.ndim > 1 else w for w in ws]
    x, eps = frames(cfg, 6), eps_for(cfg, 6)
    od, ograds, _, _ = O.loss_and_grads(cfg, ws, x, eps, dtype=torch.float64)
    for rank, d, grads in res:
        for k, v in od.items():
            if tier == "loss_only" and k in ("x_std_loss", "cross_entropy", "r_min", "r_max"):
                continue
            assert abs(d[k] - float(v)) <= 1e-6 + 3e-4 * abs(float(v)), (rank, k, d[k], float(v))
        for g, og in zip(grads, ograds):
            assert rel_err(g, og.numpy()) < 3e-4
    # both ranks hold the same reduced gradient
    for a, b in zip(res[0][2], res[1][2]):
        np.testing.assert_array_equal(a, b)


# ---------------------------------------------------------------- per-rank noise streams
def _seed_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = emu_binding()
        cb = _CB_TYPE(_allreduce_cb)
        lib.cdll.kcvae_emu_set_allreduce(cb)
        cfg = small_config("global")
        m, _ = make(cfg, "emu", weight_gain=1.6)
        m.compile()
        m.seed(77)                      # the same user seed on every rank ...
        m.distribute()
        x = frames(cfg, 3)              # ... and the same frames: only the on-device eps can differ
        m.train_step(x)                 # eps=None -> Philox N(0,1) keyed by (seed, rank)
        q.put((rank, m.debug_activation(201).copy()))
    finally:
        dist.destroy_process_group()


def test_dp_ranks_draw_different_reparameterisation_noise():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_seed_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    z0, z1 = res[0][1], res[1][1]
    assert z0.shape == z1.shape and np.abs(z0 - z1).max() > 1e-2      # i.i.d. draws, not copies


# ---------------------------------------------------------------- sharded scoring: set-level statistics
def _stats_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        sc = importlib.import_module("trustedai-cl-vae-ad_b200.scoring")
        rng = np.random.default_rng(5)
        s = torch.from_numpy((16800.0 + 40.0 * rng.standard_normal(37)).astype(np.float32))   # near-equal scores, like uniform frames
        lo = torch.from_numpy(rng.random(37).astype(np.float32) * 1e-3)
        hi = torch.from_numpy(1.0 + rng.random(37).astype(np.float32))
        mine = slice(0, 20) if rank == 0 else slice(20, 37)                                  # ragged shards
        meu, sigma, emin, emax = sc.set_statistics(s[mine], lo[mine].min(), hi[mine].max(), distributed=True)
        q.put((rank, float(meu), float(sigma), float(emin), float(emax)))
    finally:
        dist.destroy_process_group()


def test_sharded_set_statistics_equal_unsharded():
    """do_anomaly_detection.py:63-71 over a frame set sharded across two ranks == the same numbers on the whole set."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    s = (16800.0 + 40.0 * rng.standard_normal(37)).astype(np.float32)
    lo = rng.random(37).astype(np.float32) * 1e-3
    hi = 1.0 + rng.random(37).astype(np.float32)
    want = (float(s.astype(np.float64).mean()), float(s.astype(np.float64).std()), float(lo.min()), float(hi.max()))
    for rank, meu, sigma, emin, emax in res:
        assert abs(meu - want[0]) <= 1e-6 * abs(want[0]) and abs(sigma - want[1]) <= 1e-5 * want[1]
        assert emin == np.float32(want[2]) and emax == np.float32(want[3])
