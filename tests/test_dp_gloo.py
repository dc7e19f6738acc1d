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
