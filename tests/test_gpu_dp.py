"""2-GPU data-parallel parity (-m gpu, needs >= 2 GPUs; skipped otherwise): each rank trains on
its shard through libkcvae.so with the library's NCCL communicator; metrics and the reduced
gradient must equal the unsharded oracle on the full batch (batch-global kurtosis included)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from kcvae_testlib import eps_for, frames, make, small_config
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = small_config(H=32, W=60, layers=(32, 5), enc=8, dec=8, latent=8)
        Bl = 3
        m, ws = make(cfg, "cuda", weight_gain=1.6, device=rank)
        m.distribute()
        x, eps = frames(cfg, Bl * world), eps_for(cfg, Bl * world)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        d, grads = m.loss_and_grads(x[sl], eps=eps[sl])
        q.put((rank, {k: float(v) for k, v in d.items()}, [g.copy() for g in grads]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_dp2_nccl_equals_unsharded_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from kcvae_testlib import O, eps_for, frames, rel_err, small_config
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    cfg = small_config(H=32, W=60, layers=(32, 5), enc=8, dec=8, latent=8)
    ws = O.glorot_init(cfg, 1234, bias_scale=0.05)
    ws = [w * 1.6 if w.ndim > 1 else w for w in ws]
    x, eps = frames(cfg, 6), eps_for(cfg, 6)
    od, ograds, _, _ = O.loss_and_grads(cfg, ws, x, eps, dtype=torch.float64)
    for rank, d, grads in res:
        for k, v in od.items():
            assert abs(d[k] - float(v)) <= 1e-6 + 3e-4 * abs(float(v)), (rank, k, d[k], float(v))
        for g, og in zip(grads, ograds):
            assert rel_err(g, og.numpy()) < 3e-4
    for a, b in zip(res[0][2], res[1][2]):
        np.testing.assert_array_equal(a, b)
