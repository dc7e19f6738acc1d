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


CFG_KW = dict(H=32, W=60, layers=(32, 5), enc=8, dec=8, latent=8)


def _worker(rank, world, port, kind, precision, steps, q):
    import torch.distributed as dist
    from kcvae_testlib import eps_for, frames, make, pkg, small_config
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = small_config(kind, **CFG_KW)
        Bl = 3
        m, ws = make(cfg, "cuda", weight_gain=1.6, device=rank, precision=precision)
        m.distribute()
        sl = slice(rank * Bl, (rank + 1) * Bl)
        if steps == 0:
            x, eps = frames(cfg, Bl * world), eps_for(cfg, Bl * world)
            d, grads = m.loss_and_grads(x[sl], eps=eps[sl])
            q.put((rank, {k: float(v) for k, v in d.items()}, [g.copy() for g in grads], int(m.tc_status())))
        else:                        # multi-step training with fixed noise: weights and metrics after `steps` Adam steps
            m.compile(optimizer=pkg.Adam(learning_rate=1e-3))
            for s in range(steps):
                x, eps = frames(cfg, Bl * world, seed=42 + s), eps_for(cfg, Bl * world, step=s)
                d = m.train_step(x[sl], eps=eps[sl])
            q.put((rank, {k: float(v) for k, v in d.items()}, [w.copy() for w in m.get_weights()], int(m.tc_status())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_world(kind, precision, steps):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, precision, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("kind", ["global", "single"])
def test_dp2_nccl_equals_unsharded_oracle(kind):
    """fp32 path, Global ([6 + 4] moment sums) and Single ([6 + 4 L] per-dimension moment sums) on NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from kcvae_testlib import O, eps_for, frames, rel_err, small_config
    res = _run_world(kind, "fp32", 0)
    cfg = small_config(kind, **CFG_KW)
    ws = O.glorot_init(cfg, 1234, bias_scale=0.05)
    ws = [w * 1.6 if w.ndim > 1 else w for w in ws]
    x, eps = frames(cfg, 6), eps_for(cfg, 6)
    od, ograds, _, _ = O.loss_and_grads(cfg, ws, x, eps, dtype=torch.float64)
    for rank, d, grads, _tc in res:
        for k, v in od.items():
            assert abs(d[k] - float(v)) <= 1e-6 + 3e-4 * abs(float(v)), (rank, k, d[k], float(v))
        for g, og in zip(grads, ograds):
            assert rel_err(g, og.numpy()) < 3e-4
    for a, b in zip(res[0][2], res[1][2]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("kind", ["global", "single"])
def test_dp2_nccl_tensor_core_training_matches_one_gpu(kind):
    """Default (tensor-core) precision, three train_steps with fixed eps on 2 GPUs (side stream + NCCL stream live) ==
    the same three steps of the concatenated batch on ONE GPU: weights to 2e-3 of max|w|, loss terms to 1e-3."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from kcvae_testlib import eps_for, frames, make, pkg, rel_err, small_config
    steps = 3
    res = _run_world(kind, "bf16", steps)
    cfg = small_config(kind, **CFG_KW)
    m, _ = make(cfg, "cuda", weight_gain=1.6, precision="bf16")
    m.compile(optimizer=pkg.Adam(learning_rate=1e-3))
    for s in range(steps):
        d = m.train_step(frames(cfg, 6, seed=42 + s), eps=eps_for(cfg, 6, step=s))
    want_w = m.get_weights()
    for rank, dd, ws, tc in res:
        assert tc == int(m.tc_status())
        for k, v in d.items():
            assert abs(dd[k] - float(v)) <= 1e-5 + 1e-3 * abs(float(v)), (rank, k, dd[k], float(v))
        for a, b in zip(ws, want_w):
            assert rel_err(a, b) < 2e-3
    for a, b in zip(res[0][2], res[1][2]):
        np.testing.assert_array_equal(a, b)       # replicas stay bit-identical
