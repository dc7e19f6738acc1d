"""Data-parallel attachment: one process per GPU, torch.distributed for the rendezvous,
the library's own NCCL communicator for the hot-path collectives."""
import ctypes as C

import torch
import torch.distributed as dist


def attach(model):
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib, h = model._lib, model._h
    ident = (C.c_ubyte * 128)()
    if rank == 0:
        rc = lib.comm_unique_id(ident)
        if rc < 0:
            lib.check(rc, None)
    dev = model.device if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    buf = (C.c_ubyte * 128).from_buffer_copy(raw)
    lib.check(lib.comm_init(h, buf, rank, world), h)
    model._dist_world = world
    if model._dev_type == "cuda":
        lib.check(lib.broadcast_weights(h, 0, model._stream()), h)
    return model
