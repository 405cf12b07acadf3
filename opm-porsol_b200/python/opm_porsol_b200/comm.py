"""One process per GPU: wiring of the halo exchange between ranks.

The data path is entirely on the devices (peer-to-peer stores over NVLink into IPC-mapped buffers, see
csrc/eu_api.cu: halo_exchange).  torch.distributed is plumbing: it all-gathers one opaque blob per rank
(IPC handles + ghost lists) and serves the tiny min/max reductions of the CFL times and the range flag.
"""
import ctypes as C

import numpy as np

from .binding import EU_OK, EulerB200Error

_ALLREDUCE = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int)


def connect_ranks(dev, dist, device=None):
    """dev: EulerUpstream after initObj*/eu_grid_end; dist: an initialised torch.distributed module."""
    import torch
    L, h = dev.L, dev.h
    L.eu_comm_blob_size.argtypes = [C.c_void_p]
    L.eu_comm_export.argtypes = [C.c_void_p, C.c_void_p]
    L.eu_comm_connect.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.eu_comm_set_allreduce.argtypes = [C.c_void_p, _ALLREDUCE, C.c_void_p]
    size = L.eu_comm_blob_size(h)
    buf = C.create_string_buffer(size)
    if L.eu_comm_export(h, buf) != EU_OK:
        raise EulerB200Error(6, L.eu_last_error(h).decode())
    world = dist.get_world_size()
    blobs = [None]*world
    dist.all_gather_object(blobs, bytes(buf.raw))
    keep = [C.create_string_buffer(b, len(b)) for b in blobs]
    ptrs = (C.c_void_p*world)(*[C.cast(k, C.c_void_p) for k in keep])
    sizes = (C.c_int*world)(*[len(b) for b in blobs])
    if L.eu_comm_connect(h, world, ptrs, sizes) != EU_OK:
        raise EulerB200Error(6, L.eu_last_error(h).decode())
    use_cuda = dist.get_backend() == "nccl"

    def allreduce(user, values, n, op):
        arr = np.ctypeslib.as_array(values, shape=(n,))
        t = torch.tensor(arr.copy(), dtype=torch.float64, device="cuda" if use_cuda else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == 0 else dist.ReduceOp.MAX)
        arr[:] = t.cpu().numpy()

    cb = _ALLREDUCE(allreduce)
    dev._comm_keep = (cb, keep)
    if L.eu_comm_set_allreduce(h, cb, None) != EU_OK:
        raise EulerB200Error(6, L.eu_last_error(h).decode())
    dist.barrier()


def slab_ranges(n_cells, plane, world):
    """Contiguous whole-plane slabs: rank r owns cells [b[r], b[r+1])."""
    nz = n_cells//plane
    kb = [(nz*r)//world for r in range(world + 1)]
    return [k*plane for k in kb]
