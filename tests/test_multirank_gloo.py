"""World-size-2 check of the decomposition logic on CPU (gloo): slab extraction, ghost lists, the send plan of
eu_comm_plan_sends and the principle that faces on a slab boundary are evaluated redundantly on both sides
from identical operands -- so the decomposed result is BIT-IDENTICAL to the single-domain result.
The arithmetic is done by the CPU oracle here (it is the checker); the GPU version of the same test is
tests/multigpu_check.py (run under torchrun on 2+ B200)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, active_cfl_dt


def _worker(rank, world, port, case_name, q):
    sys.path.insert(0, os.path.join(ROOT, "opm-porsol_b200", "python"))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import opm_porsol_b200 as eub
    from opm_porsol_b200 import synth
    from oracle.ref import PortSolver
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = _make_case(case_name)
    glob = PortSolver(case)
    fac = glob.compute_cfl_factors()
    dt = 0.5*active_cfl_dt(case, glob.cfl_times())
    nx, ny, nz = case.dims
    bounds = [nx*ny*((nz*r)//world) for r in range(world + 1)]
    slab = synth.extract_slab(case, bounds[rank], bounds[rank + 1])
    loc = synth.local_case(case, slab)
    port_loc = PortSolver(loc, cfl_factors=fac)
    cells = slab["cells"]
    own = (cells >= bounds[rank]) & (cells < bounds[rank + 1])
    ghosts_g = cells[~own].astype(np.int32)
    ghosts_l = np.nonzero(~own)[0].astype(np.int32)
    gathered = [None]*world
    dist.all_gather_object(gathered, (ghosts_g, ghosts_l))
    lib = eub.load_library()
    ip = C.POINTER(C.c_int)
    lib.eu_comm_plan_sends.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, ip]
    g2l = {int(g): i for i, g in enumerate(cells)}
    plans = {}
    for p in range(world):
        if p == rank:
            continue
        pg, pl = gathered[p]
        sg = np.zeros(max(1, pg.size), dtype=np.int32)
        sl = np.zeros(max(1, pg.size), dtype=np.int32)
        n = lib.eu_comm_plan_sends(bounds[rank], bounds[rank + 1], pg.size, pg.ctypes.data_as(ip), pl.ctypes.data_as(ip),
                                   sg.ctypes.data_as(ip), sl.ctypes.data_as(ip))
        plans[p] = (np.array([g2l[int(g)] for g in sg[:n]], dtype=np.int64), sl[:n].astype(np.int64))
    s = loc.sat0.copy()
    s_glob = case.sat0.copy()
    ok = True
    for step in range(5):
        s_glob = glob.small_step(s_glob, dt)["sat"]
        new = port_loc.small_step(s, dt)["sat"]
        s[own] = new[own]
        # halo exchange: what the device does with peer-to-peer stores
        out = [None]*world
        dist.all_gather_object(out, {p: (dst, s[src]) for p, (src, dst) in plans.items()})
        for p in range(world):
            if p != rank and rank in out[p]:
                dst, vals = out[p][rank]
                s[dst] = vals
        ok = ok and np.array_equal(s, s_glob[cells])          # own AND ghost values, bit for bit
    q.put((rank, bool(ok), int(ghosts_g.size)))
    dist.destroy_process_group()


def _make_case(name):
    from opm_porsol_b200 import synth
    if name == "cart_zperiodic":
        c = synth.random_geometry_case(5, 4, 6, seed=31, n_rocks=2, periodic=(False, False, True))
    elif name == "faulted":
        c = synth.config_c3(12, 8, 8)
    else:
        c = synth.config_c2(8)
    return c


@pytest.mark.parametrize("case_name", ["cart", "cart_zperiodic", "faulted"])
def test_two_rank_decomposition_is_bit_identical(case_name):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + {"cart": 0, "cart_zperiodic": 1, "faulted": 2}[case_name]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(ng > 0 for _, _, ng in res)
