"""Run under torchrun on N >= 2 GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py

Each rank owns a z-slab of a small seeded case (ghost cells uploaded alongside), ghost saturations travel as
peer-to-peer stores.  Checks, per case and arithmetic mode: step counts equal the CPU oracle's, STRICT results
are BIT-IDENTICAL to the oracle (hence independent of the GPU count), FAST results within 1e-9."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "opm-porsol_b200", "python"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import opm_porsol_b200 as eub
    from opm_porsol_b200 import synth
    from opm_porsol_b200.binding import params_from_case
    from opm_porsol_b200.comm import connect_ranks
    from conftest import active_cfl_dt
    from oracle.ref import PortSolver

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("c2_cap", synth.config_c2(12)),
             ("zperiodic_2rocks", synth.random_geometry_case(6, 5, 2*world + 4, seed=41, n_rocks=2, periodic=(True, False, True))),
             ("faulted_3rocks", synth.config_c3(16, 12, 4*world)),
             ("retry", synth.random_geometry_case(5, 4, 2*world + 2, seed=21, n_rocks=1, sources=False)),
             # march-active plane shapes (nx*ny % 32 == 0): slab interiors run the z-march the bench times, the planes next
             # to a slab boundary the fused halo push
             ("c4_march_vg", synth.config_c4(64, 16, 6*world)),
             ("c4_march_vgc", synth.config_c4(32, 32, 6*world, capillary=True))]
    cases[3][1].max_steps = 2
    bad = 0
    for name, case in cases:
        port = PortSolver(case)
        fac = port.compute_cfl_factors()
        total = active_cfl_dt(case, port.cfl_times())
        time = (40.0 if name == "retry" else 17.3)*total
        want = port.transport_solve(case.sat0, time=time)
        nx, ny, nz = case.dims
        bounds = [nx*ny*((nz*r)//world) for r in range(world + 1)]
        slab = synth.extract_slab(case, bounds[rank], bounds[rank + 1])
        cells = slab["cells"]
        own = (cells >= bounds[rank]) & (cells < bounds[rank + 1])
        for mode, tol in (("strict", 0.0), ("fast", 1e-9)):
            dev = eub.EulerUpstream(device=local, mode=mode, rank=rank, world_size=world, own_begin=bounds[rank], own_end=bounds[rank + 1])
            dev.init(params_from_case(case))
            dev.initObjChunks(case, case.N, slab["n_local"], slab["n_hf"], slab["chunks"], fac)
            connect_ranks(dev, dist)
            sat = np.ascontiguousarray(case.sat0[cells])
            rep = dev.transportSolve(sat, time, case.gravity, case.hf_flux[slab["hf_index"]], (case.src_cell, case.src_rate),
                                     raise_on_error=False)
            err_own = float(np.abs(sat[own] - want["sat"][cells[own]]).max())
            err_ghost = float(np.abs(sat[~own] - want["sat"][cells[~own]]).max()) if (~own).any() else 0.0
            ok = (rep.nsteps == want["nsteps"] and rep.attempts == want["attempts"] and (rep.status != 0) == (want["status"] != 0)
                  and (want["status"] != 0 or (err_own <= tol and err_ghost <= tol)))
            t = torch.tensor([0 if ok else 1], device="cuda")
            dist.all_reduce(t)
            if rank == 0:
                print(f"{name:18s} {mode:6s} world={world} steps={rep.nsteps} attempts={rep.attempts} "
                      f"max|dS| own={err_own:.2e} ghost={err_ghost:.2e} {'OK' if int(t.item()) == 0 else 'FAILED'}", flush=True)
            bad += int(t.item())
            dev.close()
    if rank == 0:
        print("MULTIGPU CHECK PASSED" if bad == 0 else "MULTIGPU CHECK FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
