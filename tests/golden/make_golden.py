#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libeuler_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the seeded inputs' identity (case name + generator arguments are in
tests/conftest.py: small_cases / tensor_cases) and the reference's outputs: CFL factors, the three CFL
times, saturation and residual after each of 3 substeps, a full transportSolve (saturation,
step count, attempts), computeResidual with explicit method flags and the SimulatorUtilities.hpp diagnostics
(cell velocity, phase velocities, capillary pressures).  The GPU box has no /root/reference; tests compare against these files there.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "opm-porsol_b200", "python"))
sys.path.insert(0, ROOT)

from conftest import active_cfl_dt, small_cases, tensor_cases  # noqa: E402
from oracle.ref import RefSolver  # noqa: E402


def main():
    for name, case in small_cases() + tensor_cases():
        ref = RefSolver(case)
        fac = ref.cfl_factors()
        cfl, _ = ref.cfl_times()
        total = active_cfl_dt(case, cfl)
        dt = 0.5*total if case.mobility_kind == 0 else 50.0
        s = case.sat0.copy()
        step_sat, step_res = [], []
        for _ in range(3):
            o = ref.small_step(s, dt)
            assert o["status"] == 0
            s = o["sat"]
            step_sat.append(s.copy())
            step_res.append(o["residual"].copy())
        out = dict(cfl_factors=fac, cfl_times=cfl, dt=dt, step_sat=np.array(step_sat), step_res=np.array(step_res),
                   input_checksum=np.array([case.sat0.sum(), case.hf_flux.sum(), case.perm.sum(), case.hf_area.sum()]))
        # the residual as an operator (ImplicitCapillarity's call: capillary off) and the post-transport diagnostics
        out.update(res_vg=ref.compute_residual(case.sat0, (True, True, False)),
                   res_gc=ref.compute_residual(case.sat0, (False, True, True)),
                   cell_velocity=ref.cell_velocity(), cap_pressures=ref.cap_pressures(case.sat0))
        vw, vo = ref.phase_velocities(case.sat0, out["cell_velocity"])
        out.update(water_velocity=vw, oil_velocity=vo)
        if case.mobility_kind == 0:
            time = 17.3*total
            sol = ref.transport_solve(case.sat0, time=time)
            assert sol["status"] == 0
            out.update(solve_time=time, solve_sat=sol["sat"], solve_nsteps=sol["nsteps"], solve_attempts=sol["attempts"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, case.N, "cells", out.get("solve_nsteps"))


def periodic_matching():
    """tests/golden/periodic_matching.npz: findPeriodicPartners / match of the reference (BoundaryPeriodicity.hpp:86-177,
    .cpp:25-49) on the seeded boundary-face lists of tests/test_periodic_matching.py."""
    from oracle.ref import ref_find_periodic_partners
    from test_periodic_matching import CASES, FIXTURE, boundary_faces
    out = {}
    for k, (dims, per, seed) in enumerate(CASES):
        cen, area = boundary_faces(*dims, seed=seed)
        st, canon, partner, sides = ref_find_periodic_partners(cen, area, per)
        assert st == 0
        out[f"cen{k}"], out[f"canon{k}"], out[f"partner{k}"], out[f"sides{k}"] = cen, canon, partner, sides
        print("periodic", dims, per, int((partner >= 0).sum()), "matched of", area.shape[0])
    np.savez_compressed(FIXTURE, **out)


if __name__ == "__main__":
    main()
    periodic_matching()
