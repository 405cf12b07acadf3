#!/usr/bin/env python
"""Time one of the BASELINE parity configurations on a B200 (not the bench line: bench.py is).

    python tools/perf_case.py c2 --n 160          # n^3 Cartesian, rotated anisotropic K, rock table, V+G+C
    python tools/perf_case.py c2b --n 160         # the same with diagonal tensor mobility (aniso_simulator_test's class)
    python tools/perf_case.py c3 --dims 256 256 128   # faulted corner-point, lognormal K, 3 rocks, V+G+C
    python tools/perf_case.py t3 --n 100          # tensor mobility on randomised oblique normals (three-component kernel)

Prints cell-substeps/s of the resident transportSolve with a fixed number of substeps, the algorithmic-byte
roofline fraction (a*N + 8*N_hf + b*N_f) and FAST-vs-STRICT agreement after the run."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "opm-porsol_b200", "python"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=["c2", "c2b", "c3", "t3"])
    ap.add_argument("--n", type=int, default=160)
    ap.add_argument("--dims", type=int, nargs=3, default=[256, 256, 128])
    ap.add_argument("--substeps", type=int, default=50)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cfl-fraction", type=float, default=0.25)
    ap.add_argument("--check-strict", action="store_true")
    a = ap.parse_args()
    import opm_porsol_b200 as eub
    from opm_porsol_b200 import synth
    from opm_porsol_b200.binding import make_fluid, params_from_case
    t0 = time.time()
    if a.case == "t3":
        # tensor mobility on oblique normals (three-component FAST kernel): randomised geometry, 2 rocks
        twin = synth.config_c2(8)
        fluid, _ = make_fluid(twin)
        fac = np.array(fluid.cfl_factor[:])
        case = synth.random_geometry_case(a.n, a.n, a.n, seed=5, n_rocks=2, mobility_kind=1, sources=False)
        case.clamp_sat = True          # randomised (unphysical) geometry: keep the run going, this case only times the kernel
    elif a.case == "c2b":
        # tensor mobility: the CFL factors of the scalar twin stand in for the reference's (the step count is fixed here)
        twin = synth.config_c2(8)
        fluid, _ = make_fluid(twin)
        fac = np.array(fluid.cfl_factor[:])
        case = synth.config_c2b(a.n)
    else:
        case = synth.config_c2(a.n) if a.case == "c2" else synth.config_c3(*a.dims)
        fluid, _ = make_fluid(case)
        fac = np.array(fluid.cfl_factor[:])
    case.min_steps = case.max_steps = a.substeps
    dev = eub.EulerUpstream(device=0, mode="fast")
    dev.init(params_from_case(case))
    dev.initObj(case, cfl_factors=fac)
    dev.upload_state(case.sat0, case.hf_flux)
    cfl = dev.cfl_times(case.gravity)
    t_step = a.cfl_fraction*min(cfl)*case.courant*a.substeps
    setup = time.time() - t0
    dev.transportSolveResident(t_step, case.gravity)
    ms, n = 0.0, 0
    for _ in range(a.steps):
        rep = dev.transportSolveResident(t_step, case.gravity)
        assert rep.attempts == 1 and rep.nsteps == a.substeps, (rep.attempts, rep.nsteps)
        ms += rep.device_ms
        n += rep.nsteps
    N, H = case.N, case.H
    nbr = case.hf_nbr
    cell_of = np.repeat(np.arange(N), np.diff(case.hf_offset))
    n_faces = int(((nbr < 0) | (nbr > cell_of)).sum())
    abytes = 40*N + 8*H + 24*n_faces
    peak = 6456.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    kernel_ms = ms/n
    out = {"case": case.name, "cells": N, "half_faces": H, "faces": n_faces, "kernel_ms": kernel_ms,
           "cell_substeps_per_s": N/(kernel_ms*1e-3), "bytes_per_cell_substep_model": abytes/N,
           "achieved_GBs": abytes/(kernel_ms*1e-3)/1e9, "frac_of_measured_peak": abytes/(kernel_ms*1e-3)/1e9/peak,
           "regular_slot_fraction": dev.regular_fraction(), "cfl_times": list(cfl), "setup_s": round(setup, 1),
           "mode": dev.resolved_mode()}
    sat = dev.download_saturation()
    out["sat_range"] = [float(sat.min()), float(sat.max())]
    if a.check_strict:
        st = eub.EulerUpstream(device=0, mode="strict")
        st.init(params_from_case(case))
        st.initObj(case, cfl_factors=fac)
        s1, s2 = case.sat0.copy(), case.sat0.copy()
        r1 = st.transportSolve(s1, t_step/5, case.gravity, case.hf_flux)
        dev.upload_state(case.sat0, case.hf_flux)
        r2 = dev.transportSolve(s2, t_step/5, case.gravity, case.hf_flux)
        out["fast_vs_strict_max_abs"] = float(np.abs(s1 - s2).max())
        out["strict_kernel_ms"] = r1.device_ms/r1.nsteps
        assert r1.nsteps == r2.nsteps
    print(json.dumps(out))


if __name__ == "__main__":
    main()
