#!/usr/bin/env python
"""bench.py -- EulerUpstream cell-substeps/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU code, bounded sample

One "step" is one transportSolve of the workload with a fixed number of substeps
(minimum_small_steps = maximum_small_steps = --substeps) so every step is the same work.
Default workload: BASELINE config "strong scaling: 512x512x256 Cartesian (67M cells) heterogeneous
perm", viscous + gravity, split into z-slabs over the ranks (strong scaling).

  value      cell-substeps/s with saturation and fluxes already resident in HBM (eu_transport_solve_resident)
  e2e        the same through eu_transport_solve with pinned HOST buffers (H2D of S and half-face fluxes and
             D2H of S inside the timed region)
  roofline   algorithmic bytes per substep (a*N + 8*N_hf + b*N_f) / CUDA-event time of the
             substep kernel, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (compiled reference if oracle/_ref exists, else the C port) on a bounded
             sample of the same workload, 1 thread (the reference is serial)
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "opm-porsol_b200", "python"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--ny", type=int, default=512)
    ap.add_argument("--nz", type=int, default=256)
    ap.add_argument("--substeps", type=int, default=100)
    ap.add_argument("--capillary", action="store_true", help="add the capillary term (160 B/cell-substep model)")
    ap.add_argument("--mode", default="auto", choices=["auto", "fast", "strict"])
    ap.add_argument("--cpu-cells", type=int, default=128*128*64, help="cells of the bounded CPU sample")
    ap.add_argument("--cpu-substeps", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs in the line (C2, C3, C4+capillary)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (pynvml, 100 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def workload_name(a):
    terms = "viscous+gravity+capillary" if a.capillary else "viscous+gravity"
    return f"C4 strong-scaling: {a.nx}x{a.ny}x{a.nz} Cartesian, lognormal perm, 1 rock table, {terms}, Dirichlet BCs"


def cfl_factors_for(a, fluid_case, synth, eub):
    """min trace(K)/3 and max porosity over the whole grid (cheap pre-pass over the plane seeds), then
    the reference's computeCflFactors on a one-cell stand-in with those extremes."""
    npl = a.nx*a.ny
    min_kx, max_poro = np.inf, 0.0
    for k in range(a.nz):
        u1 = synth.plane_uniform(44, k, npl)
        u2 = synth.plane_uniform(44 + 7, k, npl)
        z = np.sqrt(-2.0*np.log(1.0 - u1))*np.cos(2.0*np.pi*u2)
        kx = np.exp(np.log(100.0*synth.MILLIDARCY) + z)
        min_kx = min(min_kx, float(((kx + kx) + 0.1*kx).min()/3.0))
        max_poro = max(max_poro, float((0.05 + 0.25*synth.plane_uniform(45, k, npl)).max()))
    from opm_porsol_b200.binding import make_fluid
    one = synth.c4_fluid_case(a.capillary)
    one.perm[0, :] = 0.0
    one.perm[0, [0, 4, 8]] = min_kx
    one.poro[0] = max_poro
    fluid, _ = make_fluid(one)
    return np.array(fluid.cfl_factor[:])


def cpu_sample(a, synth, threads_note=True):
    """Bounded CPU run of the same workload family: the leading z-slab of the grid, smaller in x/y."""
    n = max(8, int(round((a.cpu_cells/4.0)**(1.0/3.0))))
    nx = ny = min(a.nx, 2*n)
    nz = max(2, min(a.nz, a.cpu_cells//(nx*ny)))
    d = synth.c4_slab(nx, ny, nz, 0, nz)
    g = dict(N=d["n_cells"], hf_offset=np.arange(d["n_cells"] + 1, dtype=np.int32)*6, hf_nbr=d["hf_neighbour"],
             hf_bid=np.where(d["hf_neighbour"] < 0, 1, 0).astype(np.int32), hf_area=d["hf_area"], hf_normal=d["hf_normal"],
             hf_centroid=d["hf_centroid"], cell_volume=d["cell_volume"], cell_centroid=d["cell_centroid"],
             bid_kind=np.zeros(2, dtype=np.int32), bid_sat=np.ones(2), bid_partner=np.zeros(2, dtype=np.int32), dims=(nx, ny, nz))
    case = synth.make_case("C4-sample", g, poro=d["porosity"], perm=d["permeability"], rock_id=d["rock_id"],
                           rocks=[synth.corey_table()], sat0=d["sat0"], gravity=[0.0, 0.0, -9.80665], hf_flux=d["hf_flux"],
                           method_capillary=a.capillary)
    case.min_steps = case.max_steps = a.cpu_substeps
    return case, (nx, ny, nz)


def run_cpu(a, synth):
    from oracle import ref as oracle
    case, dims = cpu_sample(a, synth)
    if oracle.ref_available():
        try:
            solver = oracle.RefSolver(case)
            kind = "reference"
            cfl, total = solver.cfl_times()
        except OSError:
            solver = None
    else:
        solver = None
    if solver is None:
        solver = oracle.PortSolver(case)
        kind = "port"
        cfl = solver.cfl_times()
    active = min(cfl[0], cfl[1], cfl[2] if a.capillary else 1e99)*case.courant
    out = solver.transport_solve(case.sat0, time=0.5*active*a.cpu_substeps)
    secs = out["seconds"]
    value = case.N*a.cpu_substeps/secs
    sample = (f"{dims[0]}x{dims[1]}x{dims[2]} slab of the same workload ({case.N} cells) x {a.cpu_substeps} substeps, "
              f"{secs:.2f} s in the substep loop, 1 thread (the reference is serial)")
    return {"value": value, "unit": "cell-substeps/s", "cores": 1, "kind": kind, "sample": sample, "seconds": secs}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from opm_porsol_b200 import synth

    if a.impl == "reference":
        if rank != 0:
            return 0
        t0 = time.time()
        vals = []
        for _ in range(max(1, min(a.steps, 3))):
            vals.append(run_cpu(a, synth))
            if time.time() - t0 > 120:
                break
        best = max(vals, key=lambda v: v["value"])
        ms = 1e3*best.pop("seconds")
        for v in vals:
            v.pop("seconds", None)
        line = {"impl": "reference", "metric": "EulerUpstream cell-substeps/s", "value": best["value"], "unit": "cell-substeps/s",
                "n_gpus": a.gpus, "steps": len(vals), "warmup": 0, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "substeps_per_step": a.cpu_substeps, "sample": best["sample"]},
                "cpu_baseline": best,
                "e2e": {"value": best["value"], "unit": "cell-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import opm_porsol_b200 as eub
    from opm_porsol_b200.binding import params_from_case

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    fluid_case = synth.c4_fluid_case(a.capillary)
    fluid_case.min_steps = fluid_case.max_steps = a.substeps
    if os.environ.get("EU_BENCH_NOCHECK"):        # kernel timing experiments only
        fluid_case.check_sat = False
    factors = cfl_factors_for(a, fluid_case, synth, eub)

    N = a.nx*a.ny*a.nz
    npl = a.nx*a.ny
    # z-slab decomposition: contiguous plane ranges per rank, one ghost plane on each inner side
    kb = [(a.nz*r)//world for r in range(world + 1)]
    k0, k1 = kb[rank], kb[rank + 1]
    g0, g1 = max(0, k0 - 1), min(a.nz, k1 + 1)
    n_local = (g1 - g0)*npl
    dev = eub.EulerUpstream(device=local_rank, mode=a.mode, rank=rank, world_size=world, own_begin=k0*npl, own_end=k1*npl)
    dev.init(params_from_case(fluid_case))

    sat_host = torch.empty(n_local, dtype=torch.float64, pin_memory=True)
    flux_host = torch.empty(n_local*6, dtype=torch.float64, pin_memory=True)
    sat_np, flux_np = sat_host.numpy(), flux_host.numpy()

    def chunks():
        step = max(1, (1 << 21)//npl)
        for ka in range(g0, g1, step):
            kz = min(g1, ka + step)
            d = synth.c4_slab(a.nx, a.ny, a.nz, ka, kz)
            o = (ka - g0)*npl
            sat_np[o:o + d["n_cells"]] = d["sat0"]
            flux_np[6*o:6*(o + d["n_cells"])] = d["hf_flux"]
            yield d

    t_setup = time.time()
    dev.initObjChunks(fluid_case, N, n_local, n_local*6, chunks(), factors)
    if world > 1:
        from opm_porsol_b200.comm import connect_ranks
        connect_ranks(dev, dist)
    dev.upload_state(sat_np, flux_np)
    gravity = fluid_case.gravity
    cfl = dev.cfl_times(gravity)
    active = min(cfl[0], cfl[1], cfl[2] if a.capillary else 1e99)*fluid_case.courant
    t_step = 0.5*active*a.substeps            # half the CFL step: stable, saturations stay in range
    t_setup = time.time() - t_setup
    own_cells = (k1 - k0)*npl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident: W warm-up + K timed steps
    for _ in range(a.warmup):
        dev.transportSolveResident(t_step, gravity)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, substeps = 0.0, 0, 0
    for _ in range(a.steps):
        rep = dev.transportSolveResident(t_step, gravity)
        dev_ms += rep.device_ms
        launches += rep.kernel_launches
        substeps += rep.substeps_executed
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    assert rep.attempts == 1 and rep.nsteps == a.substeps, (rep.attempts, rep.nsteps)

    # ---- end to end: host buffers through eu_transport_solve
    e2e_wall = None
    if not a.no_e2e:
        dev.transportSolve(sat_np, t_step, gravity, flux_np)            # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            dev.transportSolve(sat_np, t_step, gravity, flux_np)
        barrier()
        e2e_wall = time.perf_counter() - t0

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wall = reduce_max(wall)
    dev_ms = reduce_max(dev_ms)
    if e2e_wall is not None:
        e2e_wall = reduce_max(e2e_wall)
    s_min, s_max = float(sat_np.min()), float(sat_np.max())

    if rank == 0:
        total_substeps = a.steps*a.substeps
        value = N*total_substeps/wall
        n_hf = 6*N
        n_f = (a.nx + 1)*a.ny*a.nz + a.nx*(a.ny + 1)*a.nz + a.nx*a.ny*(a.nz + 1)
        aa, bb = (40, 24) if a.capillary else (32, 16)
        bytes_per_substep_total = aa*N + 8*n_hf + bb*n_f
        bytes_per_launch = bytes_per_substep_total/world            # per GPU and launch
        kernel_ms = dev_ms/total_substeps
        achieved = bytes_per_launch/(kernel_ms*1e-3)/1e9
        peak, peak_src = measured_peak()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        key = "viscous+gravity+capillary" if a.capillary else "viscous+gravity"
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if key in tj:       # dram__bytes_read + dram__bytes_write of k_fast_step per cell, from the committed ncu capture
                traffic = tj[key]["dram_bytes_per_cell_substep"]*N/world
        line = {
            "metric": "EulerUpstream cell-substeps/s", "value": value, "unit": "cell-substeps/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3*wall/a.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "cells": N, "substeps_per_step": a.substeps,
                       "parallelism": f"z-slabs x{world}", "arithmetic_mode": a.mode,
                       "l2": "inputs (>= 8 GB per substep at full size) exceed the 126 MB L2; no flush needed",
                       "setup_s": round(t_setup, 1), "sat_range_after": [s_min, s_max],
                       "regular_slot_fraction": round(dev.regular_fraction(), 4)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "k_fast_step" if a.mode != "strict" else "k_strict_step",
                         "bytes_per_cell_substep": bytes_per_substep_total/N, "kernel_ms": kernel_ms},
            "clocks": clocks, "gpu_launches": launches,
        }
        if e2e_wall is not None:
            line["e2e"] = {"value": N*total_substeps/e2e_wall, "unit": "cell-substeps/s",
                           "h2d_bytes_per_step": 8*n_local + 8*6*n_local, "d2h_bytes_per_step": 8*n_local,
                           "ms_per_step": 1e3*e2e_wall/a.steps}
        if not a.no_cpu:
            line["cpu_baseline"] = run_cpu(a, synth)
            line["cpu_baseline"].pop("seconds", None)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    dev.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
