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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: C4 512x512x256 split over the ranks (default); weak: C5, 256x256x122 faulted corner-point cells per rank")
    ap.add_argument("--weak-planes", type=int, default=122, help="layers per rank of the weak-scaling workload (256x256x122 = 8.0 M cells)")
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


def workload_name(a, world=1):
    if a.scaling == "weak":
        return (f"C5 weak-scaling: 256x256x{a.weak_planes} faulted corner-point cells per GPU ({256*256*a.weak_planes*world} cells on "
                f"{world}), lognormal perm, 3 rock types, viscous+gravity+capillary, Dirichlet BCs")
    terms = "viscous+gravity+capillary" if a.capillary else "viscous+gravity"
    return f"C4 strong-scaling: {a.nx}x{a.ny}x{a.nz} Cartesian, lognormal perm, 1 rock table, {terms}, Dirichlet BCs"


def host_cpu():
    """CPU model and logical core count of the box (BASELINE.md section 3 asks for both next to the CPU number)."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"model": model, "logical_cores": os.cpu_count()}


def algorithmic_bytes(n_cells, n_hf, n_faces, capillary):
    """a*N + 8*N_hf + b*N_f with (a, b) = (32, 16) V+G, (40, 24) V+G+C."""
    aa, bb = (40, 24) if capillary else (32, 16)
    return aa*n_cells + 8*n_hf + bb*n_faces


def time_inmemory_config(name, case, fac, substeps, steps, cfl_fraction, peak):
    """One of the single-GPU BASELINE configurations that fits in host memory as a Case: resident transportSolve with a
    fixed number of substeps; kernel ms per substep from the library's CUDA events."""
    import opm_porsol_b200 as eub
    from opm_porsol_b200.binding import params_from_case
    case.min_steps = case.max_steps = substeps
    dev = eub.EulerUpstream(device=0, mode="fast")
    dev.init(params_from_case(case))
    dev.initObj(case, cfl_factors=fac)
    dev.upload_state(case.sat0, case.hf_flux)
    cfl = dev.cfl_times(case.gravity)
    t_step = cfl_fraction*min(cfl)*case.courant*substeps
    dev.transportSolveResident(t_step, case.gravity)
    ms, n = 0.0, 0
    for _ in range(steps):
        rep = dev.transportSolveResident(t_step, case.gravity)
        assert rep.attempts == 1 and rep.nsteps == substeps, (rep.attempts, rep.nsteps)
        ms += rep.device_ms
        n += rep.nsteps
    N, H = case.N, case.H
    cell_of = np.repeat(np.arange(N), np.diff(case.hf_offset))
    n_faces = int(((case.hf_nbr < 0) | (case.hf_nbr > cell_of)).sum())
    abytes = algorithmic_bytes(N, H, n_faces, case.method_capillary)
    kernel_ms = ms/n
    sat = dev.download_saturation()
    plan = dev.work_plan()
    dev.close()
    return {"workload": name, "cells": N, "kernel_ms": kernel_ms, "value": N/(kernel_ms*1e-3), "unit": "cell-substeps/s",
            "bytes_per_cell_substep": abytes/N, "achieved_GBs": abytes/(kernel_ms*1e-3)/1e9,
            "frac": abytes/(kernel_ms*1e-3)/1e9/peak, "substeps": n, "sat_range_after": [float(sat.min()), float(sat.max())],
            "work_units": plan["items"]}


def other_configs(a, synth, peak):
    """kernel ms / cell-substeps/s / algorithmic-byte fraction of the other single-GPU BASELINE configurations, each about a
    second of GPU time (resident state, fixed substeps): C2 100^3, C3 256x256x128, C4 with the capillary term (a 64-plane slab)."""
    from opm_porsol_b200.binding import make_fluid
    out = {}
    def factors(case):
        fluid, _ = make_fluid(case)
        return np.array(fluid.cfl_factor[:])
    t0 = time.time()
    c2 = synth.config_c2(100)
    out["C2"] = time_inmemory_config("C2: 100x100x100 Cartesian, rotated anisotropic K, 1 rock table, viscous+gravity+capillary",
                                     c2, factors(c2), 50, 3, 0.25, peak)
    del c2
    c3 = synth.config_c3(256, 256, 128)
    out["C3"] = time_inmemory_config("C3: 256x256x128 faulted corner-point, lognormal K, 3 rock types, viscous+gravity+capillary",
                                     c3, factors(c3), 30, 3, 0.25, peak)
    del c3
    c4c = synth.config_c4(512, 512, 32, capillary=True)
    out["C4+capillary"] = time_inmemory_config("C4 with the capillary term: 512x512x32 slab of the strong-scaling grid, viscous+gravity+capillary",
                                               c4c, factors(c4c), 30, 3, 0.5, peak)
    del c4c
    out["seconds"] = round(time.time() - t0, 1)
    return out





def cfl_factors_for(a, fluid_case, synth, eub, seed=44):
    """min trace(K)/3 and max porosity over the whole grid (cheap pre-pass over the plane seeds), then
    the reference's computeCflFactors on a one-cell stand-in with those extremes."""
    npl = a.nx*a.ny
    min_kx, max_poro = np.inf, 0.0
    for k in range(a.nz):
        u1 = synth.plane_uniform(seed, k, npl)
        u2 = synth.plane_uniform(seed + 7, k, npl)
        z = np.sqrt(-2.0*np.log(1.0 - u1))*np.cos(2.0*np.pi*u2)
        kx = np.exp(np.log(100.0*synth.MILLIDARCY) + z)
        min_kx = min(min_kx, float(((kx + kx) + 0.1*kx).min()/3.0))
        max_poro = max(max_poro, float((0.05 + 0.25*synth.plane_uniform(seed + 1, k, npl)).max()))
    from opm_porsol_b200.binding import make_fluid
    one = synth.c5_fluid_case() if a.scaling == "weak" else synth.c4_fluid_case(a.capillary)
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
    if a.scaling == "weak":
        # the same generator as the GPU workload (faulted corner-point, 3 rocks, V+G+C) at the sample's size
        nx = ny = min(256, nx)
        d = synth.c5_slab(nx, ny, nz, 0, nz)
        off = np.zeros(d["n_cells"] + 1, dtype=np.int32)
        np.cumsum(d["hf_count"], out=off[1:])
        g = dict(N=d["n_cells"], hf_offset=off, hf_nbr=d["hf_neighbour"], hf_bid=np.where(d["hf_neighbour"] < 0, 1, 0).astype(np.int32),
                 hf_area=d["hf_area"], hf_normal=d["hf_normal"], hf_centroid=d["hf_centroid"], cell_volume=d["cell_volume"],
                 cell_centroid=d["cell_centroid"], bid_kind=np.zeros(2, dtype=np.int32), bid_sat=np.ones(2),
                 bid_partner=np.zeros(2, dtype=np.int32), dims=(nx, ny, nz))
        case = synth.make_case("C5-sample", g, poro=d["porosity"], perm=d["permeability"], rock_id=d["rock_id"], rocks=synth.c5_rocks(),
                               sat0=d["sat0"], gravity=[0.0, 0.0, -9.80665], hf_flux=d["hf_flux"])
        case.min_steps = case.max_steps = a.cpu_substeps
        return case, (nx, ny, nz)
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
    active = min(cfl[0], cfl[1], cfl[2] if (a.capillary or a.scaling == "weak") else 1e99)*case.courant
    out = solver.transport_solve(case.sat0, time=(0.25 if a.scaling == "weak" else 0.5)*active*a.cpu_substeps)
    secs = out["seconds"]
    value = case.N*a.cpu_substeps/secs
    sample = (f"{dims[0]}x{dims[1]}x{dims[2]} slab of the same workload ({case.N} cells) x {a.cpu_substeps} substeps, "
              f"{secs:.2f} s in the substep loop, 1 thread (the reference is serial)")
    return {"value": value, "unit": "cell-substeps/s", "cores": 1, "kind": kind, "sample": sample, "seconds": secs,
            "host_cpu": host_cpu()}


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
                "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a, a.gpus), "substeps_per_step": a.cpu_substeps, "sample": best["sample"]},
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
    weak = a.scaling == "weak"
    if weak:
        a.nx, a.ny, a.nz, a.capillary = 256, 256, a.weak_planes*world, True
        fluid_case = synth.c5_fluid_case()
    else:
        fluid_case = synth.c4_fluid_case(a.capillary)
    fluid_case.min_steps = fluid_case.max_steps = a.substeps
    if os.environ.get("EU_BENCH_NOCHECK"):        # kernel timing experiments only
        fluid_case.check_sat = False
    # min trace(K)/3 and max porosity over the grid come from the plane seeds; both workloads share the property law
    # (lognormal kx around 100 mD, kz = 0.1 kx, phi in [0.05, 0.30]) up to the seed
    seed = 42 if weak else 44
    factors = cfl_factors_for(a, fluid_case, synth, eub, seed)

    N = a.nx*a.ny*a.nz
    npl = a.nx*a.ny
    # z-slab decomposition: contiguous plane ranges per rank, with the ghost planes the faces of own cells reach
    # (1 for the Cartesian grid, the largest fault throw for the corner-point grid)
    depth = synth.C5_GHOST_DEPTH if weak else 1
    kb = [(a.nz*r)//world for r in range(world + 1)]
    k0, k1 = kb[rank], kb[rank + 1]
    g0, g1 = max(0, k0 - depth), min(a.nz, k1 + depth)
    n_local = (g1 - g0)*npl
    dev = eub.EulerUpstream(device=local_rank, mode=a.mode, rank=rank, world_size=world, own_begin=k0*npl, own_end=k1*npl)
    dev.init(params_from_case(fluid_case))

    t_setup = time.time()
    step = max(1, (1 << 21)//npl)
    if weak:
        # cells of the faulted grid have 6 to 8 faces: generate the rank's chunks first, then size the buffers
        parts = [synth.c5_slab(a.nx, a.ny, a.nz, ka, min(g1, ka + step)) for ka in range(g0, g1, step)]
        n_local_hf = int(sum(int(d["hf_count"].sum()) for d in parts))
    else:
        parts = None
        n_local_hf = n_local*6
    sat_host = torch.empty(n_local, dtype=torch.float64, pin_memory=True)
    flux_host = torch.empty(n_local_hf, dtype=torch.float64, pin_memory=True)
    sat_np, flux_np = sat_host.numpy(), flux_host.numpy()
    own_hf = [0]          # half-faces / unique faces of the own cells, for the algorithmic byte count
    own_faces = [0]

    def chunks():
        o, oh = 0, 0
        gen = parts if weak else (synth.c4_slab(a.nx, a.ny, a.nz, ka, min(g1, ka + step)) for ka in range(g0, g1, step))
        for d in gen:
            n, nh = d["n_cells"], int(d["hf_count"].sum())
            sat_np[o:o + n] = d["sat0"]
            flux_np[oh:oh + nh] = d["hf_flux"]
            cell = np.repeat(np.arange(d["first_cell"], d["first_cell"] + n), d["hf_count"])
            own = (cell >= k0*npl) & (cell < k1*npl)
            own_hf[0] += int(own.sum())
            own_faces[0] += int((own & ((d["hf_neighbour"] < 0) | (d["hf_neighbour"] > cell))).sum())
            o += n
            oh += nh
            yield d

    dev.initObjChunks(fluid_case, N, n_local, n_local_hf, chunks(), factors)
    parts = None
    if world > 1:
        from opm_porsol_b200.comm import connect_ranks
        connect_ranks(dev, dist)
    dev.upload_state(sat_np, flux_np)
    gravity = fluid_case.gravity
    cfl = dev.cfl_times(gravity)
    active = min(cfl[0], cfl[1], cfl[2] if a.capillary else 1e99)*fluid_case.courant
    # well inside the CFL step.  The synthetic flux of the faulted grid (a constant velocity projected on split faces) is
    # not divergence-free, so some cells drain at a fixed rate: the weak-scaling steps are kept short in simulated time
    # (the work per substep does not depend on dt) so that 800 substeps stay inside the saturation range.
    t_step = (0.02 if weak else 0.5)*active*a.substeps
    t_setup = time.time() - t_setup

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
    plan = dev.work_plan()

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

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world > 1 else (None, None)
    wall = reduce(wall, MAX)
    dev_ms = reduce(dev_ms, MAX)
    if e2e_wall is not None:
        e2e_wall = reduce(e2e_wall, MAX)
    n_hf = int(reduce(float(own_hf[0]), SUM))
    n_f = int(reduce(float(own_faces[0]), SUM))
    h2d = int(reduce(float(8*n_local + 8*n_local_hf), SUM))
    d2h = int(reduce(float(8*n_local), SUM))
    s_min, s_max = float(sat_np.min()), float(sat_np.max())
    dev.close()

    if rank == 0:
        total_substeps = a.steps*a.substeps
        value = N*total_substeps/wall
        bytes_per_substep_total = algorithmic_bytes(N, n_hf, n_f, a.capillary)
        bytes_per_launch = bytes_per_substep_total/world            # per GPU and substep
        kernel_ms = dev_ms/total_substeps
        achieved = bytes_per_launch/(kernel_ms*1e-3)/1e9
        peak, peak_src = measured_peak()
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        key = "viscous+gravity+capillary" if a.capillary else "viscous+gravity"
        if os.path.exists(tpath) and not weak:
            with open(tpath) as f:
                tj = json.load(f)
            if key in tj:
                traffic = tj[key]["dram_bytes_per_cell_substep"]*N/world
                traffic_src = ("not measured in this run: dram__bytes_read+write per cell of the substep kernels from the committed "
                               "ncu capture " + tj[key].get("source", "profiles/r02_traffic.json") + ", scaled to this run's cells per GPU")
        line = {
            "metric": "EulerUpstream cell-substeps/s", "value": value, "unit": "cell-substeps/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3*wall/a.steps,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "cells": N, "substeps_per_step": a.substeps,
                       "parallelism": f"z-slabs x{world}", "arithmetic_mode": a.mode,
                       "l2": "inputs (>= 1 GB per substep and GPU) exceed the 126 MB L2; no flush needed",
                       "setup_s": round(t_setup, 1), "sat_range_after": [s_min, s_max],
                       "kernel_path": "box kernel (TMA-staged plane sweep)" if plan["kernel"] == "box" else "slice-class kernel",
                       "work_units": plan["items"], "planes_per_unit": plan["mean_march"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "k_box_step (+ k_box_irregular pre-pass)" if a.mode != "strict" else "k_strict_step",
                         "bytes_per_cell_substep": bytes_per_substep_total/N, "kernel_ms": kernel_ms,
                         "frac_of_dram_traffic": (traffic/(kernel_ms*1e-3)/1e9/peak) if traffic else None},
            "clocks": clocks, "gpu_launches": launches,
        }
        if e2e_wall is not None:
            line["e2e"] = {"value": N*total_substeps/e2e_wall, "unit": "cell-substeps/s",
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                           "ms_per_step": 1e3*e2e_wall/a.steps,
                           "note": "eu_transport_solve with pinned host buffers; the copies (PCIe) are inside the timed region"}
        if world == 1 and not a.no_configs and not weak:
            line["configs"] = other_configs(a, synth, peak)
        if not a.no_cpu:
            line["cpu_baseline"] = run_cpu(a, synth)
            line["cpu_baseline"].pop("seconds", None)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
