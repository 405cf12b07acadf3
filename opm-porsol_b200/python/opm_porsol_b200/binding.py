"""ctypes binding of include/euler_b200.h and the Python mirror of the reference interface.

Reference interface mirrored (opm/porsol/euler/EulerUpstream.hpp:51-102):
    EulerUpstream()                       -> EulerUpstream(device=0, mode=...)
    init(param)                           -> init(param_dict)       keys of EulerUpstream_impl.hpp:95-108
    initObj(grid, resprop, boundary)      -> initObj(case)          flat arrays of the same three objects
    transportSolve(sat, time, gravity,    -> transportSolve(sat, time, gravity, hf_flux, (src_cell, src_rate))
                   pressure_sol, inj)        sat is updated in place; raises like the reference throws
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

EU_OK, EU_ERR_ARG, EU_ERR_CUDA, EU_ERR_SAT_RANGE, EU_ERR_CFL_ZERO, EU_ERR_UNSUPPORTED, EU_ERR_COMM = range(7)
EU_HF_DIRICHLET, EU_HF_PERIODIC = 1, 2
EU_MODE_AUTO, EU_MODE_STRICT, EU_MODE_FAST = 0, 1, 2
MODES = {"auto": EU_MODE_AUTO, "strict": EU_MODE_STRICT, "fast": EU_MODE_FAST}


class EulerB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class _Config(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("device", C.c_int), ("mode", C.c_int), ("rank", C.c_int),
                ("world_size", C.c_int), ("own_begin", C.c_int), ("own_end", C.c_int)]


class _Params(C.Structure):
    _fields_ = [("courant_number", C.c_double), ("method_viscous", C.c_int), ("method_gravity", C.c_int),
                ("method_capillary", C.c_int), ("use_cfl_viscous", C.c_int), ("use_cfl_gravity", C.c_int),
                ("use_cfl_capillary", C.c_int), ("minimum_small_steps", C.c_int), ("maximum_small_steps", C.c_int),
                ("check_sat", C.c_int), ("clamp_sat", C.c_int)]


class _Chunk(C.Structure):
    _fields_ = [("first_cell", C.c_int), ("n_cells", C.c_int), ("hf_count", _ip), ("hf_neighbour", _ip),
                ("hf_area", _dp), ("hf_normal", _dp), ("hf_centroid", _dp),
                ("n_bnd", C.c_int), ("bnd_hf", _ip), ("bnd_kind", _ip), ("bnd_sat", _dp),
                ("bnd_partner_cell", _ip), ("bnd_partner_face", _ip),
                ("cell_volume", _dp), ("cell_centroid", _dp), ("porosity", _dp), ("permeability", _dp),
                ("rock_id", _ip)]


class _Fluid(C.Structure):
    _fields_ = [("mobility_kind", C.c_int), ("viscosity", C.c_double*2), ("density", C.c_double*2),
                ("cfl_factor", C.c_double*3), ("use_jfunction_scaling", C.c_int), ("sigma_cos_theta", C.c_double),
                ("n_rocks", C.c_int), ("table_offset", _ip), ("table_s", _dp), ("table_cols", _dp*7)]


class _Report(C.Structure):
    _fields_ = [("status", C.c_int), ("nsteps", C.c_int), ("attempts", C.c_int), ("substeps_executed", C.c_longlong),
                ("bad_cell", C.c_int), ("bad_value", C.c_double), ("cfl_dt", C.c_double*3), ("dt", C.c_double),
                ("device_ms", C.c_double), ("kernel_launches", C.c_int)]


@dataclass
class Report:
    status: int
    nsteps: int
    attempts: int
    substeps_executed: int
    bad_cell: int
    bad_value: float
    cfl_dt: tuple
    dt: float
    device_ms: float
    kernel_launches: int


def lib_path() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    override = os.environ.get("EU_B200_LIB")           # kernel experiments: another build of the same library
    if override:
        return override
    return os.path.normpath(os.path.join(here, "..", "..", "lib", "libeuler_b200.so"))


_LIB = None


def load_library():
    """Load libeuler_b200.so.  Fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise EulerB200Error(EU_ERR_CUDA, f"{p} is missing: build it with __graft_entry__.build() "
                                          "(make -C opm-porsol_b200/csrc); there is no CPU fallback")
    L = C.CDLL(p)
    L.eu_last_error.restype = C.c_char_p
    L.eu_last_error.argtypes = [C.c_void_p]
    L.eu_create.argtypes = [C.POINTER(_Config), C.POINTER(C.c_void_p)]
    L.eu_destroy.argtypes = [C.c_void_p]
    L.eu_set_params.argtypes = [C.c_void_p, C.POINTER(_Params)]
    L.eu_default_params.argtypes = [C.POINTER(_Params)]
    L.eu_grid_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong]
    L.eu_grid_append.argtypes = [C.c_void_p, C.POINTER(_Chunk)]
    L.eu_set_fluid.argtypes = [C.c_void_p, C.POINTER(_Fluid)]
    L.eu_grid_end.argtypes = [C.c_void_p]
    L.eu_local_cells.argtypes = [C.c_void_p]
    L.eu_resolved_mode.argtypes = [C.c_void_p]
    L.eu_local_halffaces.argtypes = [C.c_void_p]
    L.eu_local_halffaces.restype = C.c_longlong
    L.eu_regular_fraction.argtypes = [C.c_void_p]
    L.eu_regular_fraction.restype = C.c_double
    L.eu_work_plan.argtypes = [C.c_void_p, _dp]
    L.eu_transport_solve.argtypes = [C.c_void_p, _dp, C.c_double, _dp, _dp, C.c_int, _ip, _dp, C.POINTER(_Report)]
    L.eu_upload_state.argtypes = [C.c_void_p, _dp, _dp]
    L.eu_upload_saturation.argtypes = [C.c_void_p, _dp]
    L.eu_download_saturation.argtypes = [C.c_void_p, _dp]
    L.eu_transport_solve_resident.argtypes = [C.c_void_p, C.c_double, _dp, C.c_int, _ip, _dp, C.POINTER(_Report)]
    L.eu_cfl_times.argtypes = [C.c_void_p, _dp, _dp]
    L.eu_small_step.argtypes = [C.c_void_p, C.c_double, _dp, C.c_int, _ip, _dp, _dp, _ip, _dp]
    L.eu_compute_cfl_factors.argtypes = [C.POINTER(_Fluid), C.c_int, _dp, _dp, _ip, _dp]
    L.eu_match_periodic_faces.argtypes = [C.c_int, _dp, _dp, _ip, C.c_double, _ip, _ip, _dp]
    L.eu_compute_residual.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, _ip, _dp, C.c_int, C.c_int, C.c_int, _dp]
    L.eu_compute_cap_pressures.argtypes = [C.c_void_p, _dp, _dp]
    L.eu_cell_velocity.argtypes = [C.c_void_p, _dp]
    L.eu_phase_velocities.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    L.eu_fractional_flow.argtypes = [C.c_void_p, _dp, _dp]
    _LIB = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def resolve_boundary(case):
    """Boundary ids -> per-half-face boundary records, the way the reference resolves them:
    satCond(face) by boundary id (BoundaryConditions.hpp:420-425) and, for periodic faces,
    bid_to_face_[getPeriodicPartner(bid)] (EulerUpstreamResidual_impl.hpp:403-421,118-120).
    Returns (bnd_hf, kind, sat, partner_cell, partner_face) over all boundary half-faces."""
    hf_nbr = case.hf_nbr
    bnd_hf = np.nonzero(hf_nbr < 0)[0].astype(np.int32)
    bid = case.hf_bid[bnd_hf]
    kind = np.where(case.bid_kind[bid] == 1, EU_HF_PERIODIC, EU_HF_DIRICHLET).astype(np.int32)
    sat = np.ascontiguousarray(case.bid_sat[bid], dtype=np.float64)
    pcell = np.full(bnd_hf.shape[0], -1, dtype=np.int32)
    pface = np.full(bnd_hf.shape[0], -1, dtype=np.int32)
    per = kind == EU_HF_PERIODIC
    if per.any():
        n_bid = case.bid_kind.shape[0]
        bid_to_hf = np.full(n_bid, -1, dtype=np.int64)
        bid_to_hf[bid[per]] = bnd_hf[per]        # later faces win, like the reference's second loop
        partner_hf = bid_to_hf[case.bid_partner[bid[per]]]
        cell_of_hf = np.searchsorted(case.hf_offset, partner_hf, side="right") - 1
        pcell[per] = cell_of_hf
        pface[per] = partner_hf - case.hf_offset[cell_of_hf]
    return bnd_hf, kind, sat, pcell, pface


def pack_tables(case):
    n_rocks = len(case.rocks)
    off = np.zeros(n_rocks + 1, dtype=np.int32)
    for r, t in enumerate(case.rocks):
        off[r + 1] = off[r] + t.s.shape[0]
    if n_rocks == 0:
        z = np.zeros(1)
        return off, z, [z]*7
    s = np.ascontiguousarray(np.concatenate([t.s for t in case.rocks]))
    if case.mobility_kind == 0:
        cols = [np.concatenate([t.krw for t in case.rocks]), np.concatenate([t.kro for t in case.rocks]),
                np.concatenate([t.J for t in case.rocks])]
    else:
        cols = [np.concatenate([t.pc for t in case.rocks])]
        for ph in ("kr_w", "kr_o"):
            for d in range(3):
                cols.append(np.concatenate([getattr(t, ph)[:, d] for t in case.rocks]))
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    while len(cols) < 7:
        cols.append(cols[0])
    return off, s, cols


def make_fluid(case, cfl_factors=None):
    off, s, cols = pack_tables(case)
    f = _Fluid()
    f.mobility_kind = case.mobility_kind
    f.viscosity[0], f.viscosity[1] = case.visc
    f.density[0], f.density[1] = case.dens
    f.use_jfunction_scaling = int(case.use_j)
    f.sigma_cos_theta = case.sigma*np.cos(case.theta)      # RockJfunc.hpp:65-68
    f.n_rocks = len(case.rocks)
    f.table_offset = _i(off)
    f.table_s = _d(s)
    for k in range(7):
        f.table_cols[k] = _d(cols[k])
    keep = (off, s, cols)
    if cfl_factors is None:
        out = np.zeros(3)
        rid = _i(case.rock_id) if case.rock_id is not None else None
        rc = load_library().eu_compute_cfl_factors(C.byref(f), case.N, _d(case.poro), _d(case.perm), rid, _d(out))
        if rc != EU_OK:
            raise EulerB200Error(rc, "eu_compute_cfl_factors failed (pass cfl_factors for this mobility kind)")
        cfl_factors = out
    for k in range(3):
        f.cfl_factor[k] = float(cfl_factors[k])
    return f, keep


def match_periodic_faces(centroid, area, is_periodic, spatial_tolerance=1e-6):
    """findPeriodicPartners / match of the reference (BoundaryPeriodicity.hpp:86-177, .cpp:25-49) on a flat list of
    boundary faces; host only.  Returns (canon_pos, partner, side_areas)."""
    cen = np.ascontiguousarray(centroid, dtype=np.float64).reshape(-1, 3)
    ar = np.ascontiguousarray(area, dtype=np.float64)
    per = np.ascontiguousarray(is_periodic, dtype=np.int32)
    n = ar.shape[0]
    canon, partner, sides = np.zeros(max(n, 1), dtype=np.int32), np.zeros(max(n, 1), dtype=np.int32), np.zeros(6)
    rc = load_library().eu_match_periodic_faces(n, _d(cen), _d(ar), _i(per), float(spatial_tolerance), _i(canon), _i(partner), _d(sides))
    if rc != EU_OK:
        raise EulerB200Error(rc, "boundary face centroid not on the bounding box of all boundary faces")
    return canon[:n], partner[:n], sides


PARAM_KEYS = ("courant_number", "method_viscous", "method_gravity", "method_capillary", "use_cfl_viscous",
              "use_cfl_gravity", "use_cfl_capillary", "minimum_small_steps", "maximum_small_steps", "check_sat", "clamp_sat")


def params_from_case(case):
    return dict(courant_number=case.courant, method_viscous=case.method_viscous, method_gravity=case.method_gravity,
                method_capillary=case.method_capillary, use_cfl_viscous=case.use_cfl_viscous,
                use_cfl_gravity=case.use_cfl_gravity, use_cfl_capillary=case.use_cfl_capillary,
                minimum_small_steps=case.min_steps, maximum_small_steps=case.max_steps,
                check_sat=case.check_sat, clamp_sat=case.clamp_sat)


class EulerUpstream:
    """Python mirror of Opm::EulerUpstream<GridInterface, ReservoirProperties, BoundaryConditions>."""

    def __init__(self, device=0, mode="auto", rank=0, world_size=1, own_begin=0, own_end=0):
        self.L = load_library()
        cfg = _Config(1, device, MODES[mode] if isinstance(mode, str) else int(mode), rank, world_size, own_begin, own_end)
        h = C.c_void_p()
        rc = self.L.eu_create(C.byref(cfg), C.byref(h))
        if rc != EU_OK:
            raise EulerB200Error(rc, self.L.eu_last_error(None).decode())
        self.h = h
        self.par = _Params()
        self.L.eu_default_params(C.byref(self.par))
        self.case = None
        self.last_report = None

    # -- life cycle
    def close(self):
        if getattr(self, "h", None):
            self.L.eu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != EU_OK:
            raise EulerB200Error(rc, self.L.eu_last_error(self.h).decode())

    # -- EulerUpstream::init(param) (EulerUpstream_impl.hpp:95-108)
    def init(self, param=None, case=None):
        param = dict(param or {})
        for k in PARAM_KEYS:
            if k in param:
                cur = getattr(self.par, k)
                setattr(self.par, k, float(param[k]) if isinstance(cur, float) else int(param[k]))
        self._check(self.L.eu_set_params(self.h, C.byref(self.par)))
        if case is not None:
            self.initObj(case)

    def setCourantNumber(self, cn):
        self.par.courant_number = float(cn)
        self._check(self.L.eu_set_params(self.h, C.byref(self.par)))

    def display(self):
        print("\nDisplaying some members of EulerUpstream\n\ncourant_number = %g" % self.par.courant_number)

    # -- EulerUpstream::initObj(grid, resprop, boundary) (:119-127)
    def initObj(self, case, cfl_factors=None, chunk_cells=1 << 20):
        self.case = case
        N, H = case.N, case.H
        fluid, keep = make_fluid(case, cfl_factors)
        self._fluid_keep = keep
        self.cfl_factors = np.array(fluid.cfl_factor[:])
        self._check(self.L.eu_grid_begin(self.h, N, N, H))
        self._check(self.L.eu_set_fluid(self.h, C.byref(fluid)))
        bnd_hf, kind, sat, pcell, pface = resolve_boundary(case)
        counts = np.ascontiguousarray(np.diff(case.hf_offset), dtype=np.int32)
        for c0 in range(0, N, chunk_cells):
            c1 = min(N, c0 + chunk_cells)
            h0, h1 = int(case.hf_offset[c0]), int(case.hf_offset[c1])
            b0, b1 = np.searchsorted(bnd_hf, [h0, h1])
            ch = _Chunk()
            ch.first_cell, ch.n_cells = c0, c1 - c0
            cnt = counts[c0:c1]
            nbr = case.hf_nbr[h0:h1]
            rel = np.ascontiguousarray(bnd_hf[b0:b1] - h0, dtype=np.int32)
            ck, cs, cpc, cpf = (np.ascontiguousarray(a[b0:b1]) for a in (kind, sat, pcell, pface))
            ch.hf_count, ch.hf_neighbour = _i(cnt), _i(nbr)
            ch.hf_area = _d(case.hf_area[h0:h1])
            ch.hf_normal = _d(case.hf_normal[h0:h1])
            ch.hf_centroid = _d(case.hf_centroid[h0:h1])
            ch.n_bnd = int(b1 - b0)
            ch.bnd_hf, ch.bnd_kind, ch.bnd_sat = _i(rel), _i(ck), _d(cs)
            ch.bnd_partner_cell, ch.bnd_partner_face = _i(cpc), _i(cpf)
            ch.cell_volume = _d(case.cell_volume[c0:c1])
            ch.cell_centroid = _d(case.cell_centroid[c0:c1])
            ch.porosity = _d(case.poro[c0:c1])
            ch.permeability = _d(case.perm[c0:c1])
            ch.rock_id = _i(case.rock_id[c0:c1]) if case.rock_id is not None else None
            self._check(self.L.eu_grid_append(self.h, C.byref(ch)))
        self._check(self.L.eu_grid_end(self.h))

    def initObjChunks(self, fluid_case, n_global, n_local, n_local_hf, chunks, cfl_factors):
        """Streaming form of initObj for grids too large to hold as one Case: ``chunks`` yields
        dicts in eu_grid_chunk layout (global cell ids), e.g. synth.c4_slab(...).  ``cfl_factors``
        must be given (they depend on min-perm / max-poro over the whole grid)."""
        self.case = fluid_case
        fluid, keep = make_fluid(fluid_case, cfl_factors)
        self._fluid_keep = keep
        self.cfl_factors = np.array(fluid.cfl_factor[:])
        self._check(self.L.eu_grid_begin(self.h, int(n_global), int(n_local), int(n_local_hf)))
        self._check(self.L.eu_set_fluid(self.h, C.byref(fluid)))
        for d in chunks:
            ch = _Chunk()
            ch.first_cell, ch.n_cells = int(d["first_cell"]), int(d["n_cells"])
            ch.hf_count, ch.hf_neighbour = _i(d["hf_count"]), _i(d["hf_neighbour"])
            ch.hf_area, ch.hf_normal, ch.hf_centroid = _d(d["hf_area"]), _d(d["hf_normal"]), _d(d["hf_centroid"])
            ch.n_bnd = int(d["bnd_hf"].shape[0])
            ch.bnd_hf, ch.bnd_kind, ch.bnd_sat = _i(d["bnd_hf"]), _i(d["bnd_kind"]), _d(d["bnd_sat"])
            ch.bnd_partner_cell, ch.bnd_partner_face = _i(d["bnd_partner_cell"]), _i(d["bnd_partner_face"])
            ch.cell_volume, ch.cell_centroid = _d(d["cell_volume"]), _d(d["cell_centroid"])
            ch.porosity, ch.permeability = _d(d["porosity"]), _d(d["permeability"])
            ch.rock_id = _i(d["rock_id"]) if d.get("rock_id") is not None else None
            self._check(self.L.eu_grid_append(self.h, C.byref(ch)))
        self._check(self.L.eu_grid_end(self.h))

    # -- EulerUpstream::transportSolve (:151-218); host buffers in, host buffers out
    def transportSolve(self, saturation, time, gravity, hf_flux, injection_rates=None, raise_on_error=True):
        sc, sr = self._sources(injection_rates)
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        assert saturation.dtype == np.float64 and saturation.flags.c_contiguous
        fl = np.ascontiguousarray(hf_flux, dtype=np.float64)
        rep = _Report()
        rc = self.L.eu_transport_solve(self.h, _d(saturation), float(time), _d(g), _d(fl), sc.shape[0], _i(sc), _d(sr), C.byref(rep))
        self.last_report = self._report(rep)
        if rc != EU_OK and raise_on_error:
            raise EulerB200Error(rc, self.L.eu_last_error(self.h).decode())
        return self.last_report

    # -- device-resident variant
    def upload_state(self, saturation, hf_flux):
        """saturation=None keeps the resident state and replaces the fluxes only (the IMPES loop between pressure solves)."""
        s = None if saturation is None else np.ascontiguousarray(saturation, dtype=np.float64)
        fl = np.ascontiguousarray(hf_flux, dtype=np.float64)
        self._check(self.L.eu_upload_state(self.h, None if s is None else _d(s), _d(fl)))

    def upload_saturation(self, saturation):
        s = np.ascontiguousarray(saturation, dtype=np.float64)
        self._check(self.L.eu_upload_saturation(self.h, _d(s)))

    def download_saturation(self):
        out = np.empty(self.L.eu_local_cells(self.h))
        self._check(self.L.eu_download_saturation(self.h, _d(out)))
        return out

    def transportSolveResident(self, time, gravity, injection_rates=None, raise_on_error=True):
        sc, sr = self._sources(injection_rates)
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        rep = _Report()
        rc = self.L.eu_transport_solve_resident(self.h, float(time), _d(g), sc.shape[0], _i(sc), _d(sr), C.byref(rep))
        self.last_report = self._report(rep)
        if rc != EU_OK and raise_on_error:
            raise EulerB200Error(rc, self.L.eu_last_error(self.h).decode())
        return self.last_report

    def resolved_mode(self):
        """'strict' or 'fast': what EU_MODE_AUTO resolved to for this grid and property class."""
        return {1: "strict", 2: "fast"}.get(int(self.L.eu_resolved_mode(self.h)), "auto")

    def regular_fraction(self):
        return float(self.L.eu_regular_fraction(self.h))

    def work_plan(self):
        """Work plan of the FAST substep kernel after the last substep: class fraction, items, longest / mean march."""
        out = np.zeros(5)
        self._check(self.L.eu_work_plan(self.h, _d(out)))
        return {"class_fraction": float(out[0]), "items": int(out[1]), "max_march": int(out[2]), "mean_march": float(out[3]),
                "kernel": "box" if out[4] == 1.0 else "slice-class"}

    def cfl_times(self, gravity):
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        out = np.zeros(3)
        self._check(self.L.eu_cfl_times(self.h, _d(g), _d(out)))
        return out

    def small_step(self, dt, gravity, injection_rates=None):
        sc, sr = self._sources(injection_rates)
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        res = np.zeros(self.L.eu_local_cells(self.h))
        bc, bv = C.c_int(-1), C.c_double(0.0)
        rc = self.L.eu_small_step(self.h, float(dt), _d(g), sc.shape[0], _i(sc), _d(sr), _d(res), C.byref(bc), C.byref(bv))
        if rc not in (EU_OK, EU_ERR_SAT_RANGE):
            self._check(rc)
        return dict(status=rc, residual=res, bad_cell=bc.value, bad_value=bv.value)

    # -- EulerUpstreamResidual::computeResidual (Residual_impl.hpp:472-505) as an operator; hf_flux=None reuses
    #    the resident fluxes
    def computeResidual(self, saturation, gravity, hf_flux, injection_rates, method_viscous, method_gravity, method_capillary):
        sc, sr = self._sources(injection_rates)
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        s = np.ascontiguousarray(saturation, dtype=np.float64)
        fl = None if hf_flux is None else np.ascontiguousarray(hf_flux, dtype=np.float64)
        out = np.zeros(self.L.eu_local_cells(self.h))
        self._check(self.L.eu_compute_residual(self.h, _d(s), _d(g), None if fl is None else _d(fl), sc.shape[0], _i(sc), _d(sr),
                                               int(bool(method_viscous)), int(bool(method_gravity)), int(bool(method_capillary)),
                                               _d(out)))
        return out

    # -- EulerUpstreamResidual::computeCapPressures (:459-467) / computeCapPressure (SimulatorUtilities.hpp:219-230)
    def computeCapPressures(self, saturation):
        s = np.ascontiguousarray(saturation, dtype=np.float64)
        out = np.zeros(self.L.eu_local_cells(self.h))
        self._check(self.L.eu_compute_cap_pressures(self.h, _d(s), _d(out)))
        return out

    # -- diagnostics (SimulatorUtilities.hpp:59-86, :153-170, :273-279) on the resident fluxes
    def cellVelocity(self):
        out = np.zeros((self.L.eu_local_cells(self.h), 3))
        self._check(self.L.eu_cell_velocity(self.h, _d(out)))
        return out

    def phaseVelocities(self, saturation=None, cell_velocity=None):
        n = self.L.eu_local_cells(self.h)
        s = None if saturation is None else np.ascontiguousarray(saturation, dtype=np.float64)
        cv = None if cell_velocity is None else np.ascontiguousarray(cell_velocity, dtype=np.float64)
        vw, vo = np.zeros((n, 3)), np.zeros((n, 3))
        self._check(self.L.eu_phase_velocities(self.h, None if s is None else _d(s), None if cv is None else _d(cv), _d(vw), _d(vo)))
        return vw, vo

    def fractionalFlow(self, saturation=None):
        s = None if saturation is None else np.ascontiguousarray(saturation, dtype=np.float64)
        out = np.zeros(self.L.eu_local_cells(self.h))
        self._check(self.L.eu_fractional_flow(self.h, None if s is None else _d(s), _d(out)))
        return out

    @staticmethod
    def _sources(inj):
        if inj is None:
            return np.zeros(0, dtype=np.int32), np.zeros(0)
        sc, sr = inj
        return np.ascontiguousarray(sc, dtype=np.int32), np.ascontiguousarray(sr, dtype=np.float64)

    @staticmethod
    def _report(r):
        return Report(r.status, r.nsteps, r.attempts, r.substeps_executed, r.bad_cell, r.bad_value,
                      tuple(r.cfl_dt[:]), r.dt, r.device_ms, r.kernel_launches)
