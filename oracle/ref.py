"""TEST INFRASTRUCTURE -- ctypes front-ends for the two parity oracles.

  RefSolver   : oracle/_ref/libeuler_ref.so -- the UNMODIFIED reference headers compiled from
                /root/reference (oracle/ref_harness.cpp).  Prebuilt here; travels to the GPU box.
  PortSolver  : oracle/liboracle_port.so -- the plain-C restatement (oracle/euler_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (opm-porsol_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libeuler_ref.so")
PORT_LIB = os.path.join(HERE, "liboracle_port.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def port_available() -> bool:
    return os.path.exists(PORT_LIB)


def pack_tables(case):
    """Concatenate the per-rock tables: returns (offsets, s, [columns...])."""
    n_rocks = len(case.rocks)
    off = np.zeros(n_rocks + 1, dtype=np.int32)
    for r, t in enumerate(case.rocks):
        off[r + 1] = off[r] + t.s.shape[0]
    if n_rocks == 0:
        z = np.zeros(1)
        return off, z, [z, z, z]
    s = np.concatenate([t.s for t in case.rocks])
    if case.mobility_kind == 0:
        cols = [np.concatenate([t.krw for t in case.rocks]),
                np.concatenate([t.kro for t in case.rocks]),
                np.concatenate([t.J for t in case.rocks])]
    else:
        cols = [np.concatenate([t.pc for t in case.rocks])]
        for ph in ("kr_w", "kr_o"):
            for d in range(3):
                cols.append(np.concatenate([getattr(t, ph)[:, d] for t in case.rocks]))
    return off, np.ascontiguousarray(s), [np.ascontiguousarray(c, dtype=np.float64) for c in cols]


class RefSolver:
    """The reference EulerUpstream<FlatGrid, ReservoirPropertyCapillary<3>|...AnisotropicRelperm<3>,
    BasicBoundaryConditions<true,true>> behind a C API."""

    def __init__(self, case):
        self.lib = C.CDLL(REF_LIB)
        L = self.lib
        L.ref_create.restype = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_last_error.argtypes = [C.c_void_p]
        L.ref_cap_pressure.restype = C.c_double
        L.ref_cap_pressure.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ref_frac_flow.restype = C.c_double
        L.ref_frac_flow.argtypes = [C.c_void_p, C.c_int, C.c_double]
        self.case = case
        self._tmp = tempfile.TemporaryDirectory(prefix="eu_ref_")
        off, s, cols = pack_tables(case)
        colptrs = (_dp*len(cols))(*[_d(c) for c in cols])
        visc = np.asarray(case.visc, dtype=np.float64)
        dens = np.asarray(case.dens, dtype=np.float64)
        self._keep = (off, s, cols, visc, dens)
        rid = _i(case.rock_id) if case.rock_id is not None else None
        self.h = L.ref_create(
            C.c_int(case.N), _i(case.hf_offset), _i(case.hf_nbr), _i(case.hf_bid),
            _d(case.hf_area), _d(case.hf_normal), _d(case.hf_centroid),
            _d(case.cell_volume), _d(case.cell_centroid),
            _d(case.poro), _d(case.perm),
            rid, C.c_int(len(case.rocks)), _i(off), _d(s), colptrs, C.c_int(len(cols)),
            C.c_int(int(case.use_j)), C.c_double(case.sigma), C.c_double(case.theta),
            _d(visc), _d(dens),
            C.c_int(case.bid_kind.shape[0]), _i(case.bid_kind), _d(case.bid_sat), _i(case.bid_partner),
            C.c_int(case.mobility_kind), self._tmp.name.encode())
        if not self.h:
            raise RuntimeError("ref_create failed")
        self.h = C.c_void_p(self.h)
        self.set_params(case)

    def set_params(self, case):
        self.lib.ref_set_params(self.h, C.c_double(case.courant), int(case.method_viscous), int(case.method_gravity),
                                int(case.method_capillary), int(case.use_cfl_viscous), int(case.use_cfl_gravity),
                                int(case.use_cfl_capillary), int(case.min_steps), int(case.max_steps),
                                int(case.check_sat), int(case.clamp_sat))

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None
        self._tmp.cleanup()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self):
        return self.lib.ref_last_error(self.h).decode()

    def transport_solve(self, sat, time=None, gravity=None, hf_flux=None, src_cell=None, src_rate=None):
        c = self.case
        sat = np.array(sat, dtype=np.float64)
        g = np.asarray(c.gravity if gravity is None else gravity, dtype=np.float64)
        fl = np.ascontiguousarray(c.hf_flux if hf_flux is None else hf_flux, dtype=np.float64)
        sc = c.src_cell if src_cell is None else np.ascontiguousarray(src_cell, dtype=np.int32)
        sr = c.src_rate if src_rate is None else np.ascontiguousarray(src_rate, dtype=np.float64)
        ns, at, secs = C.c_int(0), C.c_int(0), C.c_double(0)
        st = self.lib.ref_transport_solve(self.h, _d(sat), C.c_double(c.time if time is None else time), _d(g), _d(fl),
                                          C.c_int(sc.shape[0]), _i(sc), _d(sr), C.byref(ns), C.byref(at), C.byref(secs))
        return dict(sat=sat, status=st, nsteps=ns.value, attempts=at.value, seconds=secs.value,
                    error=self.last_error() if st else "")

    def small_step(self, sat, dt, gravity=None, hf_flux=None, src_cell=None, src_rate=None):
        c = self.case
        sat = np.array(sat, dtype=np.float64)
        res = np.zeros(c.N)
        g = np.asarray(c.gravity if gravity is None else gravity, dtype=np.float64)
        fl = np.ascontiguousarray(c.hf_flux if hf_flux is None else hf_flux, dtype=np.float64)
        sc = c.src_cell if src_cell is None else np.ascontiguousarray(src_cell, dtype=np.int32)
        sr = c.src_rate if src_rate is None else np.ascontiguousarray(src_rate, dtype=np.float64)
        st = self.lib.ref_small_step(self.h, _d(sat), C.c_double(dt), _d(g), _d(fl),
                                     C.c_int(sc.shape[0]), _i(sc), _d(sr), _d(res))
        return dict(sat=sat, residual=res, status=st, error=self.last_error() if st else "")

    def cfl_times(self, gravity=None, hf_flux=None):
        c = self.case
        g = np.asarray(c.gravity if gravity is None else gravity, dtype=np.float64)
        fl = np.ascontiguousarray(c.hf_flux if hf_flux is None else hf_flux, dtype=np.float64)
        out = np.zeros(3)
        tot = C.c_double(0)
        self.lib.ref_cfl_times(self.h, _d(g), _d(fl), _d(out), C.byref(tot))
        return out, tot.value

    def cfl_factors(self):
        out = np.zeros(3)
        self.lib.ref_cfl_factors(self.h, _d(out))
        return out

    def compute_residual(self, sat, methods, gravity=None, hf_flux=None, src_cell=None, src_rate=None):
        """EulerUpstreamResidual::computeResidual with explicit (viscous, gravity, capillary) flags."""
        c = self.case
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        g = np.asarray(c.gravity if gravity is None else gravity, dtype=np.float64)
        fl = np.ascontiguousarray(c.hf_flux if hf_flux is None else hf_flux, dtype=np.float64)
        sc = c.src_cell if src_cell is None else np.ascontiguousarray(src_cell, dtype=np.int32)
        sr = c.src_rate if src_rate is None else np.ascontiguousarray(src_rate, dtype=np.float64)
        out = np.zeros(c.N)
        self.lib.ref_compute_residual(self.h, _d(sat), _d(g), _d(fl), C.c_int(sc.shape[0]), _i(sc), _d(sr),
                                      C.c_int(int(methods[0])), C.c_int(int(methods[1])), C.c_int(int(methods[2])), _d(out))
        return out

    def cell_velocity(self, hf_flux=None):
        fl = np.ascontiguousarray(self.case.hf_flux if hf_flux is None else hf_flux, dtype=np.float64)
        out = np.zeros((self.case.N, 3))
        self.lib.ref_cell_velocity(self.h, _d(fl), _d(out))
        return out

    def phase_velocities(self, sat, cell_velocity):
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        cv = np.ascontiguousarray(cell_velocity, dtype=np.float64)
        vw, vo = np.zeros_like(cv), np.zeros_like(cv)
        self.lib.ref_phase_velocities(self.h, _d(sat), _d(cv), _d(vw), _d(vo))
        return vw, vo

    def cap_pressures(self, sat):
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        out = np.zeros(self.case.N)
        self.lib.ref_cap_pressures(self.h, _d(sat), _d(out))
        return out

    def cap_pressure(self, cell, s):
        return self.lib.ref_cap_pressure(self.h, int(cell), float(s))

    def frac_flow(self, cell, s):
        return self.lib.ref_frac_flow(self.h, int(cell), float(s))

    def mobility(self, phase, cell, s):
        out = np.zeros(9)
        self.lib.ref_mobility(self.h, int(phase), int(cell), C.c_double(s), _d(out))
        return out


def ref_find_periodic_partners(centroid, area, is_periodic, spatial_tolerance=1e-6):
    """The reference's findPeriodicPartners + match (BoundaryPeriodicity.hpp:86-177, .cpp:25-49) over a mock GridView.
    Returns (status, canon_pos, partner, side_areas); status 1 = the reference threw."""
    lib = C.CDLL(REF_LIB)
    cen = np.ascontiguousarray(centroid, dtype=np.float64).reshape(-1, 3)
    ar = np.ascontiguousarray(area, dtype=np.float64)
    per = np.ascontiguousarray(is_periodic, dtype=np.int32)
    n = ar.shape[0]
    canon, partner, sides = np.full(max(n, 1), -9, dtype=np.int32), np.full(max(n, 1), -9, dtype=np.int32), np.zeros(6)
    st = lib.ref_find_periodic_partners(C.c_int(n), _d(cen), _d(ar), _i(per), C.c_double(spatial_tolerance),
                                        _i(canon), _i(partner), _d(sides))
    return st, canon[:n], partner[:n], sides


def ref_write_field(field, filename):
    """The reference's writeField (SimulatorUtilities.hpp:288-298)."""
    lib = C.CDLL(REF_LIB)
    f = np.ascontiguousarray(field, dtype=np.float64)
    return lib.ref_write_field(_d(f), C.c_int(f.shape[0]), C.c_char_p(filename.encode()))


# ------------------------------------------------------------------------------------
# plain-C restatement
# ------------------------------------------------------------------------------------
class _EoCase(C.Structure):
    _fields_ = [
        ("N", C.c_int),
        ("hf_offset", _ip), ("hf_nbr", _ip), ("hf_bid", _ip),
        ("hf_area", _dp), ("hf_normal", _dp), ("hf_centroid", _dp),
        ("cell_volume", _dp), ("cell_centroid", _dp),
        ("poro", _dp), ("perm", _dp), ("rock_id", _ip),
        ("n_rocks", C.c_int), ("tab_offset", _ip), ("tab_s", _dp), ("tab_cols", _dp*7),
        ("mobility_kind", C.c_int), ("use_j", C.c_int), ("sigma_cos_theta", C.c_double),
        ("visc", C.c_double*2), ("dens", C.c_double*2), ("cfl_factor", C.c_double*3),
        ("n_bid", C.c_int), ("bid_kind", _ip), ("bid_sat", _dp), ("bid_partner", _ip),
        ("courant", C.c_double),
        ("method_viscous", C.c_int), ("method_gravity", C.c_int), ("method_capillary", C.c_int),
        ("use_cfl_viscous", C.c_int), ("use_cfl_gravity", C.c_int), ("use_cfl_capillary", C.c_int),
        ("min_steps", C.c_int), ("max_steps", C.c_int), ("check_sat", C.c_int), ("clamp_sat", C.c_int),
    ]


class _EoResult(C.Structure):
    _fields_ = [("status", C.c_int), ("nsteps", C.c_int), ("attempts", C.c_int),
                ("substeps_executed", C.c_longlong), ("bad_cell", C.c_int), ("bad_value", C.c_double),
                ("cfl_dt", C.c_double*3), ("seconds", C.c_double)]


class PortSolver:
    """oracle/euler_oracle.c driven with the same flat arrays. ``cfl_factors`` default to the
    port's own restatement of computeCflFactors (scalar mobility) -- pass the reference's to
    isolate the transport arithmetic."""

    def __init__(self, case, cfl_factors=None):
        self.lib = C.CDLL(PORT_LIB)
        L = self.lib
        L.eo_cfl_gravity.restype = C.c_double
        L.eo_cfl_capillary.restype = C.c_double
        L.eo_cap_pressure.restype = C.c_double
        L.eo_fractional_flow.restype = C.c_double
        self.case = case
        off, s, cols = pack_tables(case)
        self._keep = (off, s, cols)
        ec = _EoCase()
        ec.N = case.N
        ec.hf_offset, ec.hf_nbr, ec.hf_bid = _i(case.hf_offset), _i(case.hf_nbr), _i(case.hf_bid)
        ec.hf_area, ec.hf_normal, ec.hf_centroid = _d(case.hf_area), _d(case.hf_normal), _d(case.hf_centroid)
        ec.cell_volume, ec.cell_centroid = _d(case.cell_volume), _d(case.cell_centroid)
        ec.poro, ec.perm = _d(case.poro), _d(case.perm)
        ec.rock_id = _i(case.rock_id) if case.rock_id is not None else None
        ec.n_rocks = len(case.rocks)
        ec.tab_offset, ec.tab_s = _i(off), _d(s)
        for k, col in enumerate(cols):
            ec.tab_cols[k] = _d(col)
        ec.mobility_kind = case.mobility_kind
        ec.use_j = int(case.use_j)
        ec.sigma_cos_theta = case.sigma*np.cos(case.theta)     # RockJfunc.hpp:65-68
        ec.visc[0], ec.visc[1] = case.visc
        ec.dens[0], ec.dens[1] = case.dens
        ec.n_bid = case.bid_kind.shape[0]
        ec.bid_kind, ec.bid_sat, ec.bid_partner = _i(case.bid_kind), _d(case.bid_sat), _i(case.bid_partner)
        self.ec = ec
        self.set_params(case)
        if cfl_factors is None:
            cfl_factors = self.compute_cfl_factors()
        for k in range(3):
            ec.cfl_factor[k] = cfl_factors[k]

    def set_params(self, case):
        ec = self.ec
        ec.courant = case.courant
        ec.method_viscous, ec.method_gravity, ec.method_capillary = int(case.method_viscous), int(case.method_gravity), int(case.method_capillary)
        ec.use_cfl_viscous, ec.use_cfl_gravity, ec.use_cfl_capillary = int(case.use_cfl_viscous), int(case.use_cfl_gravity), int(case.use_cfl_capillary)
        ec.min_steps, ec.max_steps = case.min_steps, case.max_steps
        ec.check_sat, ec.clamp_sat = int(case.check_sat), int(case.clamp_sat)

    def compute_cfl_factors(self):
        out = np.zeros(3)
        self.lib.eo_compute_cfl_factors(C.byref(self.ec), _d(out))
        return out

    def _inputs(self, gravity, hf_flux, src_cell, src_rate):
        c = self.case
        g = np.asarray(c.gravity if gravity is None else gravity, dtype=np.float64)
        fl = np.ascontiguousarray(c.hf_flux if hf_flux is None else hf_flux, dtype=np.float64)
        sc = c.src_cell if src_cell is None else np.ascontiguousarray(src_cell, dtype=np.int32)
        sr = c.src_rate if src_rate is None else np.ascontiguousarray(src_rate, dtype=np.float64)
        return g, fl, sc, sr

    def transport_solve(self, sat, time=None, gravity=None, hf_flux=None, src_cell=None, src_rate=None):
        c = self.case
        sat = np.array(sat, dtype=np.float64)
        g, fl, sc, sr = self._inputs(gravity, hf_flux, src_cell, src_rate)
        res = _EoResult()
        self.lib.eo_transport_solve(C.byref(self.ec), _d(sat), C.c_double(c.time if time is None else time), _d(g), _d(fl),
                                    C.c_int(sc.shape[0]), _i(sc), _d(sr), C.byref(res))
        return dict(sat=sat, status=res.status, nsteps=res.nsteps, attempts=res.attempts,
                    substeps_executed=res.substeps_executed, bad_cell=res.bad_cell, bad_value=res.bad_value,
                    cfl_dt=np.array(res.cfl_dt[:]), seconds=res.seconds)

    def small_step(self, sat, dt, gravity=None, hf_flux=None, src_cell=None, src_rate=None):
        c = self.case
        sat = np.array(sat, dtype=np.float64)
        g, fl, sc, sr = self._inputs(gravity, hf_flux, src_cell, src_rate)
        cap = np.zeros(c.N)
        res = np.zeros(c.N)
        bc, bv = C.c_int(-1), C.c_double(0)
        st = self.lib.eo_small_step(C.byref(self.ec), _d(sat), C.c_double(dt), _d(g), _d(fl), C.c_int(sc.shape[0]),
                                    _i(sc), _d(sr), _d(cap), _d(res), C.byref(bc), C.byref(bv))
        return dict(sat=sat, residual=res, status=st, bad_cell=bc.value, bad_value=bv.value, cap=cap)

    def compute_residual(self, sat, methods, gravity=None, hf_flux=None, src_cell=None, src_rate=None):
        """eo_compute_residual with explicit (viscous, gravity, capillary) flags; the case's own flags are restored."""
        c = self.case
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        g, fl, sc, sr = self._inputs(gravity, hf_flux, src_cell, src_rate)
        ec = self.ec
        saved = (ec.method_viscous, ec.method_gravity, ec.method_capillary)
        ec.method_viscous, ec.method_gravity, ec.method_capillary = (int(bool(m)) for m in methods)
        cap, res = np.zeros(c.N), np.zeros(c.N)
        try:
            self.lib.eo_compute_residual(C.byref(ec), _d(sat), _d(g), _d(fl), C.c_int(sc.shape[0]), _i(sc), _d(sr), _d(cap), _d(res))
        finally:
            ec.method_viscous, ec.method_gravity, ec.method_capillary = saved
        return res

    def cell_velocity(self, hf_flux=None):
        fl = np.ascontiguousarray(self.case.hf_flux if hf_flux is None else hf_flux, dtype=np.float64)
        out = np.zeros((self.case.N, 3))
        self.lib.eo_cell_velocity(C.byref(self.ec), _d(fl), _d(out))
        return out

    def phase_velocities(self, sat, cell_velocity):
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        cv = np.ascontiguousarray(cell_velocity, dtype=np.float64)
        vw, vo = np.zeros_like(cv), np.zeros_like(cv)
        self.lib.eo_phase_velocities(C.byref(self.ec), _d(sat), _d(cv), _d(vw), _d(vo))
        return vw, vo

    def cap_pressures(self, sat):
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        out = np.zeros(self.case.N)
        self.lib.eo_cap_pressures(C.byref(self.ec), _d(sat), _d(out))
        return out

    def frac_flows(self, sat):
        return np.array([self.frac_flow(c, s) for c, s in enumerate(np.asarray(sat, dtype=np.float64))])

    def cfl_times(self, gravity=None, hf_flux=None):
        g, fl, _, _ = self._inputs(gravity, hf_flux, None, None)
        v = C.c_double(0)
        st = self.lib.eo_cfl_velocity(C.byref(self.ec), _d(fl), C.byref(v))
        return np.array([v.value if st == 0 else np.nan, self.lib.eo_cfl_gravity(C.byref(self.ec), _d(g)),
                         self.lib.eo_cfl_capillary(C.byref(self.ec))])

    def cap_pressure(self, cell, s):
        return self.lib.eo_cap_pressure(C.byref(self.ec), C.c_int(cell), C.c_double(s))

    def mobility(self, phase, cell, s):
        out = np.zeros(9)
        self.lib.eo_mobility(C.byref(self.ec), C.c_int(phase), C.c_int(cell), C.c_double(s), _d(out))
        if self.case.mobility_kind == 0:
            out[4] = out[8] = out[0]
        return out

    def frac_flow(self, cell, s):
        return self.lib.eo_fractional_flow(C.byref(self.ec), C.c_int(cell), C.c_double(s))
