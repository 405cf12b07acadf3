"""GPU parity against the CPU ORACLE (not against the device's own STRICT mode) at the BASELINE.json sizes and on the
code path bench.py times: the box kernel of FAST mode (TMA-staged plane sweep over tiles, eu_tile.cuh) -- and, with
EU_BOX=0, the slice-class kernel with its z-march that it replaced and that still serves grids without a box numbering.

  C4 slab  512 x 512 x 4 and 64 x 32 x 8, viscous + gravity, 1 rock   (the bench workload's plane shape)
  C2       100^3, rotated anisotropic K, rock table, capillary on       (BASELINE config 1, full size)
  C3       256 x 256 x 128 faulted corner-point, 3 rocks, V+G+C         (BASELINE config 2, full size: 2 substeps)

Gates (BASELINE.json north_star): FAST max|dS| <= 1e-12 per substep, <= 1e-9 per transportSolve, identical CFL times
and step counts; STRICT bit-identical.  The oracle (oracle/euler_oracle.c) is pinned bit for bit to the compiled
reference headers on the CPU (tests/test_oracle_vs_reference.py).  The CPU side of this file costs about a minute."""
import numpy as np
import pytest

from conftest import active_cfl_dt

pytestmark = pytest.mark.gpu

TOL_SUBSTEP = 1e-12
TOL_SOLVE = 1e-9


def _oracle_run(case, n_sub, solve_steps, cfl_fraction=0.5):
    """CFL factors / times, n_sub substeps on the oracle's own trajectory, and one transportSolve."""
    from oracle.ref import PortSolver
    port = PortSolver(case)
    fac = port.compute_cfl_factors()
    cfl = port.cfl_times()
    total = active_cfl_dt(case, cfl)
    dt = cfl_fraction*total
    s = case.sat0.copy()
    steps = []
    for _ in range(n_sub):
        o = port.small_step(s, dt)
        assert o["status"] == 0
        s = o["sat"]
        steps.append((o["sat"].copy(), o["residual"].copy()))
    sol = None
    if solve_steps:
        time = (solve_steps - 0.7)*total
        sol = port.transport_solve(case.sat0, time=time)
        assert sol["status"] == 0 and sol["nsteps"] == solve_steps
        sol["time"] = time
    return dict(fac=fac, cfl=cfl, dt=dt, steps=steps, sol=sol)


def _device_check(case, ref, mode, expect_march):
    from opm_porsol_b200 import EulerUpstream
    from opm_porsol_b200.binding import params_from_case
    dev = EulerUpstream(device=0, mode=mode)
    dev.init(params_from_case(case))
    dev.initObj(case, cfl_factors=ref["fac"])
    dev.upload_state(case.sat0, case.hf_flux)
    assert np.array_equal(dev.cfl_times(case.gravity), ref["cfl"])
    inj = (case.src_cell, case.src_rate)
    worst = 0.0
    for sat_ref, res_ref in ref["steps"]:
        o = dev.small_step(ref["dt"], case.gravity, inj)
        s = dev.download_saturation()
        assert o["status"] == 0
        if mode == "strict":
            assert np.array_equal(o["residual"], res_ref)
            assert np.array_equal(s, sat_ref)
        else:
            err = float(np.abs(s - sat_ref).max())
            worst = max(worst, err)
            assert err <= TOL_SUBSTEP, err
            assert np.abs(o["residual"] - res_ref).max() <= 1e-12*(np.abs(res_ref).max() + 1e-300)
        dev.upload_saturation(sat_ref)            # every substep is an independent comparison
    if mode == "fast" and expect_march:
        plan = dev.work_plan()
        if expect_march == "box":
            # the box kernel is what ran: every own cell swept by a tile, units longer than one plane
            assert plan["kernel"] == "box" and plan["items"] > 0, plan
        else:
            # EU_BOX=0: the slice-class kernel; its marches need slice-aligned planes
            assert plan["kernel"] == "slice-class" and plan["class_fraction"] > 0.5 and plan["max_march"] >= 1, plan
    sol = ref["sol"]
    if sol is not None:
        sat = case.sat0.copy()
        rep = dev.transportSolve(sat, sol["time"], case.gravity, case.hf_flux, inj)
        assert rep.status == 0 and rep.nsteps == sol["nsteps"] and rep.attempts == sol["attempts"]
        assert np.array_equal(np.array(rep.cfl_dt), sol["cfl_dt"])
        if mode == "strict":
            assert np.array_equal(sat, sol["sat"])
        else:
            assert np.abs(sat - sol["sat"]).max() <= TOL_SOLVE
    dev.close()
    return worst


@pytest.mark.parametrize("dims", [(64, 32, 8), (512, 512, 4)], ids=["64x32x8", "512x512x4"])
def test_c4_slab_march_path_vs_oracle(dims, monkeypatch):
    """The bench workload's kernel instantiation (1 rock, no capillary term, sweep along z) against the oracle: the box
    kernel (default) and the slice-class kernel (EU_BOX=0)."""
    from opm_porsol_b200 import synth
    case = synth.config_c4(*dims)
    ref = _oracle_run(case, n_sub=4, solve_steps=18 if dims[0] < 512 else 6)
    for mode in ("fast", "strict"):
        _device_check(case, ref, mode, expect_march="box")
    monkeypatch.setenv("EU_BOX", "0")
    _device_check(case, ref, "fast", expect_march="classes")


def test_c4_slab_capillary_march_vs_oracle():
    """The same plane shape with the capillary term (the C4+capillary bench variant)."""
    from opm_porsol_b200 import synth
    case = synth.config_c4(256, 128, 6, capillary=True)
    ref = _oracle_run(case, n_sub=3, solve_steps=8)
    for mode in ("fast", "strict"):
        _device_check(case, ref, mode, expect_march="box")


@pytest.mark.parametrize("units", ["chunks", "spans"])
def test_rock_ids_drawn_per_cell_vs_oracle(units, monkeypatch):
    """Three rock types drawn per cell on a faulted grid of several tiles: the rock ids of a tile's halo cells differ from
    its own cells' in every plane, from plane 0 on (the layer bands of C3 do not test that).  Also with the other way of
    cutting the sweep into work units (EU_BOX_UNITS=spans: units of unequal length, several per tile column)."""
    from opm_porsol_b200 import synth
    monkeypatch.setenv("EU_BOX_UNITS", units)
    case = synth.config_c3(96, 40, 10, random_rocks=True)
    assert len(np.unique(case.rock_id[:96*40])) == 3
    ref = _oracle_run(case, n_sub=3, solve_steps=6, cfl_fraction=0.25)
    for mode in ("fast", "strict"):
        _device_check(case, ref, mode, expect_march="box" if mode == "fast" else False)


def test_c2_full_size_vs_oracle():
    """BASELINE config 1 at its stated size: 100^3 Cartesian, rotated anisotropic K, rock table, V+G+C."""
    from opm_porsol_b200 import synth
    case = synth.config_c2(100)
    ref = _oracle_run(case, n_sub=3, solve_steps=18, cfl_fraction=0.25)
    _device_check(case, ref, "fast", expect_march="box")          # 100 x 100 planes: no slice alignment needed
    _device_check(case, ref, "strict", expect_march=False)


def test_c3_full_size_vs_oracle(monkeypatch):
    """BASELINE config 2 at its stated size: 256 x 256 x 128 faulted corner-point, lognormal K, 3 rocks, V+G+C:
    two substeps of the whole 8.4 M-cell grid against the oracle."""
    from opm_porsol_b200 import synth
    case = synth.config_c3(256, 256, 128)
    ref = _oracle_run(case, n_sub=2, solve_steps=0, cfl_fraction=0.25)
    _device_check(case, ref, "fast", expect_march="box")          # fault faces through the pre-pass kernel
    _device_check(case, ref, "strict", expect_march=False)
    monkeypatch.setenv("EU_BOX_UNITS", "spans")                   # units of up to 110 planes that cross the rock bands
    _device_check(case, ref, "fast", expect_march="box")
