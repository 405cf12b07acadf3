"""Golden vectors generated from the reference itself (tests/golden/make_golden.py; the reference ships
none for this path).  CPU: the C oracle reproduces them bit for bit.  GPU: STRICT mode
reproduces them bit for bit, FAST mode within 1e-12 per substep / 1e-9 per transportSolve."""
import os

import numpy as np
import pytest

from conftest import ROOT, small_cases, tensor_cases

ALL = small_cases() + tensor_cases()


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.mark.parametrize("name,case", ALL, ids=[n for n, _ in ALL])
def test_oracle_reproduces_golden(name, case):
    from oracle.ref import PortSolver
    g = golden(name)
    chk = np.array([case.sat0.sum(), case.hf_flux.sum(), case.perm.sum(), case.hf_area.sum()])
    assert np.array_equal(chk, g["input_checksum"]), "seeded inputs differ from the ones the fixture was made with"
    port = PortSolver(case, cfl_factors=g["cfl_factors"])
    if case.mobility_kind == 0:
        assert np.array_equal(port.compute_cfl_factors(), g["cfl_factors"])
    assert np.array_equal(port.cfl_times(), g["cfl_times"])
    s = case.sat0.copy()
    for q in range(3):
        o = port.small_step(s, float(g["dt"]))
        s = o["sat"]
        assert np.array_equal(o["residual"], g["step_res"][q])
        assert np.array_equal(s, g["step_sat"][q])
    assert np.array_equal(port.compute_residual(case.sat0, (True, True, False)), g["res_vg"])
    assert np.array_equal(port.compute_residual(case.sat0, (False, True, True)), g["res_gc"])
    assert np.array_equal(port.cell_velocity(), g["cell_velocity"])
    assert np.array_equal(port.cap_pressures(case.sat0), g["cap_pressures"])
    vw, vo = port.phase_velocities(case.sat0, g["cell_velocity"])
    assert np.array_equal(vw, g["water_velocity"]) and np.array_equal(vo, g["oil_velocity"])
    if case.mobility_kind == 0:
        sol = port.transport_solve(case.sat0, time=float(g["solve_time"]))
        assert sol["nsteps"] == int(g["solve_nsteps"]) and sol["attempts"] == int(g["solve_attempts"])
        assert np.array_equal(sol["sat"], g["solve_sat"])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("name,case", ALL, ids=[n for n, _ in ALL])
def test_cuda_reproduces_golden(name, case, mode):
    from opm_porsol_b200 import EulerUpstream
    from opm_porsol_b200.binding import params_from_case
    g = golden(name)
    dev = EulerUpstream(device=0, mode=mode)
    dev.init(params_from_case(case))
    dev.initObj(case, cfl_factors=g["cfl_factors"])
    dev.upload_state(case.sat0, case.hf_flux)
    assert np.array_equal(dev.cfl_times(case.gravity), g["cfl_times"])
    inj = (case.src_cell, case.src_rate)
    for q in range(3):
        o = dev.small_step(float(g["dt"]), case.gravity, inj)
        s = dev.download_saturation()
        if mode == "strict":
            assert np.array_equal(o["residual"], g["step_res"][q])
            assert np.array_equal(s, g["step_sat"][q])
        else:
            assert np.abs(s - g["step_sat"][q]).max() <= 1e-12
        dev.upload_saturation(g["step_sat"][q])
    # operator entry points against the reference's own outputs
    for key, m in (("res_vg", (True, True, False)), ("res_gc", (False, True, True))):
        r = dev.computeResidual(case.sat0, case.gravity, None, inj, *m)
        if mode == "strict":
            assert np.array_equal(r, g[key])
        else:
            assert np.abs(r - g[key]).max() <= 1e-12*(np.abs(g[key]).max() + 1e-300)
    assert np.array_equal(dev.cellVelocity(), g["cell_velocity"])
    assert np.array_equal(dev.computeCapPressures(case.sat0), g["cap_pressures"])
    vw, vo = dev.phaseVelocities(case.sat0, g["cell_velocity"])
    assert np.array_equal(vw, g["water_velocity"]) and np.array_equal(vo, g["oil_velocity"])
    if case.mobility_kind == 0:
        sat = case.sat0.copy()
        rep = dev.transportSolve(sat, float(g["solve_time"]), case.gravity, case.hf_flux, inj)
        assert rep.nsteps == int(g["solve_nsteps"]) and rep.attempts == int(g["solve_attempts"])
        if mode == "strict":
            assert np.array_equal(sat, g["solve_sat"])
        else:
            assert np.abs(sat - g["solve_sat"]).max() <= 1e-9
    dev.close()


def test_survey_known_answer():
    """The known-answer case SURVEY 8c records from a separate probe of the unmodified reference headers (mock Cartesian
    10x10x10, dx = 0.1, K = diag(1e-13, 1e-13, 1e-14), phi = 0.2, no rocks, default Dirichlet S = 1 boundaries,
    v = (1e-6, 0, 0), g = (0, 0, -9.81), S0 = 0, CFL factors (0.3, 2e-4, 5e7), transportSolve(86400 s), defaults):
    CFL times 6000 / 227537 / 500 s, 346 substeps, sum S = 414.11235818963286, S[0], S[999] as below."""
    from opm_porsol_b200 import synth
    from oracle.ref import PortSolver
    g = synth.cartesian_grid(10, 10, 10, 0.1, 0.1, 0.1, unique_bids=False)
    N = g["N"]
    perm = np.zeros((N, 9))
    perm[:, 0] = perm[:, 4] = 1e-13
    perm[:, 8] = 1e-14
    case = synth.make_case("survey-kat", g, poro=np.full(N, 0.2), perm=perm, sat0=np.zeros(N), gravity=[0.0, 0.0, -9.81],
                           hf_flux=synth.constant_velocity_flux(g, (1e-6, 0.0, 0.0)), time=86400.0)
    port = PortSolver(case, cfl_factors=[0.3, 2e-4, 5e7])
    cfl = port.cfl_times()
    assert abs(cfl[0] - 6000.0) < 1e-6 and abs(cfl[2] - 500.0) < 1e-6 and abs(cfl[1] - 227537.0) < 1.0   # as printed in the survey
    out = port.transport_solve(case.sat0, time=86400.0)
    s = out["sat"]
    assert out["nsteps"] == 346 and out["attempts"] == 1
    total = 0.0
    for v in s.tolist():                       # left-to-right sum, as the probe accumulated it
        total += v
    assert total == 414.11235818963286
    assert float(s[0]) == 0.53870412522172473 and float(s[999]) == 0.30118148462420075
