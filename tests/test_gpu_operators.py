"""GPU parity of the operator-style entry points next to transportSolve:
eu_compute_residual (EulerUpstreamResidual::computeResidual, second caller ImplicitCapillarity_impl.hpp:178-180),
eu_compute_cap_pressures (computeCapPressures / computeCapPressure) and the post-transport diagnostics of
SimulatorUtilities.hpp (estimateCellVelocity, computePhaseVelocities, fractionalFlow), all through the C ABI,
against the CPU oracle on identical seeded inputs.  Bit-exact where the device follows the reference's operation
order (STRICT residual, every diagnostic in every mode); FAST residual within 1e-12 of the residual scale."""
import numpy as np
import pytest

from conftest import small_cases, tensor_aligned_cases, tensor_cases

pytestmark = pytest.mark.gpu

METHOD_SETS = [(True, True, False), (True, False, False), (False, True, True), (True, True, True)]
ALL = small_cases() + tensor_cases() + tensor_aligned_cases()


def _solvers(case, mode):
    from opm_porsol_b200 import EulerUpstream
    from opm_porsol_b200.binding import params_from_case
    from oracle.ref import PortSolver, RefSolver, ref_available
    if case.mobility_kind == 1:
        if not ref_available():
            pytest.skip("tensor-mobility CFL factors come from the compiled reference")
        fac = RefSolver(case).cfl_factors()
        port = PortSolver(case, cfl_factors=fac)
    else:
        port = PortSolver(case)
        fac = port.compute_cfl_factors()
    dev = EulerUpstream(device=0, mode=mode)
    dev.init(params_from_case(case))
    dev.initObj(case, cfl_factors=fac)
    return dev, port


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("name,case", ALL, ids=[n for n, _ in ALL])
def test_compute_residual_operator(name, case, mode):
    if case.mobility_kind == 1 and mode == "fast":
        mode = "auto"                                   # tensor mobility: FAST only on axis-aligned grids
    dev, port = _solvers(case, mode)
    inj = (case.src_cell, case.src_rate)
    exact = mode == "strict"
    rng = np.random.default_rng(5)
    sat_b = np.clip(case.sat0 + 0.05*rng.standard_normal(case.N), 0.02, 0.98)
    dev.upload_state(case.sat0, case.hf_flux)           # a resident state the operator must leave alone
    for k, m in enumerate(METHOD_SETS):
        sat = case.sat0 if k % 2 == 0 else sat_b
        want = port.compute_residual(sat, m)
        got = dev.computeResidual(sat, case.gravity, case.hf_flux if k == 0 else None, inj, *m)
        if exact:
            assert np.array_equal(got, want), (m, np.abs(got - want).max())
        else:
            assert np.abs(got - want).max() <= 1e-12*(np.abs(want).max() + 1e-300), m
    assert np.array_equal(dev.download_saturation(), case.sat0)
    # the solver's own parameters are untouched: a substep afterwards is the oracle's substep
    a = port.small_step(case.sat0, 1.0)
    b = dev.small_step(1.0, case.gravity, inj)
    if exact:
        assert np.array_equal(b["residual"], a["residual"])
    else:
        assert np.abs(b["residual"] - a["residual"]).max() <= 1e-12*(np.abs(a["residual"]).max() + 1e-300)
    dev.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("name,case", ALL, ids=[n for n, _ in ALL])
def test_diagnostics_bit_exact(name, case, mode):
    if case.mobility_kind == 1 and mode == "fast":
        pytest.skip("tensor mobility has one mode")
    dev, port = _solvers(case, mode if case.mobility_kind == 0 else "auto")
    dev.upload_state(case.sat0, case.hf_flux)
    cv = dev.cellVelocity()
    assert np.array_equal(cv, port.cell_velocity())
    vw, vo = dev.phaseVelocities()                       # resident saturation and fluxes
    vw_p, vo_p = port.phase_velocities(case.sat0, cv)
    assert np.array_equal(vw, vw_p) and np.array_equal(vo, vo_p)
    rng = np.random.default_rng(8)
    sat_b = np.clip(case.sat0 + 0.05*rng.standard_normal(case.N), 0.0, 1.0)
    vw, vo = dev.phaseVelocities(sat_b, cv)              # host saturation and velocity field
    vw_p, vo_p = port.phase_velocities(sat_b, cv)
    assert np.array_equal(vw, vw_p) and np.array_equal(vo, vo_p)
    assert np.array_equal(dev.fractionalFlow(), port.frac_flows(case.sat0))
    assert np.array_equal(dev.fractionalFlow(sat_b), port.frac_flows(sat_b))
    assert np.array_equal(dev.computeCapPressures(sat_b), port.cap_pressures(sat_b))
    assert np.array_equal(dev.download_saturation(), case.sat0)
    dev.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_resident_state_survives_flux_only_upload_after_odd_substeps(mode):
    """IMPES loop with the state resident: transportSolve (odd number of substeps, so the state ends in the second
    ping-pong buffer), new fluxes with eu_upload_state(h, NULL, flux), transportSolve again.  Must equal two
    consecutive transportSolves of the oracle."""
    from opm_porsol_b200 import EulerUpstream, synth
    from opm_porsol_b200.binding import params_from_case
    from oracle.ref import PortSolver
    case = synth.config_c4(32, 32, 6, capillary=True)
    case.min_steps = case.max_steps = 3
    port = PortSolver(case)
    fac = port.compute_cfl_factors()
    time = 1.2*min(port.cfl_times())*case.courant
    flux2 = 0.5*case.hf_flux
    a = port.transport_solve(case.sat0, time=time)
    b = port.transport_solve(a["sat"], time=time, hf_flux=flux2)
    assert a["nsteps"] == b["nsteps"] == 3
    dev = EulerUpstream(device=0, mode=mode)
    dev.init(params_from_case(case))
    dev.initObj(case, cfl_factors=fac)
    dev.upload_state(case.sat0, case.hf_flux)
    r1 = dev.transportSolveResident(time, case.gravity)
    dev.upload_state(None, flux2)
    r2 = dev.transportSolveResident(time, case.gravity)
    assert r1.nsteps == r2.nsteps == 3
    sat = dev.download_saturation()
    if mode == "strict":
        assert np.array_equal(sat, b["sat"])
    else:
        assert np.abs(sat - b["sat"]).max() <= 1e-9
    assert np.abs(sat - a["sat"]).max() > 1e-6        # the second solve did move the state
    dev.close()
