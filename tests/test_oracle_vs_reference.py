"""Pin the plain-C oracle (oracle/euler_oracle.c) against the UNMODIFIED reference headers compiled in
oracle/_ref (EulerUpstream_impl.hpp, EulerUpstreamResidual_impl.hpp, CflCalculator.hpp,
ReservoirPropertyCapillary*_impl.hpp, RockJfunc.hpp ...) on identical seeded inputs: bit for bit.
Skipped where neither /root/reference nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

from conftest import active_cfl_dt, small_cases, tensor_cases
from oracle.ref import PortSolver, RefSolver, ref_available

needs_ref = pytest.mark.skipif(not ref_available(), reason="compiled reference (oracle/_ref) not available")
ALL = small_cases() + tensor_cases()


@needs_ref
@pytest.mark.parametrize("name,case", ALL, ids=[n for n, _ in ALL])
def test_port_equals_reference(name, case):
    ref = RefSolver(case)
    fac = ref.cfl_factors()
    if case.mobility_kind == 0:
        port = PortSolver(case)
        assert np.array_equal(port.compute_cfl_factors(), fac)       # computeCflFactors restatement
    port = PortSolver(case, cfl_factors=fac)
    cfl_ref, total_ref = ref.cfl_times()
    assert np.array_equal(port.cfl_times(), cfl_ref)
    assert total_ref == active_cfl_dt(case, cfl_ref)
    dt = 0.5*total_ref if case.mobility_kind == 0 else 50.0
    s1, s2 = case.sat0.copy(), case.sat0.copy()
    for _ in range(3):
        a, b = ref.small_step(s1, dt), port.small_step(s2, dt)
        s1, s2 = a["sat"], b["sat"]
        assert a["status"] == b["status"] == 0
        assert np.array_equal(a["residual"], b["residual"])
        assert np.array_equal(s1, s2)
    if case.mobility_kind == 0:
        time = 17.3*total_ref
        a, b = ref.transport_solve(case.sat0, time=time), port.transport_solve(case.sat0, time=time)
        assert (a["status"], a["nsteps"], a["attempts"]) == (b["status"], b["nsteps"], b["attempts"]) == (0, 18, 1)
        assert np.array_equal(a["sat"], b["sat"])


@needs_ref
def test_pointwise_property_functions():
    """phaseMobility / capillaryPressure / fractionalFlow at saturations inside, on the nodes of and
    outside the rock tables (the third-party NonuniformTableLinear contract)."""
    for name, case in ALL:
        ref, port = RefSolver(case), PortSolver(case, cfl_factors=np.ones(3))
        sats = [-0.0005, 0.0, 0.03, 0.1, 0.2137, 0.5, 0.77, 0.9, 0.95, 1.0, 1.0007]
        if case.rocks:
            sats += list(case.rocks[0].s[:4]) + list(case.rocks[-1].s[-3:])
        for cell in (0, case.N//2, case.N - 1):
            for s in sats:
                for ph in (0, 1):
                    assert np.array_equal(ref.mobility(ph, cell, s), port.mobility(ph, cell, s)), (name, cell, s, ph)
                assert ref.cap_pressure(cell, s) == port.cap_pressure(cell, s)
                a, b = ref.frac_flow(cell, s), port.frac_flow(cell, s)
                assert a == b or (np.isnan(a) and np.isnan(b))


@needs_ref
def test_retry_loop_and_failure_message():
    from opm_porsol_b200 import synth
    case = synth.random_geometry_case(5, 4, 4, seed=21, n_rocks=1, sources=False)
    case.max_steps = 2
    ref, port = RefSolver(case), PortSolver(case)
    total = active_cfl_dt(case, port.cfl_times())
    a, b = ref.transport_solve(case.sat0, time=40.0*total), port.transport_solve(case.sat0, time=40.0*total)
    assert a["attempts"] > 1
    assert (a["status"], a["nsteps"], a["attempts"]) == (b["status"], b["nsteps"], b["attempts"])
    assert np.array_equal(a["sat"], b["sat"])
    # a case that cannot succeed: rethrown after 10 retries with the reference's message
    case = synth.random_geometry_case(7, 3, 5, seed=9, n_rocks=3, use_j=False)
    case.max_steps = 2
    ref, port = RefSolver(case), PortSolver(case)
    total = active_cfl_dt(case, port.cfl_times())
    a, b = ref.transport_solve(case.sat0, time=17.3*total), port.transport_solve(case.sat0, time=17.3*total)
    assert a["status"] == b["status"] == 1 and a["attempts"] == b["attempts"] == 11 and a["nsteps"] == b["nsteps"]
    assert a["error"].startswith("Saturation out of range in EulerUpstream: Cell %d " % b["bad_cell"])


@needs_ref
def test_zero_flux_cfl_throws_like_reference():
    """findCFLtimeVelocity throws when a cell gives dt == 0 (zero pore volume), CflCalculator.hpp:75-77."""
    from opm_porsol_b200 import synth
    case = synth.config_c4(4, 4, 3)
    case.poro[5] = 0.0
    ref, port = RefSolver(case), PortSolver(case)
    a, b = ref.transport_solve(case.sat0, time=10.0), port.transport_solve(case.sat0, time=10.0)
    assert a["status"] == 1 and "Cfl computation gave dt = 0.0" in a["error"]
    assert b["status"] == 2


METHOD_SETS = [(True, True, False), (True, False, False), (False, True, True), (True, True, True)]


@needs_ref
@pytest.mark.parametrize("name,case", ALL, ids=[n for n, _ in ALL])
def test_residual_operator_and_diagnostics_equal_reference(name, case):
    """EulerUpstreamResidual::computeResidual with explicit method flags (the ImplicitCapillarity call passes
    capillary = false, ImplicitCapillarity_impl.hpp:178-180) and the post-transport diagnostics of
    SimulatorUtilities.hpp:59-86,153-170,219-230: the C port against the compiled reference, bit for bit."""
    ref = RefSolver(case)
    port = PortSolver(case, cfl_factors=ref.cfl_factors())
    for m in METHOD_SETS:
        a = ref.compute_residual(case.sat0, m)
        b = port.compute_residual(case.sat0, m)
        assert np.array_equal(a, b), (m, np.abs(a - b).max())
    # the flags really are arguments: the solver's own parameters are untouched afterwards
    s1, s2 = ref.small_step(case.sat0, 1.0), port.small_step(case.sat0, 1.0)
    assert np.array_equal(s1["residual"], s2["residual"])
    cv_ref, cv_port = ref.cell_velocity(), port.cell_velocity()
    assert np.array_equal(cv_ref, cv_port)
    vw_r, vo_r = ref.phase_velocities(case.sat0, cv_ref)
    vw_p, vo_p = port.phase_velocities(case.sat0, cv_port)
    assert np.array_equal(vw_r, vw_p) and np.array_equal(vo_r, vo_p)
    assert np.array_equal(ref.cap_pressures(case.sat0), port.cap_pressures(case.sat0))
