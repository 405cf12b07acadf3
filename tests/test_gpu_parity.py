"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

STRICT mode must be bit-identical to the oracle (which is itself bit-identical to the reference
headers, tests/test_oracle_vs_reference.py).  FAST mode must agree within the tolerance BASELINE.json
states: max|dS| <= 1e-12 per substep, <= 1e-9 after a full transportSolve, identical step counts.
"""
import numpy as np
import pytest

from conftest import active_cfl_dt, small_cases, tensor_aligned_cases, tensor_cases

pytestmark = pytest.mark.gpu

TOL_SUBSTEP = 1e-12
TOL_SOLVE = 1e-9


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
@pytest.mark.parametrize("name,case", small_cases(), ids=[n for n, _ in small_cases()])
def test_cfl_and_substeps(name, case, mode):
    dev, port = _solvers(case, mode)
    dev.upload_state(case.sat0, case.hf_flux)
    cfl_dev = dev.cfl_times(case.gravity)
    cfl_cpu = port.cfl_times()
    assert np.array_equal(cfl_dev, cfl_cpu), (cfl_dev, cfl_cpu)          # bit-exact in both modes
    total = active_cfl_dt(case, cfl_cpu)
    dt = 0.5*total
    s_cpu = case.sat0.copy()
    inj = (case.src_cell, case.src_rate)
    for q in range(4):
        a = port.small_step(s_cpu, dt)
        b = dev.small_step(dt, case.gravity, inj)
        s_cpu = a["sat"]
        s_dev = dev.download_saturation()
        assert a["status"] == 0 and b["status"] == 0
        if mode == "strict":
            assert np.array_equal(b["residual"], a["residual"]), np.abs(b["residual"] - a["residual"]).max()
            assert np.array_equal(s_dev, s_cpu)
        else:
            assert np.abs(s_dev - s_cpu).max() <= TOL_SUBSTEP
            scale = np.abs(a["residual"]).max() + 1e-300
            assert np.abs(b["residual"] - a["residual"]).max() <= 1e-12*scale
        # keep both paths on the oracle's trajectory so every substep is an independent comparison
        dev.upload_saturation(s_cpu)
    dev.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("name,case", small_cases(), ids=[n for n, _ in small_cases()])
def test_transport_solve(name, case, mode):
    dev, port = _solvers(case, mode)
    total = active_cfl_dt(case, port.cfl_times())
    time = 17.3*total                                  # 18 substeps at the natural CFL
    a = port.transport_solve(case.sat0, time=time)
    sat = case.sat0.copy()
    rep = dev.transportSolve(sat, time, case.gravity, case.hf_flux, (case.src_cell, case.src_rate))
    assert a["status"] == 0 and rep.status == 0
    assert rep.nsteps == a["nsteps"] == 18
    assert rep.attempts == a["attempts"]
    assert np.array_equal(np.array(rep.cfl_dt), a["cfl_dt"]), (rep.cfl_dt, a["cfl_dt"])
    if mode == "strict":
        assert np.array_equal(sat, a["sat"])
    else:
        assert np.abs(sat - a["sat"]).max() <= TOL_SOLVE
    dev.close()


TENSOR_ALL = tensor_cases() + tensor_aligned_cases()


@pytest.mark.parametrize("name,case", TENSOR_ALL, ids=[n for n, _ in TENSOR_ALL])
def test_tensor_mobility_strict(name, case):
    dev, port = _solvers(case, "strict")
    dev.upload_state(case.sat0, case.hf_flux)
    s_cpu = case.sat0.copy()
    for q in range(3):
        a = port.small_step(s_cpu, 50.0)
        b = dev.small_step(50.0, case.gravity, (case.src_cell, case.src_rate))
        s_cpu = a["sat"]
        assert np.array_equal(b["residual"], a["residual"])
        assert np.array_equal(dev.download_saturation(), s_cpu)
    dev.close()


@pytest.mark.parametrize("mode", ["fast", "auto"])
@pytest.mark.parametrize("name,case", TENSOR_ALL, ids=[n for n, _ in TENSOR_ALL])
def test_tensor_mobility_fast(name, case, mode):
    """FAST mode for the diagonal tensor mobility, on axis-aligned grids (scalar formula with the face axis' mobility
    component) and on oblique normals (three-component kernel): per-substep and per-transportSolve gates, bit-exact
    CFL times and identical step counts, like the scalar class."""
    import copy
    case = copy.copy(case)
    # The capillary CFL factor of the reference's tensor class is not a usable number (~1e-89 s steps, see
    # tests/golden/make_golden.py): fixed 50 s substeps, 18 of them per transportSolve, exercise the update instead.
    case.min_steps = case.max_steps = 18
    dev, port = _solvers(case, mode)
    assert dev.resolved_mode() == "fast"
    dev.upload_state(case.sat0, case.hf_flux)
    assert np.array_equal(dev.cfl_times(case.gravity), port.cfl_times())
    inj = (case.src_cell, case.src_rate)
    s_cpu = case.sat0.copy()
    moved = 0.0
    for q in range(3):
        a = port.small_step(s_cpu, 50.0)
        b = dev.small_step(50.0, case.gravity, inj)
        moved = max(moved, float(np.abs(a["sat"] - s_cpu).max()))
        s_cpu = a["sat"]
        assert a["status"] == 0 and b["status"] == 0
        assert np.abs(dev.download_saturation() - s_cpu).max() <= TOL_SUBSTEP
        assert np.abs(b["residual"] - a["residual"]).max() <= 1e-12*(np.abs(a["residual"]).max() + 1e-300)
        dev.upload_saturation(s_cpu)
    assert moved > 1e-6, "the substeps are meant to change the saturation"
    a = port.transport_solve(case.sat0, time=900.0)
    sat = case.sat0.copy()
    rep = dev.transportSolve(sat, 900.0, case.gravity, case.hf_flux, inj, raise_on_error=False)
    assert (rep.status != 0) == (a["status"] != 0) and rep.nsteps == a["nsteps"] == 18 and rep.attempts == a["attempts"]
    if a["status"] == 0:
        assert np.abs(sat - a["sat"]).max() <= TOL_SOLVE
    dev.close()


def test_retry_doubling_matches():
    """An unstable step count forces the throw-and-retry loop (EulerUpstream_impl.hpp:187-213)."""
    from opm_porsol_b200 import synth
    case = synth.random_geometry_case(5, 4, 4, seed=21, n_rocks=1, sources=False)
    case.max_steps = 2
    for mode in ("strict", "fast"):
        dev, port = _solvers(case, mode)
        total = active_cfl_dt(case, port.cfl_times())
        time = 40.0*total
        a = port.transport_solve(case.sat0, time=time)
        sat = case.sat0.copy()
        rep = dev.transportSolve(sat, time, case.gravity, case.hf_flux, None, raise_on_error=False)
        assert a["attempts"] > 1, "case is meant to need retries"
        assert (rep.status != 0) == (a["status"] != 0)
        assert rep.attempts == a["attempts"] and rep.nsteps == a["nsteps"]
        if a["status"] == 0:
            assert np.abs(sat - a["sat"]).max() <= (0.0 if mode == "strict" else TOL_SOLVE)
        else:
            assert rep.bad_cell == a["bad_cell"]
        dev.close()


def test_failure_after_ten_retries_reports_reference_message():
    """Sources push a cell past 1.001 whatever the step count: the reference rethrows after 10
    retries with 'Saturation out of range in EulerUpstream: Cell <c>   sat <s>'."""
    from opm_porsol_b200 import EulerB200Error, synth
    case = synth.random_geometry_case(7, 3, 5, seed=9, n_rocks=3, use_j=False)
    case.max_steps = 2
    for mode in ("strict", "fast"):
        dev, port = _solvers(case, mode)
        total = active_cfl_dt(case, port.cfl_times())
        a = port.transport_solve(case.sat0, time=17.3*total)
        assert a["status"] == 1 and a["attempts"] == 11
        sat = case.sat0.copy()
        with pytest.raises(EulerB200Error) as ei:
            dev.transportSolve(sat, 17.3*total, case.gravity, case.hf_flux, (case.src_cell, case.src_rate))
        rep = dev.last_report
        assert ei.value.code == 3 and "Saturation out of range in EulerUpstream: Cell %d " % a["bad_cell"] in str(ei.value)
        assert rep.attempts == 11 and rep.nsteps == a["nsteps"] and rep.bad_cell == a["bad_cell"]
        assert abs(rep.bad_value - a["bad_value"]) <= (0.0 if mode == "strict" else 1e-9)
        assert rep.substeps_executed >= a["substeps_executed"]
        dev.close()


@pytest.mark.parametrize("name,case", small_cases(), ids=[n for n, _ in small_cases()])
def test_slice_class_kernel_still_matches(name, case, monkeypatch):
    """EU_BOX=0: the slice-class kernel (warp marches, generic gather) that the box kernel replaced on box-numbered grids
    and that still serves every other grid: the same per-transportSolve gate."""
    monkeypatch.setenv("EU_BOX", "0")
    dev, port = _solvers(case, "fast")
    total = active_cfl_dt(case, port.cfl_times())
    time = 17.3*total
    a = port.transport_solve(case.sat0, time=time)
    sat = case.sat0.copy()
    rep = dev.transportSolve(sat, time, case.gravity, case.hf_flux, (case.src_cell, case.src_rate))
    assert rep.nsteps == a["nsteps"] == 18 and rep.attempts == a["attempts"]
    assert np.abs(sat - a["sat"]).max() <= TOL_SOLVE
    assert dev.work_plan()["kernel"] == "slice-class"
    dev.close()
