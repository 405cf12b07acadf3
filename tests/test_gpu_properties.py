"""Size-independent properties at sizes the CPU oracle cannot cover in seconds (BASELINE configs 1-2 at or near
full size): FAST vs STRICT agreement on the device, discrete mass balance, independence of the result from how
the grid was chunked on upload, idempotence of re-uploading the same grid (initObj twice)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(case, mode, fac, chunk_cells=1 << 20):
    from opm_porsol_b200 import EulerUpstream
    from opm_porsol_b200.binding import params_from_case
    d = EulerUpstream(device=0, mode=mode)
    d.init(params_from_case(case))
    d.initObj(case, cfl_factors=fac, chunk_cells=chunk_cells)
    return d


def _factors(case):
    from opm_porsol_b200.binding import make_fluid
    fluid, _ = make_fluid(case)
    return np.array(fluid.cfl_factor[:])


def test_fast_vs_strict_c2_full_size():
    """BASELINE config 1: 100^3 Cartesian (1M cells), rotated anisotropic K, rock table, capillary on."""
    from opm_porsol_b200 import synth
    case = synth.config_c2(100)
    fac = _factors(case)
    fast, strict = _dev(case, "fast", fac), _dev(case, "strict", fac)
    fast.upload_state(case.sat0, case.hf_flux)
    strict.upload_state(case.sat0, case.hf_flux)
    cfl = strict.cfl_times(case.gravity)
    assert np.array_equal(cfl, fast.cfl_times(case.gravity))
    dt = 0.25*min(cfl)*case.courant
    for q in range(3):
        a = strict.small_step(dt, case.gravity)
        b = fast.small_step(dt, case.gravity)
        sa, sb = strict.download_saturation(), fast.download_saturation()
        assert a["status"] == 0 and b["status"] == 0
        assert np.abs(sa - sb).max() <= 1e-12
        assert np.abs(a["residual"] - b["residual"]).max() <= 1e-12*np.abs(a["residual"]).max()
        fast.upload_saturation(sa)
    # a full transportSolve: identical step counts, <= 1e-9
    s1, s2 = case.sat0.copy(), case.sat0.copy()
    r1 = strict.transportSolve(s1, 30.3*dt, case.gravity, case.hf_flux)
    r2 = fast.transportSolve(s2, 30.3*dt, case.gravity, case.hf_flux)
    assert r1.nsteps == r2.nsteps and r1.attempts == r2.attempts == 1
    assert np.abs(s1 - s2).max() <= 1e-9
    fast.close()
    strict.close()


def test_mass_balance_periodic_closed_system():
    """All-periodic grid without sources: sum(porevol * S) is conserved by every substep (each face flux is
    subtracted from one cell and added to the other)."""
    from opm_porsol_b200 import synth
    g = synth.cartesian_grid(48, 40, 32, 1.0, 1.0, 0.5, unique_bids=True, periodic=(True, True, True))
    N = g["N"]
    case = synth.make_case("closed", g, poro=0.1 + 0.2*synth.mt_uniform(5, N), perm=synth.lognormal_perm(N, 6),
                           rock_id=np.zeros(N, dtype=np.int32), rocks=[synth.corey_table()],
                           sat0=0.2 + 0.5*synth.mt_uniform(7, N), gravity=[0.0, 0.0, -9.80665],
                           hf_flux=synth.constant_velocity_flux(g, (1e-6, 5e-7, 2.5e-7)))
    fac = _factors(case)
    pv = case.cell_volume*case.poro
    for mode in ("strict", "fast"):
        dev = _dev(case, mode, fac)
        dev.upload_state(case.sat0, case.hf_flux)
        dt = 0.4*min(dev.cfl_times(case.gravity))*case.courant
        m0 = float((pv*case.sat0).sum())
        sat = case.sat0.copy()
        rep = dev.transportSolve(sat, 25*dt, case.gravity, case.hf_flux)
        assert rep.status == 0
        m1 = float((pv*sat).sum())
        assert abs(m1 - m0) <= 1e-10*abs(m0), (mode, m0, m1)
        assert np.abs(sat - case.sat0).max() > 1e-3          # something did move
        dev.close()


def test_chunking_and_reupload_do_not_change_results():
    from opm_porsol_b200 import synth
    case = synth.config_c3(48, 40, 24)               # faulted corner-point, 3 rocks, V+G+C
    fac = _factors(case)
    outs = []
    for chunk in (1 << 20, 4097, 777):
        dev = _dev(case, "fast", fac, chunk_cells=chunk)
        if chunk == 777:
            dev.initObj(case, cfl_factors=fac, chunk_cells=chunk)      # initObj twice (BCs / props may change in between)
        dev.upload_state(case.sat0, case.hf_flux)
        dt = 0.3*min(dev.cfl_times(case.gravity))*case.courant
        sat = case.sat0.copy()
        dev.transportSolve(sat, 12*dt, case.gravity, case.hf_flux)
        outs.append(sat)
        assert 0.0 <= dev.regular_fraction() < 1.0           # fault planes every 12 cells: (almost) every slot is irregular
        dev.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
