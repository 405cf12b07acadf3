"""The C-ABI library loads and exports every symbol include/euler_b200.h declares; host-side logic that
needs no GPU: CFL-factor helper, boundary resolution, synthetic generators, loud failure without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, small_cases


def test_library_exports_every_declared_symbol():
    import opm_porsol_b200 as eub
    lib = eub.load_library()
    hdr = open(os.path.join(ROOT, "include", "euler_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(eu_[a-z_0-9]+)\s*\(", hdr)))
    names = [n for n in names if n != "eu_allreduce_fn"]
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/euler_b200.h but not exported"
    assert lib.eu_abi_version() == 1


def test_no_cpu_fallback():
    """Without a device eu_create must fail loudly (skipped where a GPU is present)."""
    import opm_porsol_b200 as eub
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    with pytest.raises(eub.EulerB200Error) as ei:
        eub.EulerUpstream(device=0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opm-porsol_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle shim", "").replace("/ the oracle", "").lower() or \
                    all("import" not in l and "#include" not in l for l in txt.splitlines() if "oracle" in l.lower()), (dp, f)


@pytest.mark.parametrize("name,case", small_cases(), ids=[n for n, _ in small_cases()])
def test_cfl_factor_helper_matches_oracle(name, case):
    """eu_compute_cfl_factors (product, host) == oracle restatement of computeCflFactors, bit for bit."""
    from opm_porsol_b200.binding import make_fluid
    from oracle.ref import PortSolver
    fluid, _ = make_fluid(case)
    assert np.array_equal(np.array(fluid.cfl_factor[:]), PortSolver(case).compute_cfl_factors())


def test_boundary_resolution_periodic_pairs():
    from opm_porsol_b200 import synth
    from opm_porsol_b200.binding import resolve_boundary
    g = synth.cartesian_grid(4, 3, 2, unique_bids=True, periodic=(True, False, True))
    N = g["N"]
    case = synth.make_case("p", g, poro=np.full(N, 0.2), perm=np.tile(np.eye(3).reshape(9), (N, 1)), sat0=np.zeros(N),
                           gravity=[0, 0, -9.8], hf_flux=np.zeros(6*N))
    bnd_hf, kind, sat, pcell, pface = resolve_boundary(case)
    cell = bnd_hf//6
    face = bnd_hf % 6
    i, j, k = cell % 4, (cell//4) % 3, cell//12
    for n in range(bnd_hf.shape[0]):
        if face[n] in (0, 1):          # x periodic
            assert kind[n] == 2 and pface[n] == (1 - face[n])
            assert pcell[n] == (3 - i[n]) + 4*(j[n] + 3*k[n])
        elif face[n] in (4, 5):        # z periodic
            assert kind[n] == 2 and pface[n] == 9 - face[n]
            assert pcell[n] == i[n] + 4*(j[n] + 3*(1 - k[n]))
        else:
            assert kind[n] == 1 and pcell[n] == -1 and sat[n] == 1.0


def test_faulted_grid_is_symmetric_and_conservative():
    """Every interior half-face of the faulted corner-point generator has exactly one twin with the same
    area and opposite normal; lateral areas of a cell side sum to the full side."""
    from opm_porsol_b200 import synth
    g = synth.faulted_grid(8, 6, 5, 10.0, 10.0, 1.0, faults_i=[(3, 1.5), (6, 0.75)], faults_j=[(2, 2.25)])
    off, nbr = g["hf_offset"], g["hf_nbr"]
    N = g["N"]
    counts = np.diff(off)
    assert counts.max() >= 8 and counts.min() >= 6          # split faces on the fault planes
    cell_of = np.repeat(np.arange(N), counts)
    pairs = {}
    for h in range(nbr.shape[0]):
        if nbr[h] >= 0:
            pairs.setdefault((cell_of[h], nbr[h]), []).append(h)
    for (a, b), hs in pairs.items():
        assert len(hs) == 1
        t = pairs[(b, a)][0]
        assert abs(g["hf_area"][hs[0]] - g["hf_area"][t]) < 1e-12
        assert np.array_equal(g["hf_normal"][hs[0]], -g["hf_normal"][t])
        assert np.abs(g["hf_centroid"][hs[0]] - g["hf_centroid"][t]).max() < 1e-9
    # lateral side areas: x sides sum to dy*dz = 10, y sides to dx*dz = 10
    for c in (0, N//2, N - 1):
        hs = np.arange(off[c], off[c + 1])
        for axis, sgn in ((0, -1), (0, 1), (1, -1), (1, 1)):
            sel = hs[g["hf_normal"][hs, axis] == sgn]
            assert abs(g["hf_area"][sel].sum() - 10.0) < 1e-12


def test_table_interpolation_contract():
    """oracle eo_table_eval == the NonuniformTableLinear shim's published algorithm (slope form, binary
    search, extrapolation outside the table), evaluated here in numpy."""
    from oracle.ref import PORT_LIB
    lib = ctypes.CDLL(PORT_LIB)
    lib.eo_table_eval.restype = ctypes.c_double
    x = np.array([0.1, 0.15, 0.3, 0.55, 0.56, 0.9])
    y = np.array([0.0, 0.02, 0.2, 0.5, 0.7, 1.0])
    dp = ctypes.POINTER(ctypes.c_double)
    for s in [-0.2, 0.0, 0.1, 0.12, 0.15, 0.29999, 0.3, 0.555, 0.56, 0.8, 0.9, 0.95, 1.3]:
        j = int(np.clip(np.searchsorted(x, s, side="right") - 1, 0, len(x) - 2))
        want = (y[j + 1] - y[j])/(x[j + 1] - x[j])*(s - x[j]) + y[j]
        got = lib.eo_table_eval(ctypes.c_int(len(x)), x.ctypes.data_as(dp), y.ctypes.data_as(dp), ctypes.c_double(s))
        assert got == want, (s, got, want)
