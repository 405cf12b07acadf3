import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "opm-porsol_b200", "python"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracles():
    """The plain-C oracle builds anywhere; the compiled reference only where /root/reference exists
    (elsewhere the prebuilt oracle/_ref/libeuler_ref.so that travelled with the snapshot is used)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port", "ref"], check=True)
    yield


def small_cases():
    """Seeded cases small enough for the CPU oracle to finish in well under a second each."""
    from opm_porsol_b200 import synth
    cases = []
    c = synth.config_c1()                       # BASELINE config 0: 10^3, periodic, V+G
    cases.append(("c1_periodic_vg", c))
    c = synth.config_c2(12)                     # config 1 at reduced size: rotated K, rock table, V+G+C
    cases.append(("c2_aniso_cap", c))
    c = synth.config_c3(16, 16, 12)             # config 2 at reduced size: faulted corner-point, 3 rocks
    cases.append(("c3_faulted_3rocks", c))
    c = synth.config_c4(12, 10, 8)              # config 3 at reduced size: heterogeneous perm, V+G
    cases.append(("c4_hetero_vg", c))
    # march-active shapes (nx*ny a multiple of 32): the FAST kernel runs its z-march with the carried face flux, the
    # code path bench.py times (k_fast_step<ROCKS, !MULTIROCK, !CAP> for V+G, <.., CAP> with the capillary term)
    c = synth.config_c4(64, 16, 6)
    cases.append(("c4_march_vg", c))
    c = synth.config_c4(32, 32, 6, capillary=True)
    cases.append(("c4_march_vgc", c))
    c = synth.random_geometry_case(6, 5, 4, seed=1, n_rocks=2, periodic=(True, False, True))
    cases.append(("rand_periodic_2rocks", c))
    c = synth.random_geometry_case(5, 5, 5, seed=3, n_rocks=0, periodic=(True, True, True))
    cases.append(("rand_allperiodic_norock", c))
    c = synth.random_geometry_case(7, 3, 5, seed=10, n_rocks=3, periodic=(False, False, False), use_j=False)
    cases.append(("rand_dirichlet_3rocks_noJ", c))
    c = synth.random_geometry_case(4, 6, 3, seed=11, n_rocks=1, full_tensor=False)
    c.method_gravity = False
    cases.append(("rand_nogravity", c))
    c = synth.random_geometry_case(5, 4, 4, seed=13, n_rocks=2)
    c.method_viscous = False
    c.clamp_sat = True
    cases.append(("rand_noviscous_clamp", c))
    return cases


def tensor_cases():
    from opm_porsol_b200 import synth
    c = synth.random_geometry_case(5, 4, 3, seed=5, n_rocks=2, periodic=(False, True, False), mobility_kind=1)
    d = synth.random_geometry_case(4, 4, 4, seed=6, n_rocks=0, mobility_kind=1)
    return [("tensor_2rocks_periodic", c), ("tensor_norock", d)]


def tensor_aligned_cases():
    """Diagonal tensor mobility (ReservoirPropertyCapillaryAnisotropicRelperm, the class aniso_simulator_test uses) on
    grids with axis-aligned face normals: FAST mode applies (kernels)."""
    from opm_porsol_b200 import synth
    a = synth.random_geometry_case(6, 5, 4, seed=15, n_rocks=2, periodic=(True, False, False), mobility_kind=1,
                                   perturb_normals=False)
    b = synth.random_geometry_case(5, 7, 3, seed=16, n_rocks=3, mobility_kind=1, perturb_normals=False)
    b.method_gravity = False
    c = synth.random_geometry_case(9, 4, 4, seed=17, n_rocks=1, periodic=(False, True, True), mobility_kind=1,
                                   perturb_normals=False, sources=False)
    c.method_capillary = False
    return [("tensor_aligned_2rocks_periodic", a), ("tensor_aligned_3rocks_nogravity", b), ("tensor_aligned_1rock_nocap", c)]


def active_cfl_dt(case, cfl):
    """courant * min over the CFL terms the solver actually uses (EulerUpstream_impl.hpp:275-318)."""
    v = cfl[0] if (case.method_viscous and case.use_cfl_viscous) else 1e99
    g = cfl[1] if (case.method_gravity and case.use_cfl_gravity) else 1e99
    c = cfl[2] if (case.method_capillary and case.use_cfl_capillary) else 1e99
    return min(v, g, c)*case.courant
