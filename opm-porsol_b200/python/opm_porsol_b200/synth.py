"""Synthetic transport cases in the flat half-face form the C-ABI consumes.

The reference builds its inputs on the host (setupGridAndProps.hpp:118-136 for
``fileformat=cartesian``, corner-point decks through dune-cornerpoint, boundary
conditions through setupBoundaryConditions.hpp:55-63,136-157, fluxes from the mimetic
pressure solver, attic/euler/EulerSolverTester.hpp:73-91 for a constant-velocity flux).
None of that is on the hot path; this module fabricates equivalent inputs directly in
the flat layout (include/euler_b200.h) so that tests, the oracle and bench.py all see
identical bits.  Everything is float64 / int32 numpy.

Half-face order inside a cell is the "reference face order" the kernels must respect:
for hexahedral cells x-,x+,y-,y+,z-,z+, split faces on a fault plane in ascending z.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

DARCY = 9.869232667160130e-13
MILLIDARCY = 1e-3 * DARCY

# defaults of ReservoirPropertyCommon (ReservoirPropertyCommon_impl.hpp:259-262)
DEFAULT_DENS = (1013.9, 834.7)
DEFAULT_VISC = (1.0e-3, 3.0e-3)


@dataclass
class RockTable:
    """One rock in 'Statoil format' (RockJfunc.hpp:162-218): columns S, krw, kro, J.
    For the anisotropic-relperm variant (RockAnisotropicRelperm.hpp:109-152) the
    diagonal relperm columns per phase and the pc column are used instead."""
    s: np.ndarray
    krw: Optional[np.ndarray] = None
    kro: Optional[np.ndarray] = None
    J: Optional[np.ndarray] = None
    # anisotropic variant
    pc: Optional[np.ndarray] = None
    kr_w: Optional[np.ndarray] = None   # (npts, 3)
    kr_o: Optional[np.ndarray] = None   # (npts, 3)


@dataclass
class Case:
    name: str
    # grid
    N: int
    hf_offset: np.ndarray
    hf_nbr: np.ndarray
    hf_bid: np.ndarray
    hf_area: np.ndarray
    hf_normal: np.ndarray
    hf_centroid: np.ndarray
    cell_volume: np.ndarray
    cell_centroid: np.ndarray
    # properties
    poro: np.ndarray
    perm: np.ndarray                      # (N, 9) row-major, symmetric
    rock_id: Optional[np.ndarray] = None  # int32 or None (no-rock fallback curves)
    rocks: List[RockTable] = field(default_factory=list)
    use_j: bool = True
    sigma: float = 1.0
    theta: float = 0.0
    visc: tuple = DEFAULT_VISC
    dens: tuple = DEFAULT_DENS
    mobility_kind: int = 0                # 0 scalar, 1 diagonal tensor
    # boundary conditions, indexed by boundary id (0 = interior, unused)
    bid_kind: np.ndarray = None           # 0 Dirichlet, 1 periodic
    bid_sat: np.ndarray = None
    bid_partner: np.ndarray = None
    # solver parameters (EulerUpstream_impl.hpp:59-73)
    courant: float = 0.5
    method_viscous: bool = True
    method_gravity: bool = True
    method_capillary: bool = True
    use_cfl_viscous: bool = True
    use_cfl_gravity: bool = True
    use_cfl_capillary: bool = True
    min_steps: int = 1
    max_steps: int = 10000
    check_sat: bool = True
    clamp_sat: bool = False
    # state / per-call inputs
    sat0: np.ndarray = None
    gravity: np.ndarray = None
    hf_flux: np.ndarray = None
    src_cell: np.ndarray = None
    src_rate: np.ndarray = None
    time: float = 86400.0
    dims: tuple = None

    @property
    def H(self) -> int:
        return int(self.hf_offset[-1])


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------------
# Cartesian grids
# ----------------------------------------------------------------------------------
def _cartesian_slab_arrays(nx, ny, nz, dx, dy, dz, k0, k1):
    """Half-face arrays of the cells with k in [k0, k1) of an nx*ny*nz Cartesian grid."""
    n = nx*ny*(k1 - k0)
    c = np.arange(nx*ny*k0, nx*ny*k1, dtype=np.int64)
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx*ny)
    nbr = np.empty((n, 6), dtype=np.int64)
    nbr[:, 0] = np.where(i > 0, c - 1, -1)
    nbr[:, 1] = np.where(i < nx - 1, c + 1, -1)
    nbr[:, 2] = np.where(j > 0, c - nx, -1)
    nbr[:, 3] = np.where(j < ny - 1, c + nx, -1)
    nbr[:, 4] = np.where(k > 0, c - nx*ny, -1)
    nbr[:, 5] = np.where(k < nz - 1, c + nx*ny, -1)
    area = np.empty((n, 6))
    area[:, 0:2] = dy*dz
    area[:, 2:4] = dx*dz
    area[:, 4:6] = dx*dy
    normal = np.zeros((n, 6, 3))
    normal[:, 0, 0] = -1.0
    normal[:, 1, 0] = 1.0
    normal[:, 2, 1] = -1.0
    normal[:, 3, 1] = 1.0
    normal[:, 4, 2] = -1.0
    normal[:, 5, 2] = 1.0
    cc = np.stack([(i + 0.5)*dx, (j + 0.5)*dy, (k + 0.5)*dz], axis=1)
    fc = np.repeat(cc[:, None, :], 6, axis=1)
    fc[:, 0, 0] = i*dx
    fc[:, 1, 0] = (i + 1)*dx
    fc[:, 2, 1] = j*dy
    fc[:, 3, 1] = (j + 1)*dy
    fc[:, 4, 2] = k*dz
    fc[:, 5, 2] = (k + 1)*dz
    return c, nbr, area, normal, fc, cc


def cartesian_grid(nx, ny, nz, dx=1.0, dy=1.0, dz=1.0, unique_bids=True, periodic=(False, False, False)):
    """Flat arrays of an nx*ny*nz Cartesian grid, cell index c = i + nx*(j + ny*k).

    Boundary ids follow CpGrid: 1..6 for the sides when not unique, else 1..#boundary faces in
    half-face order (setupBoundaryConditions.hpp:48-49).  ``periodic`` marks side pairs whose
    boundary ids are periodic partners (only with unique ids, EulerUpstreamResidual.hpp:110-112).
    Returns a dict of arrays plus the bid tables.
    """
    N = nx*ny*nz
    c, nbr, area, normal, fc, cc = _cartesian_slab_arrays(nx, ny, nz, dx, dy, dz, 0, nz)
    bnd = nbr < 0
    bid = np.zeros((N, 6), dtype=np.int64)
    if unique_bids:
        nb = int(bnd.sum())
        bid[bnd] = np.arange(1, nb + 1)
        n_bid = nb + 1
    else:
        side = np.broadcast_to(np.arange(1, 7), (N, 6))
        bid[bnd] = side[bnd]
        n_bid = 7
    bid_kind = np.zeros(n_bid, dtype=np.int32)
    bid_sat = np.ones(n_bid)               # default SatBC: Dirichlet 1.0 (BoundaryConditions.hpp:179-182)
    bid_partner = np.zeros(n_bid, dtype=np.int32)
    if any(periodic):
        assert unique_bids, "periodic conditions need unique boundary ids"
        strides = (1, nx, nx*ny)
        ext = (nx, ny, nz)
        for d in range(3):
            if not periodic[d]:
                continue
            lo = np.nonzero(bnd[:, 2*d])[0]
            hi = lo + (ext[d] - 1)*strides[d]
            b_lo = bid[lo, 2*d]
            b_hi = bid[hi, 2*d + 1]
            bid_kind[b_lo] = 1
            bid_kind[b_hi] = 1
            bid_partner[b_lo] = b_hi
            bid_partner[b_hi] = b_lo
    return dict(
        N=N,
        hf_offset=_i32(np.arange(N + 1)*6),
        hf_nbr=_i32(nbr.reshape(-1)),
        hf_bid=_i32(bid.reshape(-1)),
        hf_area=_f64(area.reshape(-1)),
        hf_normal=_f64(normal.reshape(-1, 3)),
        hf_centroid=_f64(fc.reshape(-1, 3)),
        cell_volume=_f64(np.full(N, dx*dy*dz)),
        cell_centroid=_f64(cc),
        bid_kind=bid_kind, bid_sat=_f64(bid_sat), bid_partner=bid_partner,
        dims=(nx, ny, nz),
    )


# ----------------------------------------------------------------------------------
# Faulted corner-point style grid (vertical pillars, columns shifted in z at fault planes)
# ----------------------------------------------------------------------------------
def faulted_grid(nx, ny, nz, dx=10.0, dy=10.0, dz=1.0, faults_i=(), faults_j=(), unique_bids=False, k0=0, k1=None):
    """Vertical-pillar corner-point grid whose cell columns are shifted in z at fault planes.

    ``faults_i`` = [(i_plane, throw_in_layers), ...]: columns with i >= i_plane are shifted up by
    ``throw`` layers relative to i_plane-1 (likewise ``faults_j``).  Non-integer throws split each
    lateral face on the fault plane into two half-faces with different neighbours (so cells get 7
    or 8 faces, and neighbours lie in other k-layers); the unmatched parts at the top and bottom
    of a column become boundary faces.  Cell index c = i + nx*(j + ny*k) as for CpGrid.
    ``k0``, ``k1``: only the cells of the layers [k0, k1) (neighbour ids stay global): the slab a rank generates.
    """
    k1 = nz if k1 is None else k1
    N = nx*ny*(k1 - k0)
    sx = np.zeros(nx)
    for ip, t in faults_i:
        sx[ip:] += t
    sy = np.zeros(ny)
    for jp, t in faults_j:
        sy[jp:] += t
    c = np.arange(N, dtype=np.int64) + k0*nx*ny
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx*ny)
    shift = sx[i] + sy[j]                      # in layers
    zlo = (k + shift)*dz
    # slots: 0 x-(a) 1 x-(b) 2 x+(a) 3 x+(b) 4 y-(a) 5 y-(b) 6 y+(a) 7 y+(b) 8 z- 9 z+
    S = 10
    valid = np.zeros((N, S), dtype=bool)
    nbr = np.full((N, S), -1, dtype=np.int64)
    area = np.zeros((N, S))
    normal = np.zeros((N, S, 3))
    cc = np.stack([(i + 0.5)*dx, (j + 0.5)*dy, zlo + 0.5*dz], axis=1)
    fc = np.repeat(cc[:, None, :], S, axis=1)

    def lateral(slot, axis, sgn):
        """Faces of side (axis, sgn): neighbour column at +-1 along axis."""
        if axis == 0:
            other_ok = (i + sgn >= 0) & (i + sgn < nx)
            io = np.clip(i + sgn, 0, nx - 1)
            rel = sx[io] - sx[i]               # shift of the neighbour column relative to this one
            stride = sgn
            width = dy
            fpos = (i + (1 if sgn > 0 else 0))*dx
        else:
            other_ok = (j + sgn >= 0) & (j + sgn < ny)
            jo = np.clip(j + sgn, 0, ny - 1)
            rel = sy[jo] - sy[j]
            stride = sgn*nx
            width = dx
            fpos = (j + (1 if sgn > 0 else 0))*dy
        m = np.floor(rel)
        f = rel - m                            # 0 <= f < 1
        # lower part [k, k+f] faces neighbour layer k-m-1; upper part [k+f, k+1] faces layer k-m
        ka = (k - m - 1).astype(np.int64)
        kb = (k - m).astype(np.int64)
        has_a = f > 0
        for part, kk, z0, z1, present in ((0, ka, np.zeros(N), f, has_a), (1, kb, f, np.ones(N), np.ones(N, bool))):
            sl = slot + part
            valid[:, sl] = present
            inside = other_ok & (kk >= 0) & (kk < nz) & present
            nb = c + stride + (kk - k)*nx*ny
            nbr[:, sl] = np.where(inside, nb, -1)
            area[:, sl] = width*(z1 - z0)*dz
            normal[:, sl, axis] = float(sgn)
            fc[:, sl, axis] = fpos
            fc[:, sl, 2] = zlo + 0.5*(z0 + z1)*dz
        # when the neighbour column does not exist the whole side is one boundary face
        none = ~other_ok
        valid[none, slot] = False
        area[none, slot + 1] = width*dz
        fc[none, slot + 1, 2] = zlo[none] + 0.5*dz

    lateral(0, 0, -1)
    lateral(2, 0, +1)
    lateral(4, 1, -1)
    lateral(6, 1, +1)
    valid[:, 8] = True
    valid[:, 9] = True
    nbr[:, 8] = np.where(k > 0, c - nx*ny, -1)
    nbr[:, 9] = np.where(k < nz - 1, c + nx*ny, -1)
    area[:, 8:10] = dx*dy
    normal[:, 8, 2] = -1.0
    normal[:, 9, 2] = 1.0
    fc[:, 8, 2] = zlo
    fc[:, 9, 2] = zlo + dz

    counts = valid.sum(axis=1)
    hf_offset = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=hf_offset[1:])
    sel = valid.reshape(-1)
    hf_nbr = nbr.reshape(-1)[sel]
    bnd = hf_nbr < 0
    H = hf_nbr.shape[0]
    hf_bid = np.zeros(H, dtype=np.int64)
    if unique_bids:
        nb = int(bnd.sum())
        hf_bid[bnd] = np.arange(1, nb + 1)
        n_bid = nb + 1
    else:
        side = np.broadcast_to(np.array([1, 1, 2, 2, 3, 3, 4, 4, 5, 6]), (N, S)).reshape(-1)[sel]
        hf_bid[bnd] = side[bnd]
        n_bid = 7
    return dict(
        N=N, first_cell=k0*nx*ny,
        hf_offset=_i32(hf_offset),
        hf_nbr=_i32(hf_nbr),
        hf_bid=_i32(hf_bid),
        hf_area=_f64(area.reshape(-1)[sel]),
        hf_normal=_f64(normal.reshape(-1, 3)[sel]),
        hf_centroid=_f64(fc.reshape(-1, 3)[sel]),
        cell_volume=_f64(np.full(N, dx*dy*dz)),
        cell_centroid=_f64(cc),
        bid_kind=np.zeros(n_bid, dtype=np.int32), bid_sat=np.ones(n_bid), bid_partner=np.zeros(n_bid, dtype=np.int32),
        dims=(nx, ny, nz),
    )


# ----------------------------------------------------------------------------------
# Properties, fluxes, tables
# ----------------------------------------------------------------------------------
def rotation_tensor(kdiag, angle_z_deg=30.0, angle_x_deg=15.0):
    """K = R diag(k) R^T with a fixed rotation; symmetrised exactly."""
    az, ax = np.deg2rad(angle_z_deg), np.deg2rad(angle_x_deg)
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    R = Rz @ Rx
    K = R @ np.diag(kdiag) @ R.T
    K = 0.5*(K + K.T)
    return K


def corey_table(npts=24, swir=0.1, sor=0.1, j_scale=0.3, full_range=True):
    """Corey-like single rock: krw=Se^2, kro=(1-Se)^2, J = j_scale*(1-Se)/sqrt(Se) clipped.
    ``full_range`` pads the table to S in [0,1] so that any physical saturation is inside it."""
    s = np.linspace(swir, 1.0 - sor, npts)
    se = (s - swir)/(1.0 - swir - sor)
    krw = se**2
    kro = (1.0 - se)**2
    J = j_scale*(1.0 - se)/np.sqrt(np.maximum(se, 0.02))
    if full_range:
        s = np.concatenate([[0.0], s, [1.0]])
        krw = np.concatenate([[0.0], krw, [krw[-1]]])
        kro = np.concatenate([[kro[0]], kro, [0.0]])
        J = np.concatenate([[J[0]*1.5], J, [J[-1]]])
    krw[0] = 0.0
    kro[-1] = 0.0
    return RockTable(s=_f64(s), krw=_f64(krw), kro=_f64(kro), J=_f64(J))


def constant_velocity_flux(g, v):
    """Half-face fluxes of a constant velocity field: (v . n) * area
    (attic/euler/EulerSolverTester.hpp:73-91)."""
    n = g["hf_normal"]
    vn = (v[0]*n[:, 0] + v[1]*n[:, 1]) + v[2]*n[:, 2]
    return _f64(vn*g["hf_area"])


def mt_uniform(seed, n):
    """Reproducible uniforms in [0,1): 53 high bits of MT19937-64 style raw words
    (numpy's MT19937 generator; SURVEY 8d forbids library normal distributions)."""
    rs = np.random.Generator(np.random.MT19937(seed))
    raw = rs.integers(0, 2**64, size=n, dtype=np.uint64)
    return (raw >> np.uint64(11)).astype(np.float64)*(2.0**-53)


def box_muller(seed, n):
    u1 = mt_uniform(seed, n)
    u2 = mt_uniform(seed + 1000003, n)
    return np.sqrt(-2.0*np.log(1.0 - u1))*np.cos(2.0*np.pi*u2)


def make_case(name, g, **kw):
    N = g["N"]
    fields = dict(
        name=name, N=N, hf_offset=g["hf_offset"], hf_nbr=g["hf_nbr"], hf_bid=g["hf_bid"],
        hf_area=g["hf_area"], hf_normal=g["hf_normal"], hf_centroid=g["hf_centroid"],
        cell_volume=g["cell_volume"], cell_centroid=g["cell_centroid"],
        bid_kind=g["bid_kind"], bid_sat=g["bid_sat"], bid_partner=g["bid_partner"], dims=g.get("dims"),
    )
    fields.update(kw)
    c = Case(**fields)
    if c.src_cell is None:
        c.src_cell = np.zeros(0, dtype=np.int32)
        c.src_rate = np.zeros(0)
    c.src_cell = _i32(c.src_cell)
    c.src_rate = _f64(c.src_rate)
    c.gravity = _f64(c.gravity)
    c.sat0 = _f64(c.sat0)
    c.hf_flux = _f64(c.hf_flux)
    c.poro = _f64(c.poro)
    c.perm = _f64(c.perm).reshape(N, 9)
    if c.rock_id is not None:
        c.rock_id = _i32(c.rock_id)
    return c


# ----------------------------------------------------------------------------------
# The BASELINE.json configurations (SURVEY 8d "Synthetic inputs")
# ----------------------------------------------------------------------------------
def config_c1(n=10):
    """10x10x10 unit cube, K = 100 mD, phi = 0.2, no rocks, all-periodic, viscous+gravity."""
    g = cartesian_grid(n, n, n, 1.0/n, 1.0/n, 1.0/n, unique_bids=True, periodic=(True, True, True))
    N = g["N"]
    perm = np.zeros((N, 9))
    perm[:, [0, 4, 8]] = 100.0*MILLIDARCY
    sat0 = np.zeros(N)
    sat0[N//3:2*N//3] = 1.0                     # attic/euler/EulerSolverTester.hpp:169-170
    return make_case("C1", g, poro=np.full(N, 0.2), perm=perm, sat0=sat0,
                     gravity=[0.0, 0.0, -9.80665],
                     hf_flux=constant_velocity_flux(g, (1e-6, 5e-7, 2.5e-7)),
                     method_capillary=False, time=86400.0)


def config_c2(n=100, seed=7):
    """n^3 Cartesian, dx = 1 m, rotated anisotropic K, one Corey rock with J-scaling, V+G+C."""
    g = cartesian_grid(n, n, n, 1.0, 1.0, 1.0, unique_bids=False)
    N = g["N"]
    K = rotation_tensor(np.array([200.0, 100.0, 10.0])*MILLIDARCY)
    perm = np.broadcast_to(K.reshape(1, 9), (N, 9)).copy()
    sat0 = 0.2 + 0.1*(mt_uniform(seed, N) - 0.5)
    g["bid_sat"][:] = 1.0
    return make_case("C2", g, poro=np.full(N, 0.2), perm=perm, rock_id=np.zeros(N, dtype=np.int32),
                     rocks=[corey_table()], use_j=True, sigma=1.0, theta=0.0, sat0=sat0,
                     gravity=[0.0, 0.0, -9.80665],
                     hf_flux=constant_velocity_flux(g, (1e-6, 5e-7, 2.5e-7)), time=86400.0)


def config_c2b(n=100, seed=7):
    """SURVEY 8d variant C2b: the C2 grid with the property class aniso_simulator_test really uses,
    ReservoirPropertyCapillaryAnisotropicRelperm (diagonal tensor mobility): krxx != kryy != krzz tables, pc table
    without J-scaling (RockAnisotropicRelperm.hpp:57-62,79-83)."""
    c = config_c2(n, seed)
    t = c.rocks[0]
    f = np.array([1.0, 0.7, 0.4])
    c.rocks = [RockTable(s=t.s, pc=_f64(t.J*2.0e4), kr_w=_f64(t.krw[:, None]*f[None, :]), kr_o=_f64(t.kro[:, None]*f[None, ::-1]))]
    c.mobility_kind = 1
    c.use_j = False
    c.name = "C2b"
    return c


def lognormal_perm(N, seed, mean_md=100.0, sigma=1.0, kz_ratio=0.1):
    z = box_muller(seed, N)
    kx = np.exp(np.log(mean_md*MILLIDARCY) + sigma*z)
    perm = np.zeros((N, 9))
    perm[:, 0] = kx
    perm[:, 4] = kx
    perm[:, 8] = kz_ratio*kx
    return perm


def config_c3(nx=256, ny=256, nz=128, seed=42, random_rocks=False):
    """Faulted corner-point grid, lognormal K, 3 rock types in layer bands (or drawn per cell), V+G+C."""
    fi = [(nx//4, 1.5), (nx//2, 2.25), (3*nx//4, 0.75)]
    fj = [(ny//2, 3.0)]
    g = faulted_grid(nx, ny, nz, 10.0, 10.0, 1.0, faults_i=fi, faults_j=fj)
    N = g["N"]
    k = np.arange(N)//(nx*ny)
    rock_id = np.minimum((3*k)//nz, 2).astype(np.int32)
    if random_rocks:
        rock_id = np.minimum((3.0*mt_uniform(seed + 3, N)).astype(np.int32), 2)
    rocks = [corey_table(24, 0.10, 0.10, 0.30), corey_table(20, 0.15, 0.05, 0.22), corey_table(28, 0.05, 0.20, 0.40)]
    sat0 = 0.25 + 0.1*(mt_uniform(seed + 2, N) - 0.5)
    return make_case("C3", g, poro=0.05 + 0.25*mt_uniform(seed + 1, N), perm=lognormal_perm(N, seed),
                     rock_id=rock_id, rocks=rocks, sat0=sat0, gravity=[0.0, 0.0, -9.80665],
                     hf_flux=constant_velocity_flux(g, (1e-6, 0.0, 0.0)), time=86400.0)


def config_c4(nx=512, ny=512, nz=256, seed=44, capillary=False):
    """Cartesian heterogeneous-perm grid for the strong-scaling runs, 1 rock, V+G (+C)."""
    g = cartesian_grid(nx, ny, nz, 1.0, 1.0, 1.0, unique_bids=False)
    N = g["N"]
    sat0 = 0.3 + 0.2*(mt_uniform(seed + 2, N) - 0.5)
    return make_case("C4", g, poro=0.05 + 0.25*mt_uniform(seed + 1, N), perm=lognormal_perm(N, seed),
                     rock_id=np.zeros(N, dtype=np.int32), rocks=[corey_table()], sat0=sat0,
                     gravity=[0.0, 0.0, -9.80665], method_capillary=capillary,
                     hf_flux=constant_velocity_flux(g, (1e-6, 5e-7, 2.5e-7)), time=86400.0)


def random_geometry_case(nx, ny, nz, seed=1, periodic=(False, False, False), n_rocks=0, full_tensor=True,
                         sources=True, mobility_kind=0, use_j=True, perturb_normals=True):
    """Stress case: Cartesian topology with randomised (but pairwise consistent) face normals,
    areas and centroids, random full-tensor K, random porosity and saturations, oblique flux.
    The transport scheme never checks geometric consistency, so this exercises every term
    of the face flux with generic operands (oblique normals, off-diagonal K).
    ``perturb_normals=False`` keeps the axis-aligned normals of the Cartesian grid (everything else stays
    random): the geometry class on which FAST mode covers the diagonal tensor mobility."""
    g = cartesian_grid(nx, ny, nz, 1.0, 0.8, 0.5, unique_bids=True, periodic=periodic)
    N, H = g["N"], g["hf_nbr"].shape[0]
    rs = np.random.Generator(np.random.MT19937(seed))
    normal = g["hf_normal"] + 0.3*(rs.random((H, 3)) - 0.5)*(1.0 if perturb_normals else 0.0)
    normal /= np.sqrt((normal**2).sum(axis=1))[:, None]
    area = g["hf_area"]*(0.7 + 0.6*rs.random(H))
    cent = g["hf_centroid"] + 0.05*(rs.random((H, 3)) - 0.5)
    # make the two half-faces of an interior face consistent (same area/centroid, opposite normal)
    nbr = g["hf_nbr"].astype(np.int64)
    cell_of = np.repeat(np.arange(N), 6)
    local = np.tile(np.arange(6), N)
    twin = nbr*6 + (local ^ 1)
    upper = (nbr >= 0) & (cell_of > nbr)
    normal[upper] = -normal[twin[upper]]
    area[upper] = area[twin[upper]]
    cent[upper] = cent[twin[upper]]
    g["hf_normal"], g["hf_area"], g["hf_centroid"] = _f64(normal), _f64(area), _f64(cent)
    g["cell_centroid"] = _f64(g["cell_centroid"] + 0.05*(rs.random((N, 3)) - 0.5))
    g["cell_volume"] = _f64(g["cell_volume"]*(0.8 + 0.4*rs.random(N)))
    if full_tensor:
        A = rs.random((N, 3, 3)) - 0.5
        K = np.einsum("nij,nkj->nik", A, A) + 0.3*np.eye(3)[None]
        K = 0.5*(K + K.transpose(0, 2, 1))*100.0*MILLIDARCY
    else:
        K = np.zeros((N, 3, 3))
        for d in range(3):
            K[:, d, d] = (50.0 + 100.0*rs.random(N))*MILLIDARCY
    rocks, rock_id = [], None
    if n_rocks > 0:
        if mobility_kind == 0:
            rocks = [corey_table(12 + 5*r, 0.05 + 0.03*r, 0.08 + 0.02*r, 0.2 + 0.1*r) for r in range(n_rocks)]
        else:
            rocks = []
            for r in range(n_rocks):
                t = corey_table(12 + 5*r, 0.05 + 0.03*r, 0.08 + 0.02*r, 0.2 + 0.1*r)
                f = np.array([1.0, 0.8 - 0.1*r, 0.5 + 0.1*r])
                rocks.append(RockTable(s=t.s, pc=_f64(t.J*2.0e4), kr_w=_f64(t.krw[:, None]*f[None, :]),
                                       kr_o=_f64(t.kro[:, None]*f[None, ::-1])))
        rock_id = rs.integers(0, n_rocks, size=N).astype(np.int32)
    nb = g["bid_kind"].shape[0]
    g["bid_sat"] = _f64(rs.random(nb))
    src_cell = src_rate = None
    if sources:
        src_cell = np.array(sorted(rs.choice(N, size=min(4, N), replace=False)), dtype=np.int32)
        src_rate = (rs.random(src_cell.shape[0]) - 0.5)*2e-7
    v = (1e-6, -4e-7, 2.5e-7)
    return make_case("rand", g, poro=0.1 + 0.2*rs.random(N), perm=K.reshape(N, 9), rock_id=rock_id, rocks=rocks,
                     use_j=(use_j if mobility_kind == 0 else False), sigma=0.03, theta=0.3,
                     mobility_kind=mobility_kind,
                     sat0=0.05 + 0.9*rs.random(N), gravity=[0.3, -0.2, -9.80665],
                     hf_flux=constant_velocity_flux(g, v), src_cell=src_cell, src_rate=src_rate, time=3600.0)


# ----------------------------------------------------------------------------------
# Streaming generation of the big Cartesian configuration (C4) in z-slabs, so that neither
# the host nor the upload ever holds the full 27 GB "fat" grid description at once.
# ----------------------------------------------------------------------------------
def plane_uniform(seed, k, n):
    """Uniforms of plane k of a field: independent of how the grid is chunked or partitioned."""
    return mt_uniform(seed*1000003 + k, n)


def c4_slab(nx, ny, nz, k0, k1, seed=44, velocity=(1e-6, 5e-7, 2.5e-7), dirichlet_sat=1.0):
    """Cells with k in [k0, k1) of the C4 configuration (512x512x256 Cartesian, dx = 1 m,
    lognormal K with kz = 0.1 kx, phi in [0.05, 0.30], default Dirichlet S = 1 boundaries,
    constant-velocity flux, S0 = 0.3 +- 0.1).  Returns a dict in eu_grid_chunk form (global ids)
    plus ``sat0`` and ``hf_flux`` for these cells."""
    npl = nx*ny
    c, nbr, area, normal, fc, cc = _cartesian_slab_arrays(nx, ny, nz, 1.0, 1.0, 1.0, k0, k1)
    n = c.shape[0]
    kx = np.empty(n)
    poro = np.empty(n)
    sat0 = np.empty(n)
    for k in range(k0, k1):
        sl = slice((k - k0)*npl, (k - k0 + 1)*npl)
        u1 = plane_uniform(seed, k, npl)
        u2 = plane_uniform(seed + 7, k, npl)
        z = np.sqrt(-2.0*np.log(1.0 - u1))*np.cos(2.0*np.pi*u2)
        kx[sl] = np.exp(np.log(100.0*MILLIDARCY) + z)
        poro[sl] = 0.05 + 0.25*plane_uniform(seed + 1, k, npl)
        sat0[sl] = 0.3 + 0.2*(plane_uniform(seed + 2, k, npl) - 0.5)
    perm = np.zeros((n, 9))
    perm[:, 0] = kx
    perm[:, 4] = kx
    perm[:, 8] = 0.1*kx
    nbr_flat = nbr.reshape(-1)
    bnd_hf = np.nonzero(nbr_flat < 0)[0]
    nrm = normal.reshape(-1, 3)
    a = area.reshape(-1)
    flux = ((velocity[0]*nrm[:, 0] + velocity[1]*nrm[:, 1]) + velocity[2]*nrm[:, 2])*a
    return dict(
        first_cell=int(c[0]), n_cells=n,
        hf_count=_i32(np.full(n, 6)), hf_neighbour=_i32(nbr_flat),
        hf_area=_f64(a), hf_normal=_f64(nrm), hf_centroid=_f64(fc.reshape(-1, 3)),
        bnd_hf=_i32(bnd_hf), bnd_kind=_i32(np.full(bnd_hf.shape[0], 1)),
        bnd_sat=_f64(np.full(bnd_hf.shape[0], dirichlet_sat)),
        bnd_partner_cell=_i32(np.full(bnd_hf.shape[0], -1)), bnd_partner_face=_i32(np.full(bnd_hf.shape[0], -1)),
        cell_volume=_f64(np.ones(n)), cell_centroid=_f64(cc), porosity=_f64(poro), permeability=_f64(perm),
        rock_id=_i32(np.zeros(n)), sat0=_f64(sat0), hf_flux=_f64(flux),
    )


C5_FAULTS_I = ((0.25, 1.5), (0.5, 2.25), (0.75, 0.75))      # (fraction of nx, throw in layers), as in config_c3
C5_FAULTS_J = ((0.5, 3.0),)
C5_GHOST_DEPTH = 4                                           # layers a lateral neighbour can be away (max throw 3, split faces)


def c5_rocks():
    return [corey_table(24, 0.10, 0.10, 0.30), corey_table(20, 0.15, 0.05, 0.22), corey_table(28, 0.05, 0.20, 0.40)]


def c5_slab(nx, ny, nz, k0, k1, seed=42, velocity=(1e-6, 0.0, 0.0), dirichlet_sat=1.0):
    """Layers [k0, k1) of the weak-scaling configuration C5 (BASELINE configs[4]): the C3 generator -- faulted
    corner-point grid, dx = dy = 10 m, dz = 1 m, lognormal K with kz = 0.1 kx, phi in [0.05, 0.30], three rock types in
    layer bands, Dirichlet S = 1 boundaries, flux of a constant velocity projected on the true face normals -- at
    nx x ny x nz with nz = 122 layers per GPU.  Properties are seeded per layer, so the grid does not depend on how it
    is partitioned.  Returns a dict in eu_grid_chunk form (global ids) plus ``sat0`` and ``hf_flux``."""
    fi = [(int(f*nx), t) for f, t in C5_FAULTS_I]
    fj = [(int(f*ny), t) for f, t in C5_FAULTS_J]
    g = faulted_grid(nx, ny, nz, 10.0, 10.0, 1.0, faults_i=fi, faults_j=fj, k0=k0, k1=k1)
    npl = nx*ny
    n = g["N"]
    kx = np.empty(n)
    poro = np.empty(n)
    sat0 = np.empty(n)
    rock = np.empty(n, dtype=np.int32)
    for k in range(k0, k1):
        sl = slice((k - k0)*npl, (k - k0 + 1)*npl)
        u1 = plane_uniform(seed, k, npl)
        u2 = plane_uniform(seed + 7, k, npl)
        z = np.sqrt(-2.0*np.log(1.0 - u1))*np.cos(2.0*np.pi*u2)
        kx[sl] = np.exp(np.log(100.0*MILLIDARCY) + z)
        poro[sl] = 0.05 + 0.25*plane_uniform(seed + 1, k, npl)
        sat0[sl] = 0.25 + 0.1*(plane_uniform(seed + 2, k, npl) - 0.5)
        rock[sl] = min((3*k)//nz, 2)
    perm = np.zeros((n, 9))
    perm[:, 0] = kx
    perm[:, 4] = kx
    perm[:, 8] = 0.1*kx
    nbr = g["hf_nbr"]
    bnd_hf = np.nonzero(nbr < 0)[0]
    nrm = g["hf_normal"]
    flux = ((velocity[0]*nrm[:, 0] + velocity[1]*nrm[:, 1]) + velocity[2]*nrm[:, 2])*g["hf_area"]
    return dict(
        first_cell=int(g["first_cell"]), n_cells=n,
        hf_count=_i32(np.diff(g["hf_offset"])), hf_neighbour=_i32(nbr),
        hf_area=g["hf_area"], hf_normal=nrm, hf_centroid=g["hf_centroid"],
        bnd_hf=_i32(bnd_hf), bnd_kind=_i32(np.full(bnd_hf.shape[0], 1)),
        bnd_sat=_f64(np.full(bnd_hf.shape[0], dirichlet_sat)),
        bnd_partner_cell=_i32(np.full(bnd_hf.shape[0], -1)), bnd_partner_face=_i32(np.full(bnd_hf.shape[0], -1)),
        cell_volume=g["cell_volume"], cell_centroid=g["cell_centroid"], porosity=_f64(poro), permeability=_f64(perm),
        rock_id=rock, sat0=_f64(sat0), hf_flux=_f64(flux),
    )


def c5_fluid_case():
    """The fluid / rock / solver-parameter part of C5 as a tiny Case (no grid arrays): three rock tables, V+G+C."""
    g = cartesian_grid(1, 1, 1, unique_bids=False)
    perm = np.zeros((1, 9))
    perm[0, [0, 4, 8]] = 100.0*MILLIDARCY
    return make_case("C5-fluid", g, poro=np.full(1, 0.2), perm=perm, rock_id=np.zeros(1, dtype=np.int32),
                     rocks=c5_rocks(), sat0=np.zeros(1), gravity=[0.0, 0.0, -9.80665], hf_flux=np.zeros(6))


def c4_fluid_case(capillary=False):
    """The fluid/rock/solver-parameter part of C4 as a tiny Case (no grid arrays)."""
    g = cartesian_grid(1, 1, 1, unique_bids=False)
    perm = np.zeros((1, 9))
    perm[0, [0, 4, 8]] = 100.0*MILLIDARCY
    return make_case("C4-fluid", g, poro=np.full(1, 0.2), perm=perm, rock_id=np.zeros(1, dtype=np.int32),
                     rocks=[corey_table()], sat0=np.zeros(1), gravity=[0.0, 0.0, -9.80665],
                     hf_flux=np.zeros(6), method_capillary=capillary)


# ----------------------------------------------------------------------------------
# Decomposition helpers: the part of a (small, fully materialised) Case that one rank of an
# index-range (z-slab) decomposition uploads -- own cells plus the ghost cells its faces touch.
# ----------------------------------------------------------------------------------
def extract_slab(case, own_begin, own_end):
    """Returns dict(cells, hf_index, chunks, n_local, n_hf): ``cells`` = ascending global ids of the local
    (own + ghost) cells, ``hf_index`` = global half-face indices of their half-faces (so that
    sat_local = sat[cells], flux_local = hf_flux[hf_index]) and ``chunks`` in eu_grid_chunk layout."""
    from .binding import resolve_boundary
    off = case.hf_offset.astype(np.int64)
    own = np.arange(own_begin, own_end)
    bnd_hf, kind, sat, pcell, pface = resolve_boundary(case)
    h0, h1 = off[own_begin], off[own_end]
    nb = case.hf_nbr[h0:h1].astype(np.int64)
    touched = [nb[nb >= 0]]
    b0, b1 = np.searchsorted(bnd_hf, [h0, h1])
    pc = pcell[b0:b1]
    touched.append(pc[pc >= 0].astype(np.int64))
    cells = np.union1d(own, np.concatenate(touched))
    counts = (off[cells + 1] - off[cells]).astype(np.int64)
    hf_index = np.concatenate([np.arange(off[c], off[c + 1]) for c in cells]) if cells.size else np.zeros(0, np.int64)
    # contiguous runs of cells -> chunks
    breaks = np.nonzero(np.diff(cells) != 1)[0] + 1
    starts = np.concatenate([[0], breaks])
    ends = np.concatenate([breaks, [cells.size]])
    hf_starts = np.concatenate([[0], np.cumsum(counts)])
    chunks = []
    for a, b in zip(starts, ends):
        c0, c1 = int(cells[a]), int(cells[b - 1]) + 1
        g0, g1 = int(off[c0]), int(off[c1])
        k0, k1 = np.searchsorted(bnd_hf, [g0, g1])
        chunks.append(dict(
            first_cell=c0, n_cells=c1 - c0,
            hf_count=_i32(off[c0 + 1:c1 + 1] - off[c0:c1]), hf_neighbour=_i32(case.hf_nbr[g0:g1]),
            hf_area=_f64(case.hf_area[g0:g1]), hf_normal=_f64(case.hf_normal[g0:g1]), hf_centroid=_f64(case.hf_centroid[g0:g1]),
            bnd_hf=_i32(bnd_hf[k0:k1] - g0), bnd_kind=_i32(kind[k0:k1]), bnd_sat=_f64(sat[k0:k1]),
            bnd_partner_cell=_i32(pcell[k0:k1]), bnd_partner_face=_i32(pface[k0:k1]),
            cell_volume=_f64(case.cell_volume[c0:c1]), cell_centroid=_f64(case.cell_centroid[c0:c1]),
            porosity=_f64(case.poro[c0:c1]), permeability=_f64(case.perm[c0:c1]),
            rock_id=_i32(case.rock_id[c0:c1]) if case.rock_id is not None else None))
    return dict(cells=cells, hf_index=hf_index, chunks=chunks, n_local=int(cells.size), n_hf=int(hf_index.size),
                own_begin=own_begin, own_end=own_end, hf_starts=hf_starts)


def local_case(case, slab):
    """A self-contained Case over the local cells of ``slab`` (local numbering), for running the CPU oracle
    on one rank's part: neighbours that are not local become dummy Dirichlet boundary faces (they only
    occur on ghost cells, whose results are discarded)."""
    import copy
    cells = slab["cells"]
    hfi = slab["hf_index"]
    g2l = np.full(case.N, -1, dtype=np.int64)
    g2l[cells] = np.arange(cells.size)
    nbr_g = case.hf_nbr[hfi].astype(np.int64)
    nbr_l = np.where(nbr_g >= 0, g2l[np.maximum(nbr_g, 0)], -1)
    bid = case.hf_bid[hfi].astype(np.int64).copy()
    n_bid = case.bid_kind.shape[0]
    missing = (nbr_g >= 0) & (nbr_l < 0)
    bid[missing] = n_bid                               # dummy Dirichlet id
    # periodic faces of ghost cells whose partner cell is not local -> dummy as well
    from .binding import resolve_boundary
    bnd_hf, kind, sat, pcell, pface = resolve_boundary(case)
    pc_of_hf = np.full(case.H, -2, dtype=np.int64)
    pc_of_hf[bnd_hf] = pcell
    p = pc_of_hf[hfi]
    lost = (p >= 0) & (g2l[np.maximum(p, 0)] < 0)
    bid[lost] = n_bid
    c = copy.copy(case)
    c.N = int(cells.size)
    c.hf_offset = _i32(slab["hf_starts"])
    c.hf_nbr = _i32(nbr_l)
    c.hf_bid = _i32(bid)
    c.hf_area, c.hf_normal, c.hf_centroid = _f64(case.hf_area[hfi]), _f64(case.hf_normal[hfi]), _f64(case.hf_centroid[hfi])
    c.cell_volume, c.cell_centroid = _f64(case.cell_volume[cells]), _f64(case.cell_centroid[cells])
    c.poro, c.perm = _f64(case.poro[cells]), _f64(case.perm[cells])
    c.rock_id = _i32(case.rock_id[cells]) if case.rock_id is not None else None
    c.bid_kind = _i32(np.concatenate([case.bid_kind, [0]]))
    c.bid_sat = _f64(np.concatenate([case.bid_sat, [0.5]]))
    c.bid_partner = _i32(np.concatenate([case.bid_partner, [0]]))
    c.sat0 = _f64(case.sat0[cells])
    c.hf_flux = _f64(case.hf_flux[hfi])
    own = (cells >= slab["own_begin"]) & (cells < slab["own_end"])
    keep = np.isin(case.src_cell, cells[own])
    c.src_cell = _i32(g2l[case.src_cell[keep]])
    c.src_rate = _f64(case.src_rate[keep])
    return c
