"""Host-side helpers of the C ABI for the callers and data formats on either side of the path (no device):
eu_build_face_indices (GridInterfaceEuler::buildFaceIndices), eu_post_process_fluxes
(IncompFlowSolverHybrid::postProcessFluxes) and eu_write_field (writeField).

writeField is pinned against the compiled reference (oracle/_ref) and a fixture generated from it.  The other two live in
classes that need Dune (GridInterfaceEuler<CpGrid>, IncompFlowSolverHybrid): they are checked against line-by-line Python
restatements of the reference's loops (cited below) and against the properties those loops guarantee."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT


def _lib():
    import opm_porsol_b200 as eub
    L = eub.load_library()
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.eu_build_face_indices.argtypes = [C.c_int, ip, ip, ip, ip, ip]
    L.eu_post_process_fluxes.argtypes = [C.c_int, ip, ip, ip, C.c_int, ip, dp, dp]
    L.eu_write_field.argtypes = [dp, C.c_longlong, C.c_char_p]
    return L


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _grids():
    from opm_porsol_b200 import synth
    g1 = synth.cartesian_grid(5, 4, 3)
    g2 = synth.faulted_grid(8, 6, 5, faults_i=[(3, 1.5), (6, 0.75)], faults_j=[(2, 2.0)])      # split faces, 7-8 faces per cell
    g3 = synth.cartesian_grid(4, 3, 2, periodic=(True, False, True))
    return [("cartesian", g1), ("faulted", g2), ("periodic", g3)]


def build_face_indices_py(off, nbr):
    """GridInterfaceEuler.hpp:524-599 with cell index == iteration order."""
    n = len(off) - 1
    faces, fpos = [], [0]
    for c in range(n):                                           # first pass: a face is discovered by the cell visited first
        for h in range(off[c], off[c + 1]):
            c1 = nbr[h]
            if c1 >= 0 and c1 > c:                               # cell[c1] == -1
                faces.append(c1)
        fpos.append(len(faces))
    total = len(faces)
    out = np.empty(len(nbr), dtype=np.int32)
    for c in range(n):                                           # second pass
        for h in range(off[c], off[c + 1]):
            c1 = nbr[h]
            if c1 < 0:
                out[h] = total
                total += 1
            else:
                t, seek = (c, c1) if c < c1 else (c1, c)
                out[h] = fpos[t] + faces[fpos[t]:fpos[t + 1]].index(seek)      # std::find: the first match
    return out, total


@pytest.mark.parametrize("name,g", _grids(), ids=[n for n, _ in _grids()])
def test_build_face_indices(name, g):
    L = _lib()
    off = np.ascontiguousarray(g["hf_offset"], dtype=np.int32)
    nbr = np.ascontiguousarray(np.where(g["hf_nbr"] < 0, -1, g["hf_nbr"]), dtype=np.int32)
    idx = np.full(nbr.shape[0], -7, dtype=np.int32)
    nf, mx = C.c_int(0), C.c_int(0)
    assert L.eu_build_face_indices(g["N"], _i(off), _i(nbr), _i(idx), C.byref(nf), C.byref(mx)) == 0
    want, total = build_face_indices_py(off.tolist(), nbr.tolist())
    assert np.array_equal(idx, want) and nf.value == total
    assert mx.value == int(np.diff(off).max())
    # properties: boundary faces are numbered after all interior ones, each exactly once; an interior number is
    # shared by exactly the half-faces of one pair of cells
    bnd = nbr < 0
    n_int = int((~bnd).sum())//2
    assert sorted(idx[bnd].tolist()) == list(range(total - int(bnd.sum()), total))
    assert idx[~bnd].max() < total - int(bnd.sum()) and len(set(idx[~bnd].tolist())) <= n_int
    cell_of = np.repeat(np.arange(g["N"]), np.diff(off))
    pairs = {}
    for h in np.nonzero(~bnd)[0]:
        key = (min(cell_of[h], nbr[h]), max(cell_of[h], nbr[h]))
        pairs.setdefault(int(idx[h]), set()).add(key)
    assert all(len(v) == 1 for v in pairs.values())


def post_process_py(off, nbr, fidx, nf, partner, flux):
    """IncompFlowSolverHybrid.hpp:707-807 (FaceFluxes::put / get, the two passes of postProcessFluxes)."""
    acc, vis, maxmod = [0.0]*nf, [0]*nf, 0.0
    flux = list(flux)

    def put(v, f):
        acc[f] += (-1.0 if vis[f] else 1.0)*v
        vis[f] += 1
    for h in range(len(nbr)):
        f = fidx[h]
        if nbr[h] < 0:
            if partner is None:
                continue
            if partner[f] != -1:
                put(flux[h], f)
                put(flux[h], partner[f])
        else:
            put(flux[h], f)
    vis = [0]*nf
    for h in range(len(nbr)):
        f = fidx[h]

        def get(v, ff):
            nonlocal maxmod
            new = 0.5*(-1.0 if vis[ff] else 1.0)*acc[ff]
            maxmod = max(maxmod, abs(v - new))
            vis[ff] += 1
            return new
        if nbr[h] < 0:
            if partner is None:
                continue
            if partner[f] != -1:
                flux[h] = get(flux[h], f)
                get(flux[h], partner[f])
        else:
            flux[h] = get(flux[h], f)
    return np.array(flux), maxmod


@pytest.mark.parametrize("name,g", _grids(), ids=[n for n, _ in _grids()])
def test_post_process_fluxes(name, g):
    from opm_porsol_b200 import synth
    L = _lib()
    off = np.ascontiguousarray(g["hf_offset"], dtype=np.int32)
    nbr = np.ascontiguousarray(np.where(g["hf_nbr"] < 0, -1, g["hf_nbr"]), dtype=np.int32)
    fidx = np.empty(nbr.shape[0], dtype=np.int32)
    nf = C.c_int(0)
    assert L.eu_build_face_indices(g["N"], _i(off), _i(nbr), _i(fidx), C.byref(nf), None) == 0
    rng = np.random.default_rng(3)
    flux = synth.constant_velocity_flux(g, (1e-6, -3e-7, 2e-7))*(1.0 + 1e-3*rng.standard_normal(nbr.shape[0]))
    partner = None
    if name == "periodic":
        # periodic partner table per unique face: pair the boundary faces of the periodic sides through the boundary ids
        partner = np.full(nf.value, -1, dtype=np.int32)
        bid, part = g["hf_bid"], g["bid_partner"]
        face_of_bid = {int(bid[h]): int(fidx[h]) for h in np.nonzero(nbr < 0)[0]}
        for b, f in face_of_bid.items():
            if g["bid_kind"][b] == 1 and int(part[b]) in face_of_bid:
                partner[f] = face_of_bid[int(part[b])]
    got = flux.copy()
    mm = C.c_double(0.0)
    assert L.eu_post_process_fluxes(g["N"], _i(off), _i(nbr), _i(fidx), nf.value, None if partner is None else _i(partner),
                                    _d(got), C.byref(mm)) == 0
    want, maxmod = post_process_py(off.tolist(), nbr.tolist(), fidx.tolist(), nf.value, None if partner is None else partner.tolist(), flux)
    assert np.array_equal(got, want) and mm.value == maxmod and maxmod > 0.0
    # twin half-faces are exactly antisymmetric afterwards; boundary fluxes without a partner are untouched
    cell_of = np.repeat(np.arange(g["N"]), np.diff(off))
    by_face = {}
    for h in np.nonzero(nbr >= 0)[0]:
        by_face.setdefault(int(fidx[h]), []).append(got[h])
    if name != "faulted":                      # (the faulted grid has pairs of cells joined by two faces sharing a number)
        assert all(len(v) == 2 and v[0] == -v[1] for v in by_face.values())
    if partner is None:
        assert np.array_equal(got[nbr < 0], flux[nbr < 0])


def test_write_field(tmp_path):
    from oracle import ref as oracle
    L = _lib()
    rng = np.random.default_rng(11)
    field = np.concatenate([rng.random(40), [0.0, 1.0, 0.1, 1e-7, 123456.789, 2.5e10, -3.0e-12, 1.0/3.0]])
    mine = tmp_path/"mine.sat"
    assert L.eu_write_field(_d(field), field.shape[0], str(mine).encode()) == 0
    fixture = os.path.join(ROOT, "tests", "golden", "write_field.txt")
    if oracle.ref_available():
        theirs = tmp_path/"ref.sat"
        assert oracle.ref_write_field(field, str(theirs)) == 0
        assert mine.read_bytes() == theirs.read_bytes()
        if not os.path.exists(fixture):       # generated once from the reference, then committed
            open(fixture, "wb").write(theirs.read_bytes())
    assert mine.read_bytes() == open(fixture, "rb").read()
    assert L.eu_write_field(_d(field), field.shape[0], str(tmp_path/"no"/"such"/"dir.sat").encode()) != 0
