"""Work units of the box kernel's sweep (host logic, no device): eu_debug_box_units -> eu_box_make_units (csrc/eu_host.cpp).

Every own (tile, plane) pair is swept exactly once; units stay inside one tile and inside the own planes; with neighbour
ranks the planes next to a slab boundary sit in ONE flagged unit per tile (the exchange's counters assume that:
eu_api.cu launch_substep sets halo.total to the tile count), and a block's flagged units come first in its list (it
signals the neighbour behind its last one).  Both partitions: z-chunks handed out round-robin (default) and equal spans
of the tile-major sweep (EU_BOX_UNITS=spans, single-rank runs)."""
import ctypes as C

import numpy as np
import pytest


def _units(nx, ny, tx, ty, z_lo, z_hi, bnd_lo, bnd_hi, blocks, spans=0, lz=0):
    import opm_porsol_b200 as eub
    L = eub.load_library()
    ip = C.POINTER(C.c_int)
    L.eu_debug_box_units.argtypes = [C.c_int]*11 + [ip, C.c_int, ip, C.c_int, ip]
    L.eu_debug_box_units.restype = C.c_int
    nb = C.c_int(0)
    n = L.eu_debug_box_units(nx, ny, tx, ty, z_lo, z_hi, bnd_lo, bnd_hi, blocks, spans, lz, None, 0, None, 0, C.byref(nb))
    if n < 0:
        return None, None
    u = np.zeros((max(n, 1), 4), dtype=np.int32)
    st = np.zeros(nb.value + 1, dtype=np.int32)
    n2 = L.eu_debug_box_units(nx, ny, tx, ty, z_lo, z_hi, bnd_lo, bnd_hi, blocks, spans, lz,
                              u.ctypes.data_as(ip), n, st.ctypes.data_as(ip), len(st), C.byref(nb))
    assert n2 == n
    return u[:n], st


CASES = [
    # nx, ny, tx, ty, z_lo, z_hi, bnd_lo, bnd_hi, blocks
    (512, 512, 32, 8, 0, 256, 0, 0, 592),        # the bench grid, four blocks per SM
    (512, 512, 32, 8, 0, 256, 0, 0, 444),
    (512, 512, 32, 8, 1, 33, 1, 1, 592),         # an inner rank of the 8-GPU run (ghost plane below and above)
    (512, 512, 32, 8, 0, 32, 0, 1, 592),         # the first rank
    (512, 512, 32, 8, 2, 34, 2, 2, 296),         # two boundary planes (capillary term: 2 blocks per SM)
    (100, 100, 50, 5, 0, 100, 0, 0, 296),        # C2
    (256, 256, 32, 8, 0, 128, 0, 0, 296),        # C3
    (256, 256, 32, 8, 4, 126, 4, 4, 296),        # C5 slab: ghost planes to the depth of the largest throw
    (96, 40, 32, 8, 0, 10, 0, 0, 444),
    (6, 5, 6, 5, 0, 4, 0, 0, 444),               # one tile
    (64, 32, 32, 8, 3, 5, 1, 1, 444),            # two own planes, both boundary
    (64, 32, 32, 8, 3, 4, 1, 1, 444),            # one own plane that is both
    (101, 37, 32, 8, 0, 17, 0, 3, 148),          # ragged tiles
]


@pytest.mark.parametrize("spans", [0, 1], ids=["chunks", "spans"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_units_cover_the_own_planes_once(case, spans):
    nx, ny, tx, ty, z_lo, z_hi, bnd_lo, bnd_hi, blocks = case
    u, st = _units(*case, spans=spans)
    assert u is not None
    tiles_x, tiles_y = -(-nx//tx), -(-ny//ty)
    tiles = tiles_x*tiles_y
    # block ranges: monotone, cover the list, no more blocks than asked for
    assert st[0] == 0 and st[-1] == len(u) and np.all(np.diff(st) >= 0) and len(st) - 1 <= blocks
    seen = np.zeros((tiles, z_hi - z_lo), dtype=np.int32)
    for xy, z0, z1, fl in u:
        x0, y0 = int(xy) & 0xffff, int(xy) >> 16
        assert x0 % tx == 0 and y0 % ty == 0 and x0 < nx and y0 < ny and x0 % 2 == 0       # (TMA boxes start at even x)
        assert z_lo <= z0 < z1 <= z_hi
        seen[(y0//ty)*tiles_x + x0//tx, z0 - z_lo:z1 - z_lo] += 1
        # the flags say exactly: holds the first / last own plane, with a neighbour there
        assert bool(fl & 1) == (bnd_lo > 0 and z0 == z_lo)
        assert bool(fl & 2) == (bnd_hi > 0 and z1 == z_hi)
        # the boundary planes of a tile are in ONE unit
        if fl & 1:
            assert z1 - z0 >= bnd_lo
        if fl & 2:
            assert z1 - z0 >= bnd_hi
    assert np.all(seen == 1)
    assert int(np.sum(u[:, 3] & 1 != 0)) == (tiles if bnd_lo else 0)
    assert int(np.sum(u[:, 3] & 2 != 0)) == (tiles if bnd_hi else 0)
    # a block's flagged units come first
    for i in range(len(st) - 1):
        fl = u[st[i]:st[i + 1], 3] != 0
        assert not np.any(fl[1:] & ~fl[:-1])


def test_balance_of_the_two_partitions():
    """chunks: every block gets the same number of units +- 1, equal lengths +- 1 plane; spans: equal plane counts
    within the snapping distance of a cut (2 planes at either end)."""
    u, st = _units(512, 512, 32, 8, 0, 256, 0, 0, 592)
    per = np.diff(st)
    assert per.max() - per.min() <= 1
    ln = u[:, 2] - u[:, 1]
    assert ln.max() - ln.min() <= 1 and ln.max() <= 64
    u, st = _units(512, 512, 32, 8, 0, 256, 0, 0, 592, spans=1)
    planes = np.array([int(np.sum(u[st[i]:st[i + 1], 2] - u[st[i]:st[i + 1], 1])) for i in range(len(st) - 1)])
    assert len(planes) == 592 and planes.max() - planes.min() <= 5
    assert (u[:, 2] - u[:, 1]).min() >= 3


def test_fixed_chunk_length_and_bad_arguments():
    u, st = _units(512, 512, 32, 8, 0, 256, 0, 0, 444, lz=16)
    assert np.all(u[:, 2] - u[:, 1] == 16)
    assert _units(64, 32, 32, 8, 0, 2, 3, 0, 444)[0] is None          # boundary planes do not fit the slab
    assert _units(64, 32, 32, 8, 0, 8, 0, 0, 0)[0] is None
    u, st = _units(64, 32, 32, 8, 5, 5, 0, 0, 444)                    # no own planes
    assert len(u) == 0
