"""Oracle checks of bader_remap / yt_remap (bader@proc.f90:237-296, yt@proc.f90:533-594)."""
import numpy as np

import cases
import systems as S
from oracle import oracle as orc


def cell_of(x2c):
    if np.all(x2c - np.diag(np.diag(x2c)) == 0.0):
        return orc.Cell(x2c)
    return orc.Cell(x2c, ws=x2c @ S.wscell(x2c)[0].T.astype(float))


def test_bader_remap_images_are_consistent():
    for name in ("cubic48", "triclinic"):
        c = cases.make_case(name)
        n, x2c = np.array(c["n"]), c["x2c"]
        idg, nattr, xattr, _ = orc.bader_integrate(c["f"], x2c, atoms=c["atoms"])
        nattn, idg1, iatt, ilvec = orc.bader_remap(idg, xattr, cell_of(x2c))
        assert np.array_equal(iatt[:nattr], np.arange(1, nattr + 1)) and not ilvec[:, :nattr].any()
        assert np.array_equal(iatt[idg1 - 1], idg)                      # an image belongs to the point's own basin
        assert len({(int(a), *map(int, v)) for a, v in zip(iatt, ilvec.T)}) == nattn   # no image twice
        first = [np.flatnonzero(idg1.ravel(order="F") == k)[0] for k in range(nattr + 1, nattn + 1)]
        assert np.all(np.diff(first) > 0)                               # numbered by first appearance in the scan
        # the image vector brings the point next to its attractor: |p/n - xattr - ilvec| is the shortest distance
        rng = np.random.default_rng(0)
        for q in rng.integers(0, idg.size, 200):
            p = np.array(np.unravel_index(q, idg.shape, order="F"))
            k = idg1[tuple(p)] - 1
            d = x2c @ (p / n - xattr[:, iatt[k] - 1] - ilvec[:, k])
            best = min(np.linalg.norm(x2c @ (p / n - xattr[:, iatt[k] - 1] + np.array([a, b, cc])))
                       for a in range(-2, 3) for b in range(-2, 3) for cc in range(-2, 3))
            assert abs(np.linalg.norm(d) - best) <= 1e-12


def test_yt_remap_lists_every_image_of_the_weighted_points():
    c = cases.make_case("cubic48")
    n, x2c = c["n"], c["x2c"]
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], x2c, vec, area, atoms=c["atoms"])
    nattn, iatt, ilvec = orc.yt_remap(d, n, d.xattr, cell_of(x2c))
    assert nattn > d.nattr and np.all(np.diff(iatt[d.nattr:]) >= 0)      # basin-outer loop: images grouped by basin
    # the YT images contain the Bader-like images of the interior points
    nb, _, ia_b, il_b = orc.bader_remap(d.spatial_basin(n), d.xattr, cell_of(x2c))
    yt_set = {(int(a), *map(int, v)) for a, v in zip(iatt, ilvec.T)}
    assert {(int(a), *map(int, v)) for a, v in zip(ia_b, il_b.T)} <= yt_set
