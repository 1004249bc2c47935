"""Oracle checks of the yt_isosurface restatement (yt@proc.f90:233-390)."""
import numpy as np

import cases
import systems as S
from oracle import oracle as orc


def components(mask, vec):
    """Connected components of a periodic boolean grid under the stencil `vec` (union-find, numpy)."""
    n = mask.shape
    idx = np.arange(mask.size).reshape(n, order="F")
    parent = np.arange(mask.size)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for v in vec:
        nb = np.roll(idx, shift=(-v[0], -v[1], -v[2]), axis=(0, 1, 2))
        both = mask & np.roll(mask, shift=(-v[0], -v[1], -v[2]), axis=(0, 1, 2))
        for a, b in zip(idx[both].tolist(), nb[both].tolist()):
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    roots = np.array([find(a) for a in range(mask.size)]).reshape(n, order="F")
    return np.where(mask, roots, -1)


def test_isolated_regions_are_the_connected_components():
    """Above the saddle values the regions never touch: ids 1..nraw in decreasing order of the regional maxima,
    one connected component each, nothing merged."""
    c = cases.make_case("cubic48")
    n, x2c, f = c["n"], c["x2c"], c["f"]
    vec, _ = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    isov = np.quantile(f, 0.99)
    idg, nraw, nattr, xattr = orc.yt_isosurface(f, vec, isov)
    assert np.array_equal(idg > 0, f >= isov)
    comp = components(f >= isov, vec)
    assert nraw == nattr == len(np.unique(comp[comp >= 0]))
    for r in range(1, nraw + 1):   # one component per region; its maximum is the region's attractor
        assert len(np.unique(comp[idg == r])) == 1
        p = np.unravel_index(np.argmax(np.where(idg == r, f, -1.0)), n)
        assert np.allclose(np.array(p) / np.array(n), xattr[:, r - 1])
    peak = [f[tuple(np.round(xattr[:, r] * n).astype(int))] for r in range(nraw)]
    assert np.all(np.diff(peak) <= 0)


def test_merged_regions_refine_the_connected_components():
    """Below the saddles regions merge.  The reference's imap bookkeeping can LOSE merges (a later contact overwrites
    imap(b)), so its regions are unions of ascent domains that never straddle two connected components, and there
    are at least as many of them as components."""
    c = cases.make_case("cubic96")
    n, x2c, f = c["n"], c["x2c"], c["f"]
    vec, _ = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    for q in (0.5, 0.8, 0.9):
        isov = np.quantile(f, q)
        idg, nraw, nattr, _ = orc.yt_isosurface(f, vec, isov)
        comp = components(f >= isov, vec)
        ids = np.unique(idg[idg > 0])
        assert len(ids) == nattr <= nraw and ids.max() <= nraw
        assert nattr >= len(np.unique(comp[comp >= 0]))
        for r in ids:
            assert len(np.unique(comp[idg == r])) == 1
        # stable order == qcksort order on tie-free data
        idg2, _, nattr2, _ = orc.yt_isosurface(f, vec, isov, stable=True)
        assert nattr2 == nattr and np.array_equal(idg, idg2)


def test_contour_above_the_maximum_gives_no_region():
    c = cases.make_case("tiny")
    vec, _ = S.wscell(c["x2c"] / np.array(c["n"], dtype=float)[None, :])
    idg, nraw, nattr, _ = orc.yt_isosurface(c["f"], vec, c["f"].max() * 2)
    assert nraw == 0 and nattr == 0 and not idg.any()
