"""GPU parity tests of INTEGRABLE ... MULTIPOLES through the C ABI (pytest -m gpu).

Bar: every moment within 1e-10 of the oracle, relative to the sum of |terms| of that moment's basin (the moments of
an atomic basin cancel to near zero for odd l; a tolerance relative to the value itself would test rounding noise)."""
import numpy as np
import pytest

import cases
import helpers as H
import systems as S
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def cell_of(c):
    x2c = c["x2c"]
    if np.all(x2c - np.diag(np.diag(x2c)) == 0.0):
        return orc.Cell(x2c), {}
    vec, _ = S.wscell(x2c)
    ws = np.asfortranarray(x2c @ vec.T.astype(float))
    return orc.Cell(x2c, ws=ws), dict(ws=ws)


def scale_of(idg, f, xattr, cell, lmax, omega):
    """Per (lm, basin): the moments of |f| with |rrlm| bounded by r^l -- cheap upper bound via the oracle on |f|."""
    mp = np.abs(orc.multipoles_bader(idg, xattr, 0, np.abs(f), cell, omega))  # sum |f| per basin
    rmax = 0.5 * np.linalg.norm(cell.x2c, axis=0).sum()  # a shortest vector lies inside the WS cell
    return np.array([[mp[0, b] * max(1.0, rmax) ** l for b in range(mp.shape[1])] for l in range(lmax + 1) for _ in range(2 * l + 1)])


@pytest.mark.parametrize("name,lmax", [("cubic48", 5), ("triclinic", 5), ("odd_dims", 3), ("tiny", 0), ("ortho_flat", 7), ("cubic48", 10)])
def test_bader_multipoles(ctx, name, lmax):
    c = cases.make_case(name)
    n, x2c, om = c["n"], c["x2c"], S.omega(c["x2c"])
    idg, nattr, xattr, _ = orc.bader_integrate(c["f"], x2c, atoms=c["atoms"])
    cell, kw = cell_of(c)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    h = ctx.upload(c["f"])
    b = ctx.bader_assign(h, car2lat, lid)
    mp_, na, xa = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
    b.set_map(na, mp_)
    assert na == nattr and np.array_equal(b.labels(n), idg)
    f2 = cases.second_field(c["f"])
    h2 = ctx.upload(f2)
    for fh, f in ((h, c["f"]), (h2, f2)):
        got = ctx.integrate_multipoles(b, fh, lmax, xattr, x2c, om, **kw)
        want = orc.multipoles_bader(idg, xattr, lmax, f, cell, om)
        sc = scale_of(idg, f, xattr, cell, lmax, om)
        assert got.shape == want.shape
        assert np.all(np.abs(got - want) <= TOL * sc), np.abs((got - want) / sc).max()
    # the monopole is the basin population
    _, ps = ctx.integrate(b, [h], om)
    got = ctx.integrate_multipoles(b, h, 0, xattr, x2c, om, **kw)
    assert np.abs(got[0] - ps[:, 0]).max() <= TOL * np.abs(ps[:, 0]).max()
    b.free(); ctx.free(h); ctx.free(h2)


def test_bader_multipoles_discarded_basin_and_merged_maxima(ctx):
    """map(m) = 0 leaves a basin out; two maxima mapped onto one attractor are summed about that attractor."""
    c = cases.make_case("cubic48")
    n, x2c, om = c["n"], c["x2c"], S.omega(c["x2c"])
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    h = ctx.upload(c["f"])
    b = ctx.bader_assign(h, car2lat, lid)
    mp_, na, xa = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
    new = np.array([0 if v == 2 else (1 if v in (1, 3) else v - 2) for v in mp_], dtype=np.int32)
    nattr = int(new.max())
    b.set_map(nattr, new)
    idg = b.labels(n)
    xattr = np.asfortranarray(np.array([xa[:, 0]] + [xa[:, k] for k in range(3, xa.shape[1])]).T)
    got = ctx.integrate_multipoles(b, h, 4, xattr, x2c, om)
    want = orc.multipoles_bader(idg, xattr, 4, c["f"], orc.Cell(x2c), om)
    sc = scale_of(idg, c["f"], xattr, orc.Cell(x2c), 4, om)
    assert np.all(np.abs(got - want) <= TOL * sc)
    b.free(); ctx.free(h)


@pytest.mark.parametrize("name", ["cubic48", "triclinic"])
def test_yt_multipoles(ctx, name):
    c = cases.make_case(name)
    n, x2c, om = c["n"], c["x2c"], S.omega(c["x2c"])
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], x2c, vec, area, atoms=c["atoms"])
    cell, kw = cell_of(c)
    h = ctx.upload(c["f"])
    b = ctx.yt_build(h, vec, area)
    mp_, na, xa = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
    b.set_map(na, mp_)
    assert na == d.nattr
    lmax = 4
    domask = np.ones(na, dtype=np.uint8)
    domask[1] = 0                                              # docelatom(icp(2)) = .false.: the basin is skipped
    got = ctx.integrate_multipoles(b, h, lmax, d.xattr, x2c, om, domask=domask, **kw)
    ones = np.ones(n, dtype=np.int32, order="F")
    for m in range(1, na + 1):
        if not domask[m - 1]:
            assert np.all(got[:, m - 1] == 0.0)
            continue
        w = orc.yt_weights(d, m, n)
        want = orc.multipoles_weighted(w, d.xattr[:, m - 1], lmax, c["f"], cell, om)
        sc = scale_of(ones, c["f"] * w, d.xattr[:, m - 1:m], cell, lmax, om)[:, 0]
        assert np.all(np.abs(got[:, m - 1] - want) <= TOL * sc), (m, np.abs((got[:, m - 1] - want) / sc).max())
    b.free(); ctx.free(h)


def test_multipoles_error_paths(ctx):
    c = cases.make_case("tiny")
    n, x2c = c["n"], c["x2c"]
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    h = ctx.upload(c["f"])
    b = ctx.bader_assign(h, car2lat, lid)
    with pytest.raises(capi.C2GError, match="set_map"):
        b.nattr = 1
        ctx.integrate_multipoles(b, h, 2, np.zeros((3, 1)), x2c, 1.0)
    b.set_map(1, np.ones(b.nmax, dtype=np.int32))
    with pytest.raises(capi.C2GError, match="lmax"):
        ctx.integrate_multipoles(b, h, 11, np.zeros((3, 1)), x2c, 1.0)
    with pytest.raises(capi.C2GError, match="field handle"):
        ctx.integrate_multipoles(b, 999, 2, np.zeros((3, 1)), x2c, 1.0)
    b.free(); ctx.free(h)
