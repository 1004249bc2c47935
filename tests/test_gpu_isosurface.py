"""GPU parity tests of the ISOSURFACE regions (yt_isosurface, yt@proc.f90:233-390) through the C ABI (pytest -m gpu).

Bar: region ids bit-exact against the oracle (0 below the contour value, merged regions on the smaller id, surviving
regions with their discovery numbers), nraw / nattr equal, plain per-region sums <= 1e-10 relative."""
import numpy as np
import pytest

import cases
import systems as S
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cubic48", "triclinic", "cubic96", "odd_dims", "tiny"])
def test_isosurface_regions_bit_exact(ctx, name):
    c = cases.make_case(name)
    n, x2c, f = c["n"], c["x2c"], c["f"]
    om = S.omega(x2c)
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    h = ctx.upload(f)
    yt = ctx.yt_build(h, vec, area)
    f2 = cases.second_field(f)
    h2 = ctx.upload(f2)
    for q in (0.3, 0.6, 0.8, 0.9, 0.97, 0.995):
        isov = float(np.quantile(f, q))
        idg, nraw, nattr, xattr = orc.yt_isosurface(f, vec, isov)
        b, nraw_g, nattr_g = yt.isosurface(isov)
        assert (nraw_g, nattr_g) == (nraw, nattr), q
        lab = b.labels(n)
        assert np.count_nonzero(lab != idg) == 0, q
        pm = b.maxima()
        assert np.allclose((pm - 1) / np.array(n, dtype=float), xattr.T, atol=0, rtol=0)
        if nraw:
            vol, ps = ctx.integrate(b, [h, h2], om)
            vref, pref = orc.integrate_bader(idg, [f, f2], nraw, om)
            assert np.array_equal(vol, vref)
            assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
            assert np.abs(ps[:, 1] - pref[:, 1]).max() <= 1e-10 * np.abs(f2).sum() * om / f2.size
        b.free()
    yt.free(); ctx.free(h); ctx.free(h2)


def test_isosurface_multipoles_and_empty_set(ctx):
    c = cases.make_case("cubic48")
    n, x2c, f = c["n"], c["x2c"], c["f"]
    om = S.omega(x2c)
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    h = ctx.upload(f)
    yt = ctx.yt_build(h, vec, area)
    isov = float(np.quantile(f, 0.9))
    idg, nraw, nattr, xattr = orc.yt_isosurface(f, vec, isov)
    b, _, _ = yt.isosurface(isov)
    got = ctx.integrate_multipoles(b, h, 3, xattr, x2c, om)
    want = orc.multipoles_bader(idg, xattr, 3, f, orc.Cell(x2c), om)
    rmax = 0.5 * np.linalg.norm(x2c, axis=0).sum()
    scale = np.abs(want[0]).max() * rmax ** np.repeat(np.arange(4), 2 * np.arange(4) + 1)
    assert np.all(np.abs(got - want) <= 1e-10 * scale[:, None])
    b.free()
    b, nraw, nattr = yt.isosurface(float(f.max()) * 2.0)          # nothing above the contour value
    assert (nraw, nattr) == (0, 0) and not b.labels(n).any()
    b.free()
    b, nraw, nattr = yt.isosurface(float(f.min()))                # everything: one periodic region at the end
    idg, nraw_o, nattr_o, _ = orc.yt_isosurface(f, vec, float(f.min()))
    assert (nraw, nattr) == (nraw_o, nattr_o) and np.array_equal(b.labels(n), idg)
    b.free()
    with pytest.raises(capi.C2GError, match="NaN"):
        yt.isosurface(float("nan"))
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    bb = ctx.bader_assign(h, car2lat, lid)
    with pytest.raises(capi.C2GError, match="not a YT result"):
        bb.isosurface(isov)
    bb.free(); yt.free(); ctx.free(h)


def test_isosurface_quantised_density_reported(ctx):
    """CHGCAR-like 12-digit data: ties only between non-neighbours -> still bit-exact against the qcksort oracle."""
    c = cases.make_case("cubic48")
    n, x2c = c["n"], c["x2c"]
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    f = S.quantize(c["f"] * S.omega(x2c), 12)
    h = ctx.upload(f)
    yt = ctx.yt_build(h, vec, area)
    for q in (0.5, 0.9):
        isov = float(np.quantile(f, q))
        idg, nraw, nattr, _ = orc.yt_isosurface(f, vec, isov)
        b, nraw_g, nattr_g = yt.isosurface(isov)
        assert (nraw_g, nattr_g) == (nraw, nattr)
        assert np.array_equal(b.labels(n), idg)
        b.free()
    yt.free(); ctx.free(h)
