"""WCUBE on the device (int_cubew, integration@proc.f90:4428-4466): the weight field of every basin as a resident
grid and its cube value block as text, without a host round trip of the weights (pytest -m gpu)."""
import numpy as np
import pytest

import cases
import helpers as H
import systems as S
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_bader_indicator_cubes(ctx):
    c = cases.make_case("tiny")
    n, x2c = c["n"], c["x2c"]
    c2 = cases.make_case("cubic48")
    for cc in (c, c2):
        n, x2c = cc["n"], cc["x2c"]
        _, car2lat, lid = orc.bader_metrics(x2c, n)
        h = ctx.upload(cc["f"])
        b = ctx.bader_assign(h, car2lat, lid)
        mp, na, _ = H.assign_attractors(b.maxima(), n, x2c, cc["atoms"])
        b.set_map(na, mp)
        idg = b.labels(n)
        tot = np.zeros(n, order="F")
        for i in range(1, na + 1):
            hw = b.weight_grid(i)
            w = ctx.download(hw, n)
            assert np.array_equal(w, (idg == i).astype(np.float64))     # :4455-4458
            tot += w
            if cc is c:   # the cube value block of writegrid_cube, (1p,6(" ",E12.5E3)), byte for byte
                assert ctx.format_text(hw, 1, 12, 5, 1) == orc.format_text_grid(w, 1, 12, 5, 1)
            ctx.free(hw)
        assert np.array_equal(tot, np.ones(n))
        with pytest.raises(capi.C2GError, match="unknown basin"):
            b.weight_grid(na + 1)
        b.free(); ctx.free(h)


def test_yt_weight_cubes_and_isosurface_indicator(ctx):
    c = cases.make_case("cubic48")
    n, x2c = c["n"], c["x2c"]
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], x2c, vec, area, atoms=c["atoms"])
    h = ctx.upload(c["f"])
    y = ctx.yt_build(h, vec, area)
    mp, na, _ = H.assign_attractors(y.maxima(), n, x2c, c["atoms"])
    y.set_map(na, mp)
    tot = np.zeros(n, order="F")
    for i in range(1, na + 1):
        hw = y.weight_grid(i)
        w = ctx.download(hw, n)
        assert np.abs(w - orc.yt_weights(d, i, n)).max() <= 1e-12          # :4451
        assert np.array_equal(w, y.yt_weights(i, n))                          # same kernels as the host-copy call
        tot += w
        ctx.free(hw)
    assert np.abs(tot - 1.0).max() <= 1e-12                                   # partition of unity
    isov = float(np.quantile(c["f"], 0.9))
    reg, nraw, _ = y.isosurface(isov)
    idg = reg.labels(n)
    for i in range(1, nraw + 1):
        hw = reg.weight_grid(i)
        assert np.array_equal(ctx.download(hw, n), (idg == i).astype(np.float64))
        ctx.free(hw)
    reg.free(); y.free(); ctx.free(h)
