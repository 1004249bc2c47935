"""GPU parity tests of the DELOC attractor images (bader_remap, bader@proc.f90:237-296; yt_remap, yt@proc.f90:533-594)
through the C ABI (pytest -m gpu).  Bar: nattn, iatt, ilvec and idg1 bit-exact, in the reference's numbering."""
import numpy as np
import pytest

import cases
import helpers as H
import systems as S
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def cell_of(x2c):
    if np.all(x2c - np.diag(np.diag(x2c)) == 0.0):
        return orc.Cell(x2c), {}
    ws = np.asfortranarray(x2c @ S.wscell(x2c)[0].T.astype(float))
    return orc.Cell(x2c, ws=ws), dict(ws=ws)


@pytest.mark.parametrize("name", ["cubic48", "triclinic", "odd_dims", "ortho_flat", "tiny"])
def test_bader_remap_bit_exact(ctx, name):
    c = cases.make_case(name)
    n, x2c = c["n"], c["x2c"]
    idg, nattr, xattr, _ = orc.bader_integrate(c["f"], x2c, atoms=c["atoms"])
    cell, kw = cell_of(x2c)
    nattn_o, idg1_o, iatt_o, ilvec_o = orc.bader_remap(idg, xattr, cell)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    h = ctx.upload(c["f"])
    b = ctx.bader_assign(h, car2lat, lid)
    mp_, na, _ = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
    b.set_map(na, mp_)
    assert na == nattr
    nattn, idg1, iatt, ilvec = ctx.basins_remap(b, xattr, x2c, shape=n, **kw)
    assert nattn == nattn_o
    assert np.array_equal(iatt, iatt_o) and np.array_equal(ilvec, ilvec_o)
    assert np.array_equal(idg1, idg1_o)
    # a too small capacity reports the size needed
    if nattn > nattr:
        with pytest.raises(capi.C2GError, match="attractor images"):
            ctx.basins_remap(b, xattr, x2c, shape=n, maxattn=nattr, **kw)
    # lists only (no idg1)
    nattn2, none, iatt2, _ = ctx.basins_remap(b, xattr, x2c, want_idg1=False, **kw)
    assert none is None and nattn2 == nattn and np.array_equal(iatt2, iatt)
    b.free(); ctx.free(h)


@pytest.mark.parametrize("name", ["cubic48", "triclinic"])
def test_yt_remap_bit_exact(ctx, name):
    c = cases.make_case(name)
    n, x2c = c["n"], c["x2c"]
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], x2c, vec, area, atoms=c["atoms"])
    cell, kw = cell_of(x2c)
    nattn_o, iatt_o, ilvec_o = orc.yt_remap(d, n, d.xattr, cell)
    h = ctx.upload(c["f"])
    y = ctx.yt_build(h, vec, area)
    mp_, na, _ = H.assign_attractors(y.maxima(), n, x2c, c["atoms"])
    y.set_map(na, mp_)
    nattn, _, iatt, ilvec = ctx.basins_remap(y, d.xattr, x2c, want_idg1=False, **kw)
    assert nattn == nattn_o and np.array_equal(iatt, iatt_o) and np.array_equal(ilvec, ilvec_o)
    with pytest.raises(capi.C2GError, match="no idg1"):
        ctx.basins_remap(y, d.xattr, x2c, shape=n, **kw)
    y.free(); ctx.free(h)
