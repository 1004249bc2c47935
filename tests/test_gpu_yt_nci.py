"""GPU parity tests of YT and NCIPLOT through the C ABI (pytest -m gpu).

YT bar: spatial basin ids (0 = IAS point) bit-exact against the oracle's faithful qcksort + sweep on
tie-free data; volumes/populations <= 1e-10 relative; weight fields <= 1e-12 absolute.
NCI bar: RDG <= 1e-12 relative; sign(lambda_2) identical except where |lambda_2| is at rounding level."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
import helpers as H
import systems as S
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")


@pytest.mark.parametrize("name", ["cubic48", "triclinic", "odd_dims", "tiny"])
def test_yt_labels_weights_integrals(ctx, name):
    c = cases.make_case(name)
    n, x2c = c["n"], c["x2c"]
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], x2c, vec, area, atoms=c["atoms"])
    f2 = cases.second_field(c["f"])
    vref, pref = orc.integrate_yt(d, [c["f"], f2], S.omega(x2c))
    h, h2 = ctx.upload(c["f"]), ctx.upload(f2)
    b = ctx.yt_build(h, vec, area)
    mp, na, xa = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
    b.set_map(na, mp)
    assert na == d.nattr
    assert np.array_equal(b.labels(n), d.spatial_basin(n))
    vol, ps = ctx.integrate(b, [h, h2], S.omega(x2c))
    assert np.abs(vol - vref).max() <= 1e-10 * np.abs(vref).max()
    assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
    scale = np.abs(f2).sum() * S.omega(x2c) / f2.size
    assert np.abs(ps[:, 1] - pref[:, 1]).max() <= 1e-10 * scale
    for idb in range(1, na + 1):
        assert np.abs(b.yt_weights(idb, n) - orc.yt_weights(d, idb, n)).max() <= 1e-12
    b.free(); ctx.free(h); ctx.free(h2)


def test_yt_golden(ctx):
    g = json.load(open(GOLDEN))["yt"]
    for name, ref in g.items():
        c = cases.make_case(name)
        vec, area = S.wscell(c["x2c"] / np.array(c["n"], dtype=float)[None, :])
        assert len(area) == ref["nvec"]
        h = ctx.upload(c["f"])
        b = ctx.yt_build(h, vec, area)
        mp, na, _ = H.assign_attractors(b.maxima(), c["n"], c["x2c"], c["atoms"])
        b.set_map(na, mp)
        lab = b.labels(c["n"])
        assert hashlib.sha256(np.ascontiguousarray(lab.ravel(order="F")).tobytes()).hexdigest() == ref["labels_sha256"]
        vol, ps = ctx.integrate(b, [h], S.omega(c["x2c"]))
        assert np.allclose(ps[:, 0], ref["pop"], rtol=1e-10, atol=0)
        assert np.allclose(vol, ref["vol"], rtol=1e-10, atol=0)
        b.free(); ctx.free(h)


def test_yt_chgcar_like_quantised_density(ctx):
    """CHGCAR-like data (12 significant digits, values scaled by the cell volume): ties appear only
    between non-neighbours, so labels and weights must still match the faithful oracle (qcksort order);
    a coarsely quantised variant (6 digits) creates tied neighbours -- mismatches are counted and reported
    against both the qcksort oracle and the stable-order oracle (the order the GPU defines)."""
    c = cases.make_case("cubic48")
    n, x2c = c["n"], c["x2c"]
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    for digits, exact in ((12, True), (6, False)):
        f = S.quantize(c["f"] * S.omega(x2c), digits)
        d = orc.yt_integrate(f, x2c, vec, area, atoms=c["atoms"])
        ds = orc.yt_integrate(f, x2c, vec, area, atoms=c["atoms"], stable=True)
        h = ctx.upload(f)
        b = ctx.yt_build(h, vec, area)
        mp, na, _ = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
        b.set_map(na, mp)
        lab = b.labels(n)
        m_q = int(np.count_nonzero(lab != d.spatial_basin(n)))
        m_s = int(np.count_nonzero(lab != ds.spatial_basin(n)))
        print(f"YT quantised to {digits} digits: label mismatches vs qcksort oracle {m_q}, vs stable oracle {m_s}")
        assert m_s == 0
        if exact:
            assert m_q == 0 and na == d.nattr
            vol, ps = ctx.integrate(b, [h], S.omega(x2c))
            vref, pref = orc.integrate_yt(d, [f], S.omega(x2c))
            assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
        b.free(); ctx.free(h)


def _nci_compare(crho, cgrad, crho_o, cgrad_o, lam2):
    rel = np.abs(cgrad - cgrad_o) / np.maximum(np.abs(cgrad_o), 1e-300)
    # RDG bar: 1e-12 relative.  It holds bit-for-bit / to an ulp wherever the lattice point is exactly a grid
    # node.  Where the coordinate chain lands an ulp below a node (floor() selects the previous cell, fractional
    # offset 1-4e-16) the reference evaluates the full tricubic polynomial through a 64x64 matrix-vector product
    # with coefficients up to +-27: its own rounding noise is ~1e-16 * sum|terms| / |gradient|, i.e. 1e-12..1e-10
    # relative where the gradient is small (and BLAS-order dependent in the real reference).  So:
    # >= 99.9 % of the points within 1e-12, every point within 1e-10 relative or 1e-12 of the median RDG.
    assert (rel <= 1e-12).mean() >= 0.999, f"only {(rel <= 1e-12).mean():.5f} of the points within 1e-12"
    ok = (rel <= 1e-10) | (np.abs(cgrad - cgrad_o) <= 1e-12 * np.median(cgrad_o))
    assert ok.all(), f"max rel RDG error {rel.max():.3e}"
    assert np.abs(np.abs(crho) - np.abs(crho_o)).max() <= 1e-12 * np.abs(crho_o).max()
    flips = np.sign(crho) != np.sign(crho_o)
    # sign(lambda_2) may only differ where lambda_2 is at rounding level
    if flips.any():
        assert (np.abs(lam2[flips]) <= 1e-9 * np.abs(lam2).max()).all()
    return rel.max(), int(flips.sum())


@pytest.mark.parametrize("name", ["triclinic", "odd_dims", "cubic48"])
def test_nci_node_aligned(ctx, name):
    c = cases.make_case(name)
    crho_o, cgrad_o, lam2 = orc.nci_rdg(c["f"], c["x2c"], want_lam2=True)
    h = ctx.upload(c["f"])
    crho, cgrad = ctx.nci_rdg(h, c["x2c"], c["n"])
    r, fl = _nci_compare(crho, cgrad, crho_o, cgrad_o, lam2)
    print(f"NCI {name}: max rel RDG error {r:.2e}, sign flips {fl}")
    ctx.free(h)


def test_nci_general_lattice_and_nucleus_rule(ctx):
    c = cases.make_case("triclinic")
    x2c = c["x2c"]
    nstep = (23, 19, 17)
    x0 = x2c @ np.array([0.013, -0.021, 1.034])      # also exercises the wrap outside [-1e-4, 1+1e-4]
    xmat = x2c / np.array(nstep, dtype=float)[None, :] * 0.93
    nuc = (x2c @ c["atoms"].T).T
    args = dict(nstep=nstep, x0=x0, xmat=xmat, nuclei_cart=nuc)
    crho_o, cgrad_o, lam2 = orc.nci_rdg(c["f"], x2c, want_lam2=True, **args)
    h = ctx.upload(c["f"])
    crho, cgrad = ctx.nci_rdg(h, x2c, c["n"], **args)
    _nci_compare(crho, cgrad, crho_o, cgrad_o, lam2)
    # node-aligned with nuclei on nodes: RDG is exactly zero there on both sides
    crho_o, cgrad_o, lam2 = orc.nci_rdg(c["f"], x2c, nuclei_cart=nuc, want_lam2=True)
    crho, cgrad = ctx.nci_rdg(h, x2c, c["n"], nuclei_cart=nuc)
    _nci_compare(crho, cgrad, crho_o, cgrad_o, lam2)
    assert np.count_nonzero(cgrad == 0.0) >= len(nuc) and np.array_equal(cgrad == 0.0, cgrad_o == 0.0)
    ctx.free(h)


def test_nci_golden_and_resident(ctx):
    ref = json.load(open(GOLDEN))["nci"]["triclinic"]
    c = cases.make_case("triclinic")
    h = ctx.upload(c["f"])
    crho, cgrad = ctx.nci_rdg(h, c["x2c"], c["n"])
    for (k, j, i), r, g in zip(ref["points_kji"], ref["crho"], ref["cgrad"]):
        assert abs(crho[k, j, i] - r) <= 1e-12 * abs(r)
        assert abs(cgrad[k, j, i] - g) <= 1e-12 * abs(g)
    assert abs(cgrad.sum() - ref["sum_cgrad"]) <= 1e-11 * ref["sum_cgrad"]
    hr, hg = ctx.nci_rdg_resident(h, c["x2c"], c["n"])
    assert np.array_equal(ctx.download(hg, cgrad.shape), cgrad)
    for hh in (h, hr, hg):
        ctx.free(hh)


def _close_pairs_case():
    """Two pairs of maxima 1 bohr apart plus two isolated ones, off the nuclei list (NOATOMS): with RATOM 2 the
    reference merges each pair into ONE attractor while it sweeps (are_lclose, yt@proc.f90:142-150)."""
    n = (48, 48, 48)
    x2c = S.cell_x2c(9.0, 9.0, 9.0)
    at = np.array([[0.25, 0.25, 0.25], [0.25 + 1.0 / 9.0, 0.25, 0.25], [0.75, 0.70, 0.30], [0.75, 0.70 + 1.0 / 9.0, 0.30],
                   [0.30, 0.75, 0.75], [0.80, 0.20, 0.80]])
    at = S.snap_to_grid(at, n)
    z = np.array([4.0, 3.5, 6.0, 6.5, 2.0, 5.0]); al = np.array([2.0, 2.1, 2.4, 2.3, 1.6, 1.9])
    f = orc.promolecular(n, x2c, at, z, al, nimg=1)
    return n, x2c, at, f


@pytest.mark.parametrize("mode", ["ratom2", "discard"])
def test_yt_merged_and_discarded_maxima_follow_the_map(ctx, mode):
    """`yt ratom 2` (002_yt_options) and a DISCARDed maximum: the interior / IAS classification must use the basin ids
    the reference's sweep writes into ibasin -- a point between two maxima of one attractor is interior, every point
    below a discarded maximum is an IAS point (yt@proc.f90:129-186)."""
    n, x2c, at, f = _close_pairs_case()
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    h = ctx.upload(f)
    b = ctx.yt_build(h, vec, area)
    assert b.nmax == 6
    if mode == "ratom2":
        d = orc.yt_integrate(f, x2c, vec, area, atoms=None, ratom=2.0, atexist=False)
        mp, na, _ = H.assign_attractors(b.maxima(), n, x2c, None, ratom=2.0, atexist=False)
        assert na == d.nattr == 4
        ref = d.spatial_basin(n)
    else:
        # the oracle has no DISCARD expression: emulate it by merging nothing and blanking basin 2 afterwards is NOT the
        # same thing (the IAS set changes), so compare against the restated rule on the raw oracle data instead:
        # maxima keep their ids, maximum 2 is discarded -> its whole catchment becomes ibasin 0
        d = orc.yt_integrate(f, x2c, vec, area, atoms=None, ratom=1e-90, atexist=False)
        mp = np.arange(1, 7, dtype=np.int32); mp[1] = 0
        na = 6
        ref = orc.yt_reclassify(d, f, vec, mp, n)
    b.set_map(na, mp)
    lab = b.labels(n)
    assert np.array_equal(lab, ref), f"{int(np.count_nonzero(lab != ref))} labels differ"
    if mode == "ratom2":
        vol, ps = ctx.integrate(b, [h], S.omega(x2c))
        vref, pref = orc.integrate_yt(d, [f], S.omega(x2c))
        assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
        assert np.abs(vol - vref).max() <= 1e-10 * np.abs(vref).max()
        for idb in range(1, na + 1):
            assert np.abs(b.yt_weights(idb, n) - orc.yt_weights(d, idb, n)).max() <= 1e-12
        # int_reorder_gridout style relabel afterwards only renames basins: the IAS set must not move
        ias_before = lab == 0
        b.relabel(np.array([1, 1, 2, 3], dtype=np.int32), 3)
        lab2 = b.labels(n)
        assert np.array_equal(lab2 == 0, ias_before)
    b.free(); ctx.free(h)


@pytest.mark.parametrize("name", ["cubic48", "triclinic", "odd_dims"])
def test_yt_export_reproduces_the_ytdata_record(ctx, name):
    """c2g_yt_export against the oracle's ytdata (yt.f90:36-45: nlo, ibasin, iio, inear, fnear as yt_integrate writes them
    to bas%luw, yt@proc.f90:191-199): integers bit-exact, flux fractions bit-exact (same operations, same order)."""
    c = cases.make_case(name)
    n, x2c = c["n"], c["x2c"]
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], x2c, vec, area, atoms=c["atoms"])
    h = ctx.upload(c["f"])
    b = ctx.yt_build(h, vec, area)
    mp, na, _ = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
    b.set_map(na, mp)
    nlo, ibasin, iio, inear, fnear = b.yt_export(n, len(area))
    assert np.array_equal(iio, d.iio)
    assert np.array_equal(ibasin, d.ibasin)
    assert np.array_equal(nlo, d.nlo)
    assert np.array_equal(inear, d.inear)
    assert np.array_equal(fnear, d.fnear)
    # the reference's own yt_weights run on the exported record gives the weights of c2g_yt_weights
    dd = orc.YtData(d.nn, d.nvec)
    dd.nlo, dd.ibasin, dd.iio, dd.inear, dd.fnear, dd.nattr = nlo, ibasin, iio, inear, fnear, na
    for idb in (1, na):
        assert np.abs(orc.yt_weights(dd, idb, n) - b.yt_weights(idb, n)).max() <= 1e-12
    nlo2, ib2, iio2, _, _ = b.yt_export(n, len(area), full=False)
    assert np.array_equal(nlo2, nlo) and np.array_equal(ib2, ibasin) and np.array_equal(iio2, iio)
    b.free(); ctx.free(h)
