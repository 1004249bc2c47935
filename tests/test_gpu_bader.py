"""GPU parity tests of BADER + INTEGRABLE through the C ABI (run on the B200 box: pytest -m gpu).

Bar: basin labels and attractor counts bit-exact against the oracle's faithful sequential
restatement of bader_integrate; integrated volumes exact, populations <= 1e-10 relative."""
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


def gpu_bader(ctx, c, algo, order=capi.ORDER_INDEX, atexist=True):
    _, car2lat, lid = orc.bader_metrics(c["x2c"], c["n"])
    h = ctx.upload(c["f"])
    b = ctx.bader_assign(h, car2lat, lid, algo=algo, order=order)
    mp, na, xa = H.assign_attractors(b.maxima(), c["n"], c["x2c"], c["atoms"] if atexist else None, atexist=atexist)
    b.set_map(na, mp)
    return h, b, na


@pytest.mark.parametrize("name", [c[0] for c in cases.SMALL_CASES])
@pytest.mark.parametrize("algo", [capi.BADER_EXACT, capi.BADER_FAST])
def test_labels_bit_exact_vs_reference_scan(ctx, name, algo):
    c = cases.make_case(name)
    idg, nattr, xattr, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    h, b, na = gpu_bader(ctx, c, algo)
    lab = b.labels(c["n"])
    assert na == nattr
    assert b.nmax == nattr
    assert np.count_nonzero(lab != idg) == 0
    # INTEGRABLE: volume, rho, second field
    f2 = cases.second_field(c["f"])
    h2 = ctx.upload(f2)
    vol, ps = ctx.integrate(b, [h, h2], S.omega(c["x2c"]))
    vref, pref = orc.integrate_bader(idg, [c["f"], f2], nattr, S.omega(c["x2c"]))
    assert np.array_equal(vol, vref)                                   # integer counts: exact
    assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
    scale = np.abs(f2).sum() * S.omega(c["x2c"]) / f2.size             # Laplacian-like: integrals ~ 0
    assert np.abs(ps[:, 1] - pref[:, 1]).max() <= 1e-10 * scale
    assert int(b.counts().sum()) == c["f"].size
    b.free(); ctx.free(h); ctx.free(h2)


def test_degenerate_thin_cell_reported(ctx):
    """Scan-order dependent residue of the sequential reference (see tests/cases.py): bounded and reported,
    not bit-exact.  The exact-walk referee must still equal the oracle's own-trajectory labelling."""
    c = cases.make_case("thin_cell")
    idg, nattr, _, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    term, _ = orc.bader_canonical(c["f"], c["x2c"])
    for algo in (capi.BADER_EXACT, capi.BADER_FAST):
        h, b, na = gpu_bader(ctx, c, algo)
        lab = b.labels(c["n"])
        mism = int(np.count_nonzero(lab != idg))
        print(f"thin_cell algo={algo}: {mism} of {lab.size} labels differ from the sequential reference")
        assert na == nattr and mism <= 1e-4 * lab.size
        if algo == capi.BADER_EXACT:
            pm = b.maxima() - 1
            lin = pm[:, 0] + c["n"][0] * (pm[:, 1] + c["n"][1] * pm[:, 2])
            t2 = np.searchsorted(np.sort(lin), term)                      # terminal -> index in the sorted maxima list
            b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
            assert np.array_equal(b.labels(c["n"]), (t2 + 1).astype(np.int32))
        b.free(); ctx.free(h)


def test_golden_labels(ctx):
    g = json.load(open(GOLDEN))["bader"]
    for name, ref in g.items():
        c = cases.make_case(name)
        h, b, na = gpu_bader(ctx, c, capi.BADER_FAST)
        lab = b.labels(c["n"])
        assert na == ref["nattr"]
        assert hashlib.sha256(np.ascontiguousarray(lab.ravel(order="F")).tobytes()).hexdigest() == ref["labels_sha256"]
        vol, ps = ctx.integrate(b, [h], S.omega(c["x2c"]))
        assert np.allclose(ps[:, 0], ref["pop"], rtol=1e-10, atol=0)
        assert np.allclose(vol, ref["vol"], rtol=1e-13, atol=0)
        b.free(); ctx.free(h)


def test_noatoms_mode_partition_and_scan_order(ctx):
    """NOATOMS: numbering follows discovery order in the reference; the GPU returns maxima ordered by
    the scan key of the first point of each basin.  The partition must be identical; the numbering is
    compared modulo a permutation and the permutation is reported."""
    c = cases.make_case("cubic96")
    idg, nattr, xattr, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=None, atexist=False)
    h, b, na = gpu_bader(ctx, c, capi.BADER_FAST, order=capi.ORDER_SCAN, atexist=False)
    lab = b.labels(c["n"])
    assert na == nattr
    pairs = set(zip(idg.ravel().tolist(), lab.ravel().tolist()))
    assert len(pairs) == nattr  # a bijection between the two numberings
    same = sum(1 for a, bb in pairs if a == bb)
    print(f"NOATOMS numbering: {same}/{nattr} attractors keep the reference's id")
    b.free(); ctx.free(h)


def test_relabel_composes_maps(ctx):
    c = cases.make_case("cubic48")
    h, b, na = gpu_bader(ctx, c, capi.BADER_FAST)
    lab = b.labels(c["n"])
    assigned = np.array([2, 2, 1, 3, 3, 1], dtype=np.int32)  # int_reorder_gridout style merge
    b.relabel(assigned, 3)
    lab2 = b.labels(c["n"])
    assert np.array_equal(lab2, assigned[lab - 1])
    vol, ps = ctx.integrate(b, [h], S.omega(c["x2c"]))
    assert abs(vol.sum() - S.omega(c["x2c"])) < 1e-9
    b.free(); ctx.free(h)


def test_discarded_attractor_leaves_points_unassigned(ctx):
    c = cases.make_case("cubic48")
    _, car2lat, lid = orc.bader_metrics(c["x2c"], c["n"])
    h = ctx.upload(c["f"])
    b = ctx.bader_assign(h, car2lat, lid)
    mp = np.arange(1, b.nmax + 1, dtype=np.int32)
    mp[0] = 0
    b.set_map(b.nmax, mp)
    lab = b.labels(c["n"])
    cnt = b.counts()
    assert np.count_nonzero(lab == 0) == cnt[0]
    b.free(); ctx.free(h)


def test_rough_field_reported_not_asserted(ctx):
    """30 % multiplicative noise (hundreds of noise basins).  On such fields the discrete near-grid map has
    isolated points whose own trajectory ends elsewhere than all of their neighbours'; the sequential reference
    keeps for them whatever an earlier path left behind (scan-order dependent, SURVEY.md appendix D: 1e-5..1e-4
    of the points), the hierarchical algorithm keeps the fill value, the exact referee their own terminal.
    All three partitions must agree except for such residues, which are counted and printed."""
    c = cases.make_case("cubic48")
    rng = np.random.default_rng(77)
    f = np.asfortranarray(c["f"] * (1.0 + 0.3 * (rng.random(c["n"]) - 0.5)))
    c2 = dict(c, f=f)
    h1, b1, _ = gpu_bader(ctx, c2, capi.BADER_EXACT, atexist=False)
    h2, b2, _ = gpu_bader(ctx, c2, capi.BADER_FAST, atexist=False)
    term, _ = orc.bader_canonical(f, c["x2c"])
    b1.set_map(b1.nmax, np.arange(1, b1.nmax + 1, dtype=np.int32))
    pm = b1.maxima() - 1
    lin = pm[:, 0] + c["n"][0] * (pm[:, 1] + c["n"][1] * pm[:, 2])
    # the exact referee equals the oracle's own-trajectory labelling on every point, rough field or not
    assert np.array_equal(b1.labels(c["n"]), (np.searchsorted(lin, term) + 1).astype(np.int32))
    idg, nattr, _, _ = orc.bader_integrate(f, c["x2c"], atoms=None, atexist=False)

    def partition_mismatch(a, b):
        pairs = {}
        for u, v in zip(a.ravel().tolist(), b.ravel().tolist()):
            pairs.setdefault(u, {}).setdefault(v, 0)
            pairs[u][v] += 1
        return sum(sum(v.values()) - max(v.values()) for v in pairs.values())

    b2.set_map(b2.nmax, np.arange(1, b2.nmax + 1, dtype=np.int32))
    l1, l2 = b1.labels(c["n"]), b2.labels(c["n"])
    m_exact, m_fast, m_ef = partition_mismatch(idg, l1), partition_mismatch(idg, l2), partition_mismatch(l1, l2)
    print(f"rough field: attractors ref={nattr} exact={b1.nmax} fast={b2.nmax}; points differing from the sequential "
          f"reference: exact {m_exact}, fast {m_fast}; fast vs exact {m_ef} (of {f.size})")
    assert max(m_exact, m_fast, m_ef) <= 2e-3 * f.size
    assert abs(b1.nmax - nattr) <= 0.02 * nattr + 2 and abs(b2.nmax - nattr) <= 0.02 * nattr + 2
    for b, hh in ((b1, h1), (b2, h2)):
        b.free(); ctx.free(hh)


@pytest.mark.parametrize("N,side", [(256, 4)])
def test_fast_equals_exact_at_config_size(ctx, N, side):
    """configs[1]-sized grid (256^3, GPU-generated promolecular-like density): the hierarchical
    algorithm must reproduce the exact-walk referee on every point; counts sum to N^3; every maximum
    is labelled with itself."""
    n = (N, N, N)
    x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
    at, z, al = S.jittered_lattice(side, 5)
    at = S.snap_to_grid(at, n)
    h = ctx.alloc(n)
    ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    labs = []
    for algo in (capi.BADER_EXACT, capi.BADER_FAST):
        b = ctx.bader_assign(h, car2lat, lid, algo=algo)
        assert b.nmax == side**3
        b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
        lab = b.labels(n)
        pm = b.maxima() - 1
        assert np.array_equal(lab[pm[:, 0], pm[:, 1], pm[:, 2]], np.arange(1, b.nmax + 1))
        assert int(b.counts().sum()) == N**3
        vol, ps = ctx.integrate(b, [h], S.omega(x2c))
        assert abs(vol.sum() - S.omega(x2c)) < 1e-9 * S.omega(x2c)
        labs.append(lab)
        b.free()
    assert np.array_equal(labs[0], labs[1])
    ctx.free(h)


def test_headline_size_properties(ctx):
    """BASELINE.json configs[4] at full size (1024^3, 512 atoms, generated in HBM): size-independent properties of
    the assignment + integration through the C ABI -- every point belongs to exactly one basin (counts sum to N^3,
    volumes sum to the cell volume), the basin populations add up to the integral over the whole cell obtained with
    all maxima mapped to one basin (1e-10 relative, north_star), one basin per atom, and the result is reproducible
    (second call: identical counts; populations to 1e-13, the order of the shared-memory atomics inside a block is free)."""
    N, side = 1024, 8
    n = (N, N, N)
    x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
    at, z, al = S.jittered_lattice(side, 5)
    at = S.snap_to_grid(at, n)
    h = ctx.alloc(n)
    ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    om = S.omega(x2c)
    runs = []
    for rep in range(2):
        b = ctx.bader_assign(h, car2lat, lid, algo=capi.BADER_FAST)
        assert b.nmax == side**3
        cnt = b.counts()
        assert int(cnt.sum()) == N**3 and cnt.min() > 0
        b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
        vol, ps = ctx.integrate(b, [h], om)
        assert abs(vol.sum() - om) <= 1e-12 * om
        b.set_map(1, np.ones(b.nmax, dtype=np.int32))
        vol1, ps1 = ctx.integrate(b, [h], om)
        assert abs(ps[:, 0].sum() - ps1[0, 0]) <= 1e-10 * abs(ps1[0, 0])
        # every maximum sits on an atom (the density model has one cusp per snapped atom)
        pm = (b.maxima() - 1) / np.array(n, dtype=float)
        d = np.abs(pm[:, None, :] - at[None, :, :])
        d = np.minimum(d, 1.0 - d).max(axis=2).min(axis=1)
        assert d.max() <= 1.5 / N
        runs.append((cnt.copy(), ps.copy()))
        if rep == 1:
            # INTEGRABLE ... MULTIPOLES at the headline size: the monopole of every basin is its population, the
            # first moments about the attractor are small against (basin radius) x population (near-spherical atoms)
            b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
            mpole = ctx.integrate_multipoles(b, h, 5, pm.T, x2c, om)
            assert mpole.shape == (36, b.nmax)
            assert np.abs(mpole[0] - ps[:, 0]).max() <= 1e-10 * np.abs(ps[:, 0]).max()
            assert np.all(np.isfinite(mpole)) and np.abs(mpole[1:4]).max() <= 2.5 * np.abs(ps[:, 0]).max()
        b.free()
    assert np.array_equal(runs[0][0], runs[1][0])
    assert np.abs(runs[0][1] - runs[1][1]).max() <= 1e-13 * np.abs(runs[0][1]).max()
    ctx.free(h)


def test_error_paths(ctx):
    with pytest.raises(capi.C2GError):
        ctx.bader_assign(12345, np.eye(3), np.ones(27))
    c = cases.make_case("tiny")
    _, car2lat, lid = orc.bader_metrics(c["x2c"], c["n"])
    h = ctx.upload(c["f"])
    b = ctx.bader_assign(h, car2lat, lid)
    with pytest.raises(capi.C2GError):
        b.labels(c["n"])  # no map yet
    with pytest.raises(capi.C2GError):
        b.set_map(1, np.array([5], dtype=np.int32))
    b.free(); ctx.free(h)


def test_async_upload_and_labels_match_the_synchronous_calls(ctx):
    """c2g_grid_upload_async / c2g_basins_labels_async (the e2e path of bench.py): same labels and sums as the
    synchronous entry points, with the second field uploaded while the assignment runs."""
    import torch
    c = cases.make_case("odd_dims")
    n = c["n"]
    f2 = cases.second_field(c["f"])
    _, car2lat, lid = orc.bader_metrics(c["x2c"], n)
    # synchronous reference
    h, b, na = gpu_bader(ctx, c, capi.BADER_FAST)
    h2 = ctx.upload(f2)
    lab_ref = b.labels(n)
    vol_ref, ps_ref = ctx.integrate(b, [h, h2], S.omega(c["x2c"]))
    b.free(); ctx.free(h); ctx.free(h2)
    # asynchronous path from pinned host buffers
    nn = int(np.prod(n))
    pa = torch.empty(nn, dtype=torch.float64, pin_memory=True)
    pb = torch.empty(nn, dtype=torch.float64, pin_memory=True)
    pl = torch.empty(nn, dtype=torch.int32, pin_memory=True)
    pa.numpy()[:] = c["f"].ravel(order="F")
    pb.numpy()[:] = f2.ravel(order="F")
    for rep in range(2):  # the second pass reuses cached device blocks
        ha = ctx.upload_ptr_async(pa.data_ptr(), n)
        hb = ctx.upload_ptr_async(pb.data_ptr(), n)
        bb = ctx.bader_assign(ha, car2lat, lid)
        mp2, na2, _ = H.assign_attractors(bb.maxima(), n, c["x2c"], c["atoms"])
        bb.set_map(na2, mp2)
        pl.zero_()
        bb.labels_ptr_async(pl.data_ptr())
        vol, ps = ctx.integrate(bb, [ha, hb], S.omega(c["x2c"]))
        ctx.synchronize()
        lab = pl.numpy().reshape(n, order="F")
        assert np.array_equal(lab, lab_ref)
        assert np.array_equal(vol, vol_ref)
        assert np.abs(ps - ps_ref).max() <= 1e-12 * np.abs(ps_ref).max()
        bb.free(); ctx.free(ha); ctx.free(hb)


@pytest.mark.parametrize("env", [{"C2G_SYNC_LEVELS": "1"}, {"C2G_WALK3": "0"}, {"C2G_WALK3": "0", "C2G_WALK2": "1"},
                                 {"C2G_WALK3": "5"}, {"C2G_CERT": "2"}, {"C2G_NO_EARLY_STOP": "1"}, {"C2G_FILL_OLD": "1"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
@pytest.mark.parametrize("name", ["cubic48", "odd_dims", "triclinic"])
def test_fallback_flows_and_kernels_give_the_same_labels(ctx, name, env):
    """The switches of DESIGN.md 9b select older kernels and flows that stay in the library for measurements (host-driven
    level loop, k_walk, k_walk2, mixed kernels, cube certificates, no early stop, the full-stencil fill pass): every one
    of them must still produce the oracle's labels."""
    c = cases.make_case(name)
    idg, nattr, _, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    os.environ.update(env)
    try:
        h, b, na = gpu_bader(ctx, c, capi.BADER_FAST)
    finally:
        for k in env:
            del os.environ[k]
    assert na == nattr and np.count_nonzero(b.labels(c["n"]) != idg) == 0
    b.free(); ctx.free(h)
