"""CPU tests of the oracle's YT / tricubic / NCI restatements -- no GPU."""
import os
import re

import numpy as np
import pytest

import cases
import systems as S
from oracle import oracle as orc

REF = "/root/reference/src"


def test_qcksort_sorts_and_matches_stable_on_tie_free_data():
    rng = np.random.default_rng(5)
    a = rng.random(20011)
    io = orc.qcksort(a)
    assert sorted(io.tolist()) == list(range(1, a.size + 1))
    assert (np.diff(a[io - 1]) >= 0).all()
    assert (io - 1 == np.argsort(a, kind="stable")).all()


def test_qcksort_with_ties_is_a_sorted_permutation():
    rng = np.random.default_rng(6)
    a = np.round(rng.random(5000), 2)
    io = orc.qcksort(a)
    assert sorted(io.tolist()) == list(range(1, a.size + 1))
    assert (np.diff(a[io - 1]) >= 0).all()


@pytest.mark.parametrize("cellp,n", [((7, 7, 7, 90, 90, 90), (24, 24, 24)), ((8, 8.5, 7.5, 80, 95, 105), (24, 26, 22))])
def test_yt_weights_partition_of_unity(cellp, n):
    x2c = S.cell_x2c(*cellp)
    at, z, al = S.random_atoms(3, 31, x2c)
    at = S.snap_to_grid(at, n)
    f = orc.promolecular(n, x2c, at, z, al, nimg=1)
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    assert len(area) in (6, 8, 12, 14)
    d = orc.yt_integrate(f, x2c, vec, area, atoms=at)
    d2 = orc.yt_integrate(f, x2c, vec, area, atoms=at, stable=True)
    assert (d.spatial_basin(n) == d2.spatial_basin(n)).all()
    assert d.nattr == 3
    wsum = sum(orc.yt_weights(d, i + 1, n) for i in range(d.nattr))
    assert np.abs(wsum - 1.0).max() < 1e-12
    vol, ps = orc.integrate_yt(d, [f], S.omega(x2c))
    assert abs(vol.sum() - S.omega(x2c)) < 1e-9 * S.omega(x2c)
    assert abs(ps[:, 0].sum() - f.sum() * S.omega(x2c) / f.size) < 1e-10 * abs(ps[:, 0].sum())


def test_wscell_cubic_and_fcc():
    vec, area = S.wscell(np.eye(3) * 0.2)
    assert len(vec) == 6 and np.allclose(area, 0.04)
    fcc = 0.5 * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float).T
    vec, area = S.wscell(fcc)
    assert len(vec) == 12  # rhombic dodecahedron
    assert np.allclose(area, area[0])


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present on this machine")
def test_tricubic_matrix_equals_reference_table():
    """The regenerated Lekien-Marsden matrix equals the table in grid3mod@proc.f90:76-340."""
    src = open(os.path.join(REF, "grid3mod@proc.f90")).read()
    start = src.index("real*8, parameter :: c(64,64) = reshape((/")
    end = src.index("/),shape(c))", start)
    body = re.sub(r"!.*", "", src[start:end].split("reshape((/", 1)[1])
    vals = [float(v.replace("d0", "")) for v in re.findall(r"-?\d+d0", body)]
    assert len(vals) == 4096
    cref = np.array(vals).reshape((64, 64), order="F")
    assert (cref == orc.tricubic_matrix()).all()


def test_tricubic_interpolates_nodes_and_is_catmull_rom():
    rng = np.random.default_rng(3)
    n = (7, 8, 9)
    f = np.asfortranarray(rng.random(n))
    c2x = np.eye(3)
    for idx in [(0, 0, 0), (3, 4, 5), (6, 7, 8)]:
        xi = np.array(idx) / np.array(n)
        y, yp, ypp = orc.grid_interp_tricubic(f, c2x, xi)
        assert y == f[idx]
        cd = 0.5 * (f[(idx[0] + 1) % 7, idx[1], idx[2]] - f[(idx[0] - 1) % 7, idx[1], idx[2]]) * 7
        assert yp[0] == cd
    # off-node: tensor-product Catmull-Rom weights
    def w(u):
        return 0.5 * np.array([-u**3 + 2 * u**2 - u, 3 * u**3 - 5 * u**2 + 2, -3 * u**3 + 4 * u**2 + u, u**3 - u**2])
    xi = np.array([0.31, 0.47, 0.83])
    y, _, _ = orc.grid_interp_tricubic(f, c2x, xi)
    t = xi * np.array(n)
    i0 = np.floor(t).astype(int)
    u = t - i0
    acc = 0.0
    for a in range(4):
        for b in range(4):
            for c in range(4):
                acc += w(u[0])[a] * w(u[1])[b] * w(u[2])[c] * f[(i0[0] + a - 1) % 7, (i0[1] + b - 1) % 8, (i0[2] + c - 1) % 9]
    assert abs(acc - y) < 1e-13


def test_nci_rdg_exponential_density_known_answer():
    """rho = exp(-a x) along one axis on a fine grid: s = a rho^(-1/3)/(2 (3 pi^2)^(1/3)) analytically;
    the node-aligned tricubic path uses central differences, so compare with those."""
    n = (64, 8, 8)
    L = 8.0
    x2c = np.diag([L, 2.0, 2.0])
    xs = np.arange(n[0]) / n[0]
    rho1 = 2.0 + np.cos(2 * np.pi * xs)
    f = np.asfortranarray(np.broadcast_to(rho1[:, None, None], n).copy())
    crho, cgrad = orc.nci_rdg(f, x2c)
    cst = 2.0 * (3.0 * np.pi**2) ** (1.0 / 3.0)
    g = 0.5 * (np.roll(rho1, -1) - np.roll(rho1, 1)) * n[0] / L
    s_expect = np.abs(g) / (cst * rho1 ** (4.0 / 3.0))
    got = cgrad[0, 0, :]  # (k,j,i)
    assert np.allclose(got, s_expect, rtol=1e-13, atol=1e-15)
    assert np.allclose(np.abs(crho[0, 0, :]), rho1 * 100.0, rtol=1e-15)
    # lambda_2 = 0 exactly here (1-D field): sign(rho, +0) = +
    assert (crho > 0).all()


def test_yt_classification_follows_the_attractor_map():
    """The interior / IAS rule of yt@proc.f90:170-186 uses the MERGED attractor ids: the pure-Python replay of the
    loop (oracle.yt_reclassify) reproduces the C++ oracle with one basin per maximum and with `ratom 2` merging two
    pairs of maxima; the merged run has far fewer IAS points than the raw one."""
    import helpers as H
    n = (32, 32, 32)
    x2c = S.cell_x2c(6.0, 6.0, 6.0)
    at = S.snap_to_grid(np.array([[0.25, 0.25, 0.25], [0.25 + 1.0 / 6.0, 0.25, 0.25], [0.75, 0.70, 0.30], [0.30, 0.75, 0.75]]), n)
    f = orc.promolecular(n, x2c, at, np.array([4.0, 3.5, 6.0, 2.0]), np.array([2.0, 2.1, 2.4, 1.6]), nimg=1)
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    d0 = orc.yt_integrate(f, x2c, vec, area, atoms=None, ratom=1e-90, atexist=False)
    assert d0.nattr == 4
    assert np.array_equal(orc.yt_reclassify(d0, f, vec, np.arange(1, 5), n), d0.spatial_basin(n))
    d2 = orc.yt_integrate(f, x2c, vec, area, atoms=None, ratom=2.0, atexist=False)
    assert d2.nattr == 3
    pm = np.round(d0.xattr.T * np.array(n)).astype(int) + 1
    mp, na, _ = H.assign_attractors(pm, n, x2c, None, ratom=2.0, atexist=False)
    assert na == 3
    r2 = orc.yt_reclassify(d0, f, vec, mp, n)
    assert np.array_equal(r2, d2.spatial_basin(n))
    assert (r2 == 0).sum() < (d0.spatial_basin(n) == 0).sum()
