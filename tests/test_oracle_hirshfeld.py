"""Oracle checks of the HIRSHFELD restatement: grid1%interp (grid1mod@proc.f90:86-137), promolecular_array3
(crystalmod@complex.f90:436-470 on promolecular_atom, crystalmod@env.f90:622-748) and the loop of
intgrid_hirshfeld_fields (integration@proc.f90:1552-1596)."""
import numpy as np

import systems as S
from oracle import oracle as orc


def slater_tables(zs, alphas, ngrid=400, a=1e-4, rmax=9.0, rcut=None):
    """Synthetic atomic grids: rho(r) = Z alpha^3 / (8 pi) exp(-alpha r), normalised to Z electrons."""
    b = np.log(rmax / a) / (ngrid - 1)
    r = a * np.exp(b * np.arange(ngrid))
    out = []
    for z, al in zip(zs, alphas):
        t = dict(a=a, b=b, ngrid=ngrid, f=z * al ** 3 / (8 * np.pi) * np.exp(-al * r))
        if rcut is not None:
            t["rcut"] = rcut
        out.append(t)
    return orc.AtomicGrids(out)


def small_system():
    x2c = S.cell_x2c(7.0, 7.5, 8.0, 90, 95, 100)
    atoms = np.array([[0.1, 0.2, 0.3], [0.6, 0.7, 0.25], [0.35, 0.8, 0.75]])
    ispc = np.array([1, 2, 1], dtype=np.int32)
    grids = slater_tables([8.0, 1.0], [2.2, 1.9])
    return x2c, atoms, ispc, grids


def test_interp_reproduces_the_tabulated_function_and_the_nodes():
    g = slater_tables([6.0], [2.0])
    r = g.rtab[: g.ngrid[0]]
    for k in (0, 1, 57, 200, 398):
        assert abs(g.interp(0, r[k]) - g.ftab[k]) <= 1e-13 * g.ftab[k]       # a node: the Lagrange weights are 0/1
    rng = np.random.default_rng(0)
    for r0 in np.exp(rng.uniform(np.log(2e-4), np.log(8.9), 200)):
        want = 6.0 * 8.0 / (8 * np.pi) * np.exp(-2.0 * r0)
        assert abs(g.interp(0, r0) - want) <= 1e-2 * want                      # cubic interpolation on a 400-node log grid (coarse at large r)
    # four-node Lagrange interpolation is exact for cubics: pins the weights, the node window and its clamping at both ends
    poly = lambda x: 1.0 + 2.0 * x - 0.5 * x ** 2 + 0.1 * x ** 3
    gp = orc.AtomicGrids([dict(a=1e-4, b=g.b[0], ngrid=400, f=poly(r))])
    for r0 in np.concatenate([np.exp(rng.uniform(np.log(1e-4), np.log(8.99), 300)), r[:3] * 1.3, r[-3:] * 0.999]):
        assert abs(gp.interp(0, r0) - poly(r0)) <= 1e-11 * abs(poly(r0))
    assert abs(g.interp(0, 1e-6) - g.ftab[0]) <= 1e-14 * g.ftab[0]                                      # below r(1): the value at r(1)
    assert g.interp(0, 9.0) == 0.0 and g.interp(0, 20.0) == 0.0                # r0 >= rmax


def test_promolecular_grid_against_a_direct_lattice_sum():
    x2c, atoms, ispc, grids = small_system()
    n = (12, 13, 14)
    f = orc.promolecular_grid(n, x2c, atoms, ispc, grids)
    rng = np.random.default_rng(1)
    for _ in range(40):
        p = np.array([rng.integers(0, n[k]) for k in range(3)])
        xc = x2c @ (p / np.array(n))
        want = 0.0
        for a, isp in zip(atoms, ispc):
            for L in np.ndindex(7, 7, 7):
                d = np.linalg.norm(xc - x2c @ (a + np.array(L) - 3))
                if d <= grids.rcut[isp - 1]:
                    want += max(grids.interp(isp - 1, max(d, grids.rtab[grids.off[isp - 1]], 1e-14)), 0.0)
        assert abs(f[tuple(p)] - want) <= 1e-12 * want
    # fragments add up to the whole (promolecular_array3 with fr = one atom, hirshfeld@proc.f90:83)
    parts = sum(orc.promolecular_grid(n, x2c, atoms, ispc, grids, infrag=np.eye(3, dtype=np.uint8)[k]) for k in range(3))
    assert np.abs(parts - f).max() <= 1e-13 * f.max()


def test_hirshfeld_weights_partition_the_cell():
    x2c, atoms, ispc, grids = small_system()
    n = (16, 16, 18)
    om = S.omega(x2c)
    promol = orc.promolecular_grid(n, x2c, atoms, ispc, grids)
    vol, ps = orc.hirshfeld_fields(promol, x2c, atoms, ispc, grids, [promol], om)
    # sum over atoms of w_A = sum_A rho_A / rho_promol = 1 wherever no clamp acts
    assert abs(vol.sum() - om) <= 1e-10 * om
    assert abs(ps[:, 0].sum() - promol.sum() * om / promol.size) <= 1e-10 * ps[:, 0].sum()
    # each atom's Hirshfeld population of the promolecular density is its own density integrated (~Z, grid error)
    assert abs(ps[0, 0] / ps[1, 0] - 8.0) < 0.8
    # ONLY: masked atoms stay at zero, the others are unchanged
    vol2, ps2 = orc.hirshfeld_fields(promol, x2c, atoms, ispc, grids, [promol], om, domask=[1, 0, 1])
    assert vol2[1] == 0.0 and ps2[1, 0] == 0.0 and np.array_equal(vol2[[0, 2]], vol[[0, 2]])
