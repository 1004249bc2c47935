"""GPU parity tests of HIRSHFELD on a grid through the C ABI (pytest -m gpu): promolecular_array3
(crystalmod@complex.f90:436-470) and the loop of intgrid_hirshfeld_fields (integration@proc.f90:1552-1596).

Bar: floating-point sums over the same set of atom images in a different order -> 1e-10 relative to the largest value."""
import numpy as np
import pytest

import systems as S
from critic2_b200 import capi
from oracle import oracle as orc
from test_oracle_hirshfeld import slater_tables

pytestmark = pytest.mark.gpu
TOL = 1e-10


def tab_of(g):
    return dict(ngrid=g.ngrid, off=g.off, a=g.a, b=g.b, rmax=g.rmax, rcut=g.rcut, rtab=g.rtab, ftab=g.ftab)


SYSTEMS = {
    # triclinic, 3 atoms of 2 species, cutoff (9 bohr) larger than the cell: several images of every atom contribute
    "triclinic": (S.cell_x2c(7.0, 7.5, 8.0, 90, 95, 100), np.array([[0.1, 0.2, 0.3], [0.6, 0.7, 0.25], [0.35, 0.8, 0.75]]),
                  np.array([1, 2, 1], dtype=np.int32), ([8.0, 1.0], [2.2, 1.9]), (20, 21, 23), None),
    # orthorhombic, a species without a grid (skipped like z = 0) and a cutoff below rmax; n1 not a multiple of the tile
    "ortho_cut": (S.cell_x2c(9.0, 6.0, 11.0), np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5], [0.25, 0.1, 0.8], [1.7, -0.3, 0.4]]),
                  np.array([1, 2, 3, 1], dtype=np.int32), ([6.0, 3.0, 1.0], [2.0, 1.6, 2.4]), (37, 18, 25), 5.5),
}


def make(name):
    x2c, atoms, ispc, (zs, als), n, rcut = SYSTEMS[name]
    g = slater_tables(zs, als, rcut=rcut)
    if name == "ortho_cut":
        g.ngrid[1] = 0          # species 2: no usable grid
    return x2c, atoms, ispc, g, n


@pytest.mark.parametrize("name", list(SYSTEMS))
def test_promolecular_grid_and_hirshfeld_sums(ctx, name):
    x2c, atoms, ispc, g, n = make(name)
    om = S.omega(x2c)
    ref = orc.promolecular_grid(n, x2c, atoms, ispc, g)
    h = ctx.promolecular_grid(n, x2c, atoms, ispc, tab_of(g))
    got = ctx.download(h, n)
    assert np.abs(got - ref).max() <= TOL * ref.max()
    # one-atom fragment = the numerator of hirsh_weights (hirshfeld@proc.f90:78-86)
    fr = np.zeros(len(atoms), dtype=np.uint8); fr[0] = 1
    hf = ctx.promolecular_grid(n, x2c, atoms, ispc, tab_of(g), infrag=fr)
    assert np.abs(ctx.download(hf, n) - orc.promolecular_grid(n, x2c, atoms, ispc, g, infrag=fr)).max() <= TOL * ref.max()
    ctx.free(hf)
    # the Hirshfeld sums: volume + two properties, then with an ONLY mask
    rng = np.random.default_rng(5)
    f2 = np.asfortranarray(ref * (1.0 + 0.3 * rng.standard_normal(n)))
    h2 = ctx.upload(f2)
    for dm in (None, np.array([1] + [0] * (len(atoms) - 2) + [1], dtype=np.uint8)):
        vol_o, ps_o = orc.hirshfeld_fields(ref, x2c, atoms, ispc, g, [ref, f2], om, domask=dm)
        vol, ps = ctx.hirshfeld_integrate(h, x2c, atoms, ispc, tab_of(g), [h, h2], om, domask=dm)
        assert np.abs(vol - vol_o).max() <= TOL * np.abs(vol_o).max()
        assert np.abs(ps - ps_o).max() <= TOL * np.abs(ps_o).max()
        if dm is not None:
            assert vol[1] == 0.0 and not ps[1].any()
    vol0, ps0 = ctx.hirshfeld_integrate(h, x2c, atoms, ispc, tab_of(g), [], om)      # volumes only
    vol_o, _ = orc.hirshfeld_fields(ref, x2c, atoms, ispc, g, [], om)
    assert ps0.shape == (len(atoms), 0) and np.abs(vol0 - vol_o).max() <= TOL * np.abs(vol_o).max()
    ctx.free(h); ctx.free(h2)


def test_hirshfeld_error_paths(ctx):
    x2c, atoms, ispc, g, n = make("triclinic")
    with pytest.raises(capi.C2GError, match="species"):
        ctx.promolecular_grid(n, x2c, atoms, np.array([1, 5, 1], dtype=np.int32), tab_of(g))
    with pytest.raises(capi.C2GError, match="handle"):
        ctx.hirshfeld_integrate(999, x2c, atoms, ispc, tab_of(g), [], 1.0)


def test_promolecular_grid_of_urea_reproduces_the_reference_cube(ctx):
    """PINNED by critic2's own output: the device evaluates the promolecular density of the library urea crystal on the
    10x10x10 grid of the reference's nodata test 005_plot/016_cube_grid (atomic grids built from the reference's data
    files, tests/golden/urea_atomic_grids.npz) and must reproduce the 1000 values of the reference's cube file, 14
    printed digits, to 1e-11; the Hirshfeld volumes of that density then partition the cell."""
    from test_oracle_promolecular_reference import golden_urea_grid, urea_fixture
    x2c, atoms, ispc, g = urea_fixture()
    n = (10, 10, 10)
    h = ctx.promolecular_grid(n, x2c, atoms, ispc, tab_of(g))
    rho = ctx.download(h, n)
    assert np.abs(rho / golden_urea_grid() - 1.0).max() <= 1e-11
    om = S.omega(x2c)
    vol, ps = ctx.hirshfeld_integrate(h, x2c, atoms, ispc, tab_of(g), [h], om)
    assert abs(vol.sum() - om) <= 1e-10 * om
    vol_o, ps_o = orc.hirshfeld_fields(rho, x2c, atoms, ispc, g, [rho], om)
    assert np.abs(ps - ps_o).max() <= TOL * np.abs(ps_o).max() and np.abs(vol - vol_o).max() <= TOL * np.abs(vol_o).max()
    ctx.free(h)


def test_multi_round_image_culling(ctx):
    """A 3.2 bohr cell with 24 atoms and 9 bohr cutoffs: about 2700 atom images reach every point, more than the 2048
    entries the per-tile culling list holds, so every tile needs a second (and third) culling round (hirshfeld.cu,
    HB_LIST)."""
    x2c = S.cell_x2c(3.2, 3.0, 3.4, 90, 96, 90)
    rng = np.random.default_rng(3)
    atoms = rng.uniform(0, 1, (24, 3))
    ispc = np.array([1, 2] * 12, dtype=np.int32)
    g = slater_tables([6.0, 1.0], [2.1, 1.7])
    n = (12, 11, 13)
    om = S.omega(x2c)
    ref = orc.promolecular_grid(n, x2c, atoms, ispc, g)
    h = ctx.promolecular_grid(n, x2c, atoms, ispc, tab_of(g))
    got = ctx.download(h, n)
    assert np.abs(got - ref).max() <= TOL * ref.max()
    vol_o, ps_o = orc.hirshfeld_fields(ref, x2c, atoms, ispc, g, [ref], om)
    vol, ps = ctx.hirshfeld_integrate(h, x2c, atoms, ispc, tab_of(g), [h], om)
    assert np.abs(vol - vol_o).max() <= TOL * np.abs(vol_o).max()
    assert np.abs(ps - ps_o).max() <= TOL * np.abs(ps_o).max()
    assert abs(vol.sum() - om) <= 1e-9 * om
    ctx.free(h)
