"""GPU parity test of VORONOI on a grid through the C ABI (pytest -m gpu): voronoi_grid (hirshfeld@proc.f90:93-122) =
crystal%nearest_atom_grid (crystalmod@proc.f90:1138-1167), then the atomic integrals of intgrid_fields on its idg.

Bar: the nearest-atom id of every node whose nearest atom is unique (distance gap to the second-nearest atom above
1e-12 relative) bit-exact against a brute-force restatement; nodes equidistant from two atoms (the reference resolves
them by the traversal order of list_near_atoms) go to the lower id on the device and are counted."""
import numpy as np
import pytest

import systems as S
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CASES = {
    "triclinic": (S.cell_x2c(7.0, 7.5, 8.0, 84, 95, 103), 5, (22, 25, 27)),
    "ortho": (S.cell_x2c(9.0, 6.0, 11.0), 7, (37, 18, 25)),
    "one_atom": (S.cell_x2c(4.0, 4.2, 4.4), 1, (9, 10, 11)),
    "small_cell_many_atoms": (S.cell_x2c(3.2, 3.0, 3.4, 90, 96, 90), 24, (12, 11, 13)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_voronoi_grid_and_atomic_integrals(ctx, name):
    x2c, nat, n = CASES[name]
    rng = np.random.default_rng(17)
    atoms = rng.uniform(0, 1, (nat, 3))
    idg_o, gap = orc.voronoi_grid(n, x2c, atoms)
    b = ctx.voronoi_grid(n, x2c, atoms)
    idg = b.labels(n)
    unique = gap > 1e-12
    assert np.array_equal(idg[unique], idg_o[unique])
    ties = int((~unique).sum())
    assert np.array_equal(idg[~unique], idg_o[~unique]) or ties > 0   # ties: lower id on both sides by construction
    print(f"VORONOI {name}: {idg.size} nodes, {ties} equidistant from two atoms")
    # the atomic integrals: plain sums over idg like the Bader branch (integration@proc.f90:1208-1218, :1289-1299)
    f = np.asfortranarray(1.0 + rng.uniform(0, 1, n))
    h = ctx.upload(f)
    om = S.omega(x2c)
    vol, ps = ctx.integrate(b, [h], om)
    vref, pref = orc.integrate_bader(idg_o, [f], nat, om)
    if ties == 0:
        assert np.array_equal(vol, vref)
        assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
    assert abs(vol.sum() - om) <= 1e-12 * om
    b.free(); ctx.free(h)


def test_voronoi_symmetric_structure_ties_go_to_the_lower_id(ctx):
    """Two atoms at (0,0,0) and (1/2,1/2,1/2) of a cubic cell on an even grid: whole planes of nodes are equidistant."""
    x2c = S.cell_x2c(6.0, 6.0, 6.0)
    atoms = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    n = (12, 12, 12)
    idg_o, gap = orc.voronoi_grid(n, x2c, atoms)
    b = ctx.voronoi_grid(n, x2c, atoms)
    idg = b.labels(n)
    unique = gap > 1e-12
    assert (~unique).sum() > 0
    assert np.array_equal(idg[unique], idg_o[unique])
    # exactly representable coordinates: the squared distances tie exactly and both sides pick atom 1
    assert np.array_equal(idg, idg_o)
    b.free()
