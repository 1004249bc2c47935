"""CPU known-answer tests of the oracle's nearest_atom_grid restatement (crystalmod@proc.f90:1138-1167)."""
import numpy as np

import systems as S
from oracle import oracle as orc


def test_one_atom_owns_every_node():
    idg, gap = orc.voronoi_grid((5, 6, 7), S.cell_x2c(4.0, 4.2, 4.4), np.array([[0.3, 0.1, 0.9]]))
    assert idg.min() == idg.max() == 1 and (gap == 1.0).all()   # no second atom: the gap to 'infinity' is 1


def test_two_atoms_along_x_split_the_cell_at_the_midplanes():
    # atoms at x = 0.1 and x = 0.5 of a 10 bohr orthorhombic cell: midplanes at x = 0.3 and x = 0.8 (periodic)
    n = (20, 4, 4)
    idg, gap = orc.voronoi_grid(n, S.cell_x2c(10.0, 7.0, 8.0), np.array([[0.1, 0.5, 0.5], [0.5, 0.5, 0.5]]))
    x = np.arange(20) / 20.0
    want = np.where((x > 0.3 + 1e-9) & (x < 0.8 - 1e-9), 2, 1)
    tie = (np.abs(x - 0.3) < 1e-9) | (np.abs(x - 0.8) < 1e-9)
    for i in range(20):
        if not tie[i]:
            assert (idg[i] == want[i]).all()
        else:
            assert (gap[i] < 1e-12).all() and (idg[i] == 1).all()   # ties: lower id


def test_periodic_images_count():
    # the nearest image of the only other atom lies across the cell boundary
    idg, _ = orc.voronoi_grid((10, 1, 1), S.cell_x2c(10.0, 30.0, 30.0), np.array([[0.05, 0.0, 0.0], [0.55, 0.0, 0.0]]))
    assert idg[9, 0, 0] == 1   # x = 0.9: 1.5 bohr from the image of atom 1 at x = 1.05, 3.5 bohr from atom 2
    assert idg[8, 0, 0] == 1 and idg[5, 0, 0] == 2
