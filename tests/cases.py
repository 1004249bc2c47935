"""Seeded parity cases shared by the CPU (oracle) and GPU tests."""
import numpy as np

import systems as S
from oracle import oracle as orc

# (name, n, number of atoms, cell parameters, seed)
SMALL_CASES = [
    ("cubic48", (48, 48, 48), 6, (9, 9, 9, 90, 90, 90), 1),
    ("triclinic", (64, 68, 72), 8, (10, 10.5, 11, 85, 95, 100), 2),
    ("cubic96", (96, 96, 96), 32, (16, 16, 16, 90, 90, 90), 3),
    ("odd_dims", (50, 61, 47), 5, (8, 9.5, 7.7, 90, 90, 90), 7),   # not multiples of 4 / 2
    ("ortho_flat", (36, 80, 48), 4, (6, 13, 8, 90, 90, 90), 8),
    ("tiny", (9, 10, 11), 1, (4, 4.2, 4.4, 90, 90, 90), 9),        # a single basin, grid smaller than a tile
]


# A deliberately degenerate case: a 5-bohr-thin cell in which the atoms interact with their own periodic
# images.  In the low-density valley between the images the discrete near-grid map is chaotic (isolated
# points whose own trajectory ends in another basin than all of their neighbours'), and the sequential
# reference keeps for such points whatever label an earlier path left behind: its result is scan-order
# dependent there and no order-independent algorithm can reproduce it bit for bit (5 of 80640 points).
DEGENERATE_CASES = [
    ("thin_cell", (36, 80, 28), 4, (6, 13, 5, 90, 90, 90), 8),
]


def make_case(name):
    for c in SMALL_CASES + DEGENERATE_CASES:
        if c[0] == name:
            _, n, nat, cellp, seed = c
            x2c = S.cell_x2c(*cellp)
            at, z, al = S.random_atoms(nat, seed, x2c, dmin=1.6 if nat > 1 else 0.0)
            at = S.snap_to_grid(at, n)
            f = orc.promolecular(n, x2c, at, z, al, nimg=1)
            return dict(name=name, n=n, x2c=x2c, atoms=at, z=z, alpha=al, f=f)
    raise KeyError(name)


def second_field(f):
    """A second INTEGRABLE field on the same grid (a crude Laplacian-like stencil; only has to be
    the same array on both sides)."""
    out = np.zeros_like(f)
    for ax in range(3):
        out += np.roll(f, 1, ax) + np.roll(f, -1, ax) - 2.0 * f
    return np.asfortranarray(out)
