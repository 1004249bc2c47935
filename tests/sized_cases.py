"""BASELINE.json-sized parity cases, shared by tests/test_gpu_at_size.py and
tools/golden_at_size.py.  The densities are generated in HBM by c2g_promolecular and downloaded for the oracle,
so both sides always see the same array.

name -> dict(n, x2c, atoms, z, alpha, nimg, rc)
"""
import numpy as np

import systems as S


def headline(size):
    """bench.py's workload: configs[4] scaled to size^3 (5 bohr atom spacing, 128 points per spacing)."""
    side = max(2, size // 128)
    n = (size, size, size)
    x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
    at, z, al = S.jittered_lattice(side, 5)
    return dict(n=n, x2c=x2c, atoms=S.snap_to_grid(at, n), z=z, alpha=al, nimg=1, rc=8.0)


def urea(N=256):
    """configs[1]: tetragonal cell 10.52 x 10.52 x 8.85 bohr, 16 atoms, N^3 (same case as tools/bench_paths.py)."""
    n = (N, N, N)
    x2c = S.cell_x2c(10.52, 10.52, 8.85)
    at, z, al = S.random_atoms(16, 2, x2c, dmin=2.0)
    return dict(n=n, x2c=x2c, atoms=S.snap_to_grid(at, n), z=z, alpha=al, nimg=1, rc=0.0)


def hetero(N=192):
    """Heterogeneous basins in a triclinic cell: four heavy, compact atoms, each with two light diffuse atoms
    1.8 bohr away (X-H like): the light basins are several times smaller than the mean basin."""
    n = (N, N + 8, N - 8)
    x2c = S.cell_x2c(14.0, 15.0, 13.0, 84.0, 97.0, 105.0)
    rng = np.random.default_rng(11)
    heavy, _, _ = S.random_atoms(4, 11, x2c, dmin=5.5)
    c2x = np.linalg.inv(x2c)
    pts, z, al = [], [], []
    for h in heavy:
        pts.append(h); z.append(8.0); al.append(2.6)
        for _ in range(2):
            d = rng.normal(size=3); d *= 1.8 / np.linalg.norm(d)
            pts.append((h + c2x @ d) % 1.0); z.append(1.0); al.append(1.3)
    return dict(n=n, x2c=x2c, atoms=S.snap_to_grid(np.array(pts), n), z=np.array(z), alpha=np.array(al), nimg=1, rc=0.0)


def molecule_in_vacuum(N=256):
    """A benzene-dimer-like molecule (24 atoms) in a 30 bohr cubic box: most of the cell is a low-density tail
    shared by a few huge basins, next to 1-2 bohr wide hydrogen basins."""
    n = (N, N, N)
    x2c = S.cell_x2c(30.0, 30.0, 30.0)
    ang = np.arange(6) * np.pi / 3
    ring = np.stack([2.64 * np.cos(ang), 2.64 * np.sin(ang), np.zeros(6)], 1)
    hyd = np.stack([4.69 * np.cos(ang), 4.69 * np.sin(ang), np.zeros(6)], 1)
    mono = np.concatenate([ring, hyd])
    dimer = np.concatenate([mono + [15.0, 15.0, 11.7], mono + [15.0 + 3.0, 15.0, 11.7 + 6.6]])
    z = np.array(([6.0] * 6 + [1.0] * 6) * 2); al = np.array(([2.2] * 6 + [1.9] * 6) * 2)
    return dict(n=n, x2c=x2c, atoms=S.snap_to_grid(dimer / 30.0, n), z=z, alpha=al, nimg=0, rc=0.0)


def flat_cell(N=160, degenerate=False):
    """Non-cubic grid and cell: 2N x N x 0.8N+4 points, orthorhombic 20 x 10 x 9 bohr, 12 atoms.
    degenerate=True: 20 x 10 x 8.2 bohr with 16 atoms -- a cell in which long trajectories run along the interatomic
    surfaces (ridge runners): the sequential reference then keeps scan-order dependent labels (its refine_edge needs 21
    iterations instead of 2-3) and 19 of 6.8e6 points differ between the oracle and the own-trajectory labelling that
    the exact-walk referee computes.  Such cases are REPORTED by the tests, not asserted (like tests/cases.py
    DEGENERATE_CASES); no parallel algorithm can reproduce labels that depend on the reference's scan order."""
    n = (2 * N, N, (4 * N) // 5 + 4)
    if degenerate:
        x2c = S.cell_x2c(20.0, 10.0, 8.2)
        at, z, al = S.random_atoms(16, 21, x2c, dmin=2.2)
    else:
        x2c = S.cell_x2c(20.0, 10.0, 9.0)
        at, z, al = S.random_atoms(12, 21, x2c, dmin=2.5)
    return dict(n=n, x2c=x2c, atoms=S.snap_to_grid(at, n), z=z, alpha=al, nimg=1, rc=0.0)


CASES = {
    "urea256": lambda: urea(256),
    "head256": lambda: headline(256),
    "head512": lambda: headline(512),
    "head1024": lambda: headline(1024),
    "hetero192": lambda: hetero(192),
    "molvac256": lambda: molecule_in_vacuum(256),
    "flat160": lambda: flat_cell(160),
    "flat160d": lambda: flat_cell(160, degenerate=True),
}


def atom_map(pmax, n, x2c, atoms):
    """Vectorised attractor identification for the atoms mode with nuclei on grid nodes (bader@proc.f90:160-175,
    identify_atom with distmax = ratom): nearest atom (minimum image over +-2 cells) of every maximum.
    Returns (map[nmax] 1-based, dist[nmax] in bohr)."""
    dv = (np.asarray(pmax, dtype=float) - 1.0) / np.asarray(n, dtype=float)[None, :]
    d = dv[:, None, :] - np.asarray(atoms, dtype=float)[None, :, :]
    d -= np.round(d)
    best = np.full(d.shape[:2], np.inf)
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            for c in (-1, 0, 1):
                v = (d + np.array([a, b, c], dtype=float)) @ np.asarray(x2c).T
                best = np.minimum(best, np.sqrt((v * v).sum(axis=2)))
    k = best.argmin(axis=1)
    return (k + 1).astype(np.int32), best[np.arange(len(k)), k]
