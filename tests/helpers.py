"""Helpers shared by the parity tests."""
from __future__ import annotations

import numpy as np


def shortest(x2c, dx):
    dx = dx - np.round(dx)
    best = 1e300
    for a in (-2, -1, 0, 1, 2):
        for b in (-2, -1, 0, 1, 2):
            for c in (-2, -1, 0, 1, 2):
                d = np.linalg.norm(x2c @ (dx + np.array([a, b, c], dtype=float)))
                best = min(best, d)
    return best


def assign_attractors(pmax, n, x2c, atoms=None, ratom=1.0, atexist=True):
    """Attractor identification of bader_integrate (bader@proc.f90:160-199) applied to an ordered
    list of maxima (1-based grid coordinates).  Returns (map[nmax] 1-based ids, nattr, xattr[3,nattr])."""
    n = np.asarray(n, dtype=float)
    atoms = np.zeros((0, 3)) if atoms is None else np.asarray(atoms, dtype=float)
    xattr = [a for a in atoms] if (atexist and len(atoms)) else []
    mp = []
    for p in pmax:
        dv = (np.asarray(p, dtype=float) - 1.0) / n
        lab = 0
        if atexist and len(atoms):
            d = [shortest(x2c, dv - a) for a in atoms]
            k = int(np.argmin(d))
            if d[k] <= ratom:
                lab = k + 1
        if lab == 0 and ratom > 1e-80:
            for l, xa in enumerate(xattr):
                if shortest(x2c, dv - xa) < ratom:
                    lab = l + 1
                    break
        if lab == 0:
            xattr.append(dv)
            lab = len(xattr)
        mp.append(lab)
    return np.array(mp, dtype=np.int32), len(xattr), np.array(xattr).T if xattr else np.zeros((3, 0))
