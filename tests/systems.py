"""Seeded synthetic systems shared by the tests and bench.py.

Pure numpy/scipy host-side helpers: cells, atoms, the Voronoi-relevant grid
stencil (a restatement of what init_geometry/wscell hand to the hot path,
grid3mod@proc.f90:3167-3215, tools@proc.f90:679-898) and metrics.
"""
from __future__ import annotations

import itertools

import numpy as np


def cell_x2c(a, b, c, alpha=90.0, beta=90.0, gamma=90.0):
    """Crystallographic -> Cartesian matrix (columns = cell vectors), bohr."""
    al, be, ga = np.deg2rad([alpha, beta, gamma])
    va = np.array([a, 0.0, 0.0])
    vb = np.array([b * np.cos(ga), b * np.sin(ga), 0.0])
    cx = c * np.cos(be)
    cy = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
    cz = np.sqrt(max(c * c - cx * cx - cy * cy, 0.0))
    m = np.stack([va, vb, np.array([cx, cy, cz])], axis=1)
    m[np.abs(m) < 1e-13 * np.abs(m).max()] = 0.0
    return m


def random_atoms(nat, seed, x2c=None, dmin=1.6):
    """nat random fractional positions with a minimum-image separation >= dmin bohr."""
    rng = np.random.default_rng(seed)
    pts = []
    x2c = np.eye(3) * 10.0 if x2c is None else x2c
    tries = 0
    while len(pts) < nat:
        tries += 1
        if tries > 100000:
            raise RuntimeError("cannot place atoms")
        x = rng.random(3)
        ok = True
        for y in pts:
            d = x - y
            d -= np.round(d)
            if np.linalg.norm(x2c @ d) < dmin:
                ok = False
                break
        if ok:
            pts.append(x)
    z = rng.uniform(1.0, 8.0, nat)
    alpha = rng.uniform(1.2, 2.7, nat)
    return np.array(pts), z, alpha


def jittered_lattice(m, seed, jitter=0.15):
    """m^3 atoms on a jittered simple-cubic lattice (config 5: 8x8x8 = 512 atoms)."""
    rng = np.random.default_rng(seed)
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / m
    g = g + rng.uniform(-jitter, jitter, g.shape) / m
    g = g % 1.0
    z = rng.uniform(1.0, 8.0, len(g))
    alpha = rng.uniform(1.2, 2.7, len(g))
    return g, z, alpha


def snap_to_grid(atoms, n):
    """Move atoms onto grid nodes so that the density maxima coincide with nuclei."""
    n = np.asarray(n, dtype=np.float64)
    return (np.round(atoms * n) % n) / n


def omega(x2c):
    return abs(np.linalg.det(x2c))


def wscell(x2cg, nmax=3):
    """Voronoi-relevant lattice vectors and facet areas of the lattice with basis
    x2cg (columns), i.e. grid%vec / grid%area (tools@proc.f90:784-836).
    Returns (vec[nvec,3] int32, area[nvec]).  Uses scipy's qhull binding."""
    from scipy.spatial import Voronoi

    rng = range(-nmax, nmax + 1)
    ijk = np.array(list(itertools.product(rng, rng, rng)), dtype=np.int64)
    pts = ijk @ x2cg.T
    vor = Voronoi(pts)
    i0 = int(np.where((ijk == 0).all(axis=1))[0][0])
    vecs, areas = [], []
    for (p, q), verts in zip(vor.ridge_points, vor.ridge_vertices):
        if i0 not in (p, q) or -1 in verts:
            continue
        other = q if p == i0 else p
        v = vor.vertices[verts]
        # area of the (convex, planar) facet: order vertices around the centroid
        cen = v.mean(axis=0)
        nrm = pts[other] - pts[i0]
        nrm = nrm / np.linalg.norm(nrm)
        e1 = v[0] - cen
        e1 -= nrm * (e1 @ nrm)
        e1 /= np.linalg.norm(e1)
        e2 = np.cross(nrm, e1)
        ang = np.arctan2((v - cen) @ e2, (v - cen) @ e1)
        v = v[np.argsort(ang)]
        ar = 0.0
        for k in range(len(v)):
            ar += 0.5 * np.linalg.norm(np.cross(v[k] - cen, v[(k + 1) % len(v)] - cen))
        if ar < 1e-10 * np.linalg.norm(pts[other]) ** 2:
            continue  # degenerate (zero-area) facet: not Voronoi-relevant
        vecs.append(ijk[other])
        areas.append(ar)
    order = np.lexsort(np.array(vecs).T[::-1])
    return np.array(vecs, dtype=np.int32)[order], np.array(areas)[order]


def quantize(f, digits):
    """Round to `digits` significant digits (CHGCAR E18.11 -> digits=12)."""
    out = np.array([float(f"{v:.{digits - 1}E}") for v in f.ravel(order="F")])
    return out.reshape(f.shape, order="F")
