"""PINNED by critic2's own output: the promolecular density restatement (promolecular_atom + grid1%interp, the core of
the HIRSHFELD path) against the density cube that the reference's nodata test 015_grdplot/005_nciplot_basic writes for
the library urea crystal on a 2x2x2 lattice (NCIPLOT stores sign(lambda_2) * rho * 100, six significant digits).

The atomic radial grids are built from the reference's data files dat/wfc/*_pbe.wfc with a restatement of read_critic
(grid1mod@proc.f90:206-318); the structure is dat/lib/crystal.dat's `urea`.  Needs /root/reference (build container);
skipped elsewhere.  The eight golden values are committed in tests/golden/cube_golden.json ("nci_dens")."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

REF = "/root/reference"
CUTRAD = {1: 2.149886192475e+01, 6: 2.167675592180e+01, 7: 1.749805708313e+01, 8: 1.465173060207e+01}   # global.f90:53-56
SYMBOL = {1: "h_", 6: "c_", 7: "n_", 8: "o_"}
CORE_CUTDENS = 1e-8                                                                                 # grid1mod@proc.f90:42


def read_critic(z):
    """grid1%read_critic (grid1mod@proc.f90:206-318) for the neutral atom: r(i), f(i) = sum_orb occ psi^2 / (4 pi r^2),
    cut where the density falls below core_cutdens."""
    tok = open(os.path.join(REF, "dat", "wfc", SYMBOL[z] + "_pbe.wfc")).read().split()
    norb = int(tok[0])
    occ = np.array([float(t) for t in tok[1 + norb: 1 + 2 * norb]])
    assert abs(occ.sum() - z) < 1e-12
    ngrid = int(tok[1 + 3 * norb])
    data = np.array(tok[2 + 3 * norb: 2 + 3 * norb + ngrid * (norb + 1)], dtype=np.float64).reshape(ngrid, norb + 1)
    r, psi = data[:, 0], data[:, 1:]
    rr0 = (psi ** 2) @ occ
    small = np.flatnonzero((rr0 / (4 * np.pi * r ** 2) < CORE_CUTDENS) & (np.arange(ngrid) > 0))
    if small.size:
        ngrid = int(small[0]) + 1                      # g%ngrid = i (the point itself is kept)
    r, rr0 = r[:ngrid], rr0[:ngrid]
    return dict(a=r[0], b=np.log(r[1] / r[0]), ngrid=ngrid, f=rr0 / r ** 2 / (4 * np.pi), r=r)


def urea_system():
    # structure urea, dat/lib/crystal.dat:2146-2167
    cell = (10.51632592951, 10.51632592951, 8.85147720644)
    atoms, zs = [], []
    for ln in open(os.path.join(REF, "dat", "lib", "crystal.dat")).read().split("structure urea")[1].split("endcrystal")[0].split("\n"):
        w = ln.split()
        if w and w[0] == "neq":
            atoms.append([float(w[1]), float(w[2]), float(w[3])])
            zs.append({"C": 6, "O": 8, "N": 7, "H": 1}[w[4]])
    assert len(atoms) == 16
    order = [6, 8, 7, 1]
    tabs = []
    for z in order:
        t = read_critic(z)
        # the log grid of the file is the one interp assumes: r(i) = a exp(b (i-1))
        assert np.abs(t["a"] * np.exp(t["b"] * np.arange(t["ngrid"])) / t["r"] - 1).max() < 1e-9
        t["rcut"] = min(CUTRAD[z], t["r"][-1])
        tabs.append(t)
    g = orc.AtomicGrids(tabs)
    for k, t in enumerate(tabs):                        # the nodes as stored in the file, not the regenerated ones
        g.rtab[g.off[k]: g.off[k] + t["ngrid"]] = t["r"]
        g.rmax[k] = t["r"][-1]
        g.rcut[k] = t["rcut"]
    ispc = np.array([order.index(z) + 1 for z in zs], dtype=np.int32)
    return np.diag(cell), np.array(atoms), ispc, g


def _gold(key):
    b = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cube_golden.json")))[key]
    return np.array(b["text"].split(), dtype=np.float64).reshape(2, 2, 2)                         # file order: i, j, k (k fastest)


needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dat", "wfc")), reason="needs the reference's data files")


@needs_ref
def test_promolecular_density_of_urea_matches_the_reference_cube():
    x2c, atoms, ispc, g = urea_system()
    rho = orc.promolecular_grid((2, 2, 2), x2c, atoms, ispc, g)
    want = np.abs(_gold("nci_dens")) / 100.0
    assert np.abs(rho / want - 1.0).max() <= 6e-6, (rho, want)                                    # six printed digits


@needs_ref
def test_reduced_density_gradient_formula_matches_the_reference_cube():
    """The grad cube of the same run: s = |grad rho| / (2 (3 pi^2)^(1/3) rho^(4/3)) (nci@proc.f90:91, :554, :571), 100 where
    rho exceeds the NCIPLOT cutoff, ~1e-15 at the symmetry-fixed points.  |grad rho| comes from central differences of the
    pinned promolecular density (atoms displaced by -+h): this pins the constant and the exponent that nci.cu uses."""
    x2c, atoms, ispc, g = urea_system()
    c2x = np.linalg.inv(x2c)
    rho = orc.promolecular_grid((2, 2, 2), x2c, atoms, ispc, g)
    h = 1e-3
    grad = np.zeros((3, 2, 2, 2))
    for k in range(3):
        d = np.zeros(3); d[k] = h
        shift = c2x @ d
        grad[k] = (orc.promolecular_grid((2, 2, 2), x2c, atoms - shift, ispc, g) -
                   orc.promolecular_grid((2, 2, 2), x2c, atoms + shift, ispc, g)) / (2 * h)
    s = np.sqrt((grad ** 2).sum(axis=0)) / (2.0 * (3.0 * np.pi ** 2) ** (1.0 / 3.0) * rho ** (4.0 / 3.0))
    want = _gold("nci_grad")
    sel = (want > 1e-3) & (want < 99.0)                    # the points that carry a genuine RDG value
    assert sel.sum() == 2
    assert np.abs(s[sel] / want[sel] - 1.0).max() <= 2e-5, (s, want)
    assert np.all(s[want < 1e-3] < 1e-4)                   # symmetry-fixed points: zero gradient (finite-difference noise only)


def urea_fixture():
    """The same system from the committed fixture (tests/golden/urea_atomic_grids.npz, made by make_urea_grids.py)."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "urea_atomic_grids.npz"))
    g = orc.AtomicGrids([dict(a=float(z["a"][k]), b=float(z["b"][k]), ngrid=int(z["ngrid"][k]),
                              f=z["ftab"][z["off"][k]: z["off"][k] + z["ngrid"][k]]) for k in range(4)])
    g.rtab[:] = z["rtab"]; g.rmax[:] = z["rmax"]; g.rcut[:] = z["rcut"]
    return z["x2c"], z["atoms"], z["ispc"], g


def golden_urea_grid():
    b = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cube_golden.json")))["shift"]
    return np.array(b["plain_text"].split(), dtype=np.float64).reshape(b["n"])           # (ix, iy, iz)


def test_promolecular_grid_of_urea_reproduces_the_reference_cube_to_14_digits():
    """PINNED: `load as "$0" 10 10 10` + `cube grid` of the reference's nodata test 005_plot/016_cube_grid -- 1000 values of
    the promolecular density of urea with 14 printed digits.  The restatement (read_critic tables, grid1%interp, the
    image sum with the cutrad cutoffs) reproduces every one of them to 1e-12 (measured 5e-14)."""
    x2c, atoms, ispc, g = urea_fixture()
    rho = orc.promolecular_grid((10, 10, 10), x2c, atoms, ispc, g)
    want = golden_urea_grid()
    assert np.abs(rho / want - 1.0).max() <= 1e-12


@needs_ref
def test_the_fixture_is_what_read_critic_builds_from_the_reference_data():
    x2c, atoms, ispc, g = urea_system()
    x2, a2, i2, g2 = urea_fixture()
    assert np.array_equal(x2c, x2) and np.array_equal(atoms, a2) and np.array_equal(ispc, i2)
    for k in ("ngrid", "off", "a", "b", "rmax", "rcut", "rtab", "ftab"):
        assert np.array_equal(getattr(g, k), getattr(g2, k)), k
