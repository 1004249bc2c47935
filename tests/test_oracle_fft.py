"""CPU tests of the oracle's grid3%fft / trilinear / NCI-FOURIER restatements -- no GPU.

The DFT inside the oracle is checked against numpy.fft (an independent implementation) through a numpy
restatement of the reference formulas (grid3mod@proc.f90:1785-1866)."""
import numpy as np
import pytest

import cases
from oracle import oracle as orc


def numpy_fft_derivative(f, x2c, what):
    n = f.shape
    a = [x2c[:, i] for i in range(3)]
    vol = abs(np.linalg.det(x2c))
    b = [2 * np.pi / vol * np.cross(a[2], a[1]), 2 * np.pi / vol * np.cross(a[0], a[2]), 2 * np.pi / vol * np.cross(a[1], a[0])]

    def fr(nn):  # i in (-n/2, n/2]  (:1792-1794)
        i = np.arange(nn)
        return np.where(i <= nn // 2, i, i - nn)

    i1, i2, i3 = np.meshgrid(fr(n[0]), fr(n[1]), fr(n[2]), indexing="ij")
    v = [i1 * b[0][d] + i2 * b[1][d] + i3 * b[2][d] for d in range(3)]
    F = np.fft.fftn(f) / f.size          # cfftnd forward is scaled by 1/ntot
    inv = lambda Z: np.fft.ifftn(Z) * f.size
    if what == "grad":
        out = 0
        for d in range(3):
            out = out + inv(v[d] * 1j * F).real ** 2
        return np.sqrt(out)
    m = {"x": -v[0] * 1j, "y": -v[1] * 1j, "z": -v[2] * 1j, "xx": -v[0] * v[0], "xy": -v[0] * v[1], "xz": -v[0] * v[2],
         "yy": -v[1] * v[1], "yz": -v[1] * v[2], "zz": -v[2] * v[2], "lap": -(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)}
    if what == "pot":
        v2 = v[0] ** 2 + v[1] ** 2 + v[2] ** 2
        mm = np.where(v2 < 1e-12, 0, -1 / np.where(v2 < 1e-12, 1, v2))
        return -4 * np.pi * inv(mm * F).real
    return inv(m[what] * F).real


@pytest.mark.parametrize("n", [(12, 10, 9), (8, 8, 8), (7, 5, 3), (16, 6, 15)])
def test_fft_derivative_matches_numpy(n):
    rng = np.random.default_rng(sum(n))
    x2c = np.array([[5.0, 0.3, 0.1], [0.0, 4.5, 0.2], [0.0, 0.0, 6.0]])
    f = np.asfortranarray(rng.standard_normal(n))
    for w in orc.FT_CODES:
        o = orc.fft_derivative(f, x2c, w)
        r = numpy_fft_derivative(f, x2c, w)
        assert np.abs(o - r).max() <= 1e-13 * np.abs(r).max(), w


def test_fft_laplacian_of_a_plane_wave_is_analytic():
    """lap(cos(G.x)) = -|G|^2 cos(G.x) for a reciprocal vector below the Nyquist index."""
    n = (16, 12, 10)
    x2c = np.array([[6.0, 0.5, 0.0], [0.0, 5.0, 0.4], [0.0, 0.0, 7.0]])
    B = 2 * np.pi * np.linalg.inv(x2c).T       # columns: true reciprocal vectors
    k = np.array([2, -1, 3])
    G = B @ k
    i, j, l = np.meshgrid(*[np.arange(m) / m for m in n], indexing="ij")
    phase = 2 * np.pi * (k[0] * i + k[1] * j + k[2] * l)
    f = np.asfortranarray(np.cos(phase))
    lap = orc.fft_derivative(f, x2c, "lap")
    assert np.abs(lap + (G @ G) * f).max() <= 1e-12 * (G @ G)
    gx = orc.fft_derivative(f, x2c, "x")
    assert np.abs(gx + G[0] * np.sin(phase)).max() <= 1e-12 * abs(G).max()
    gm = orc.fft_derivative(f, x2c, "grad")
    assert np.abs(gm - np.linalg.norm(G) * np.abs(np.sin(phase))).max() <= 1e-12 * np.linalg.norm(G)


def test_trilinear_reproduces_nodes_and_midpoints():
    rng = np.random.default_rng(3)
    f = np.asfortranarray(rng.standard_normal((6, 5, 7)))
    n = np.array(f.shape)
    for p in [(0, 0, 0), (5, 4, 6), (2, 3, 1)]:
        assert orc.grid_interp_trilinear(f, np.array(p) / n) == f[p]
    # midpoint of an x edge, wrapping around the cell
    v = orc.grid_interp_trilinear(f, np.array([5.5, 2.0, 3.0]) / n)
    assert abs(v - 0.5 * (f[5, 2, 3] + f[0, 2, 3])) <= 1e-15


def test_nci_fourier_node_aligned_is_pointwise():
    """On the node-aligned lattice every field%grd call takes the node short-cut (fieldmod@proc.f90:948-961)."""
    c = cases.make_case("cubic48")
    f, x2c = c["f"], c["x2c"]
    der = tuple(orc.fft_derivative(f, x2c, w) for w in ("grad", "xx", "yy", "zz"))
    crho, cgrad = orc.nci_rdg_fourier(f, x2c, derived=der)
    cst = 2.0 * (3.0 * np.pi ** 2) ** (1.0 / 3.0)
    s = der[0] / (cst * np.maximum(f, 1e-80) ** (4.0 / 3.0))
    assert np.abs(cgrad - s.transpose(2, 1, 0)).max() <= 1e-13 * np.abs(s).max()
    npos = (der[1] > 0).astype(int) + (der[2] > 0) + (der[3] > 0)
    sign = np.where(npos >= 2, 1.0, -1.0)
    assert np.array_equal(crho, (sign * np.abs(f) * 100.0).transpose(2, 1, 0))
