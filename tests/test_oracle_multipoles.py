"""Oracle checks for the basin multipoles (integration@proc.f90:1302-1361; genrlm_real, tools_math@proc.f90:273-306).

Pinned by data the reference itself contains: the comment block of genrlm_real lists the order and the closed
forms of the first nine real regular solid harmonics (C00 = 1, C11 = x, C10 = z, S11 = y, C22 = sqrt(3)/2 (x^2-y^2),
C21 = sqrt(3) x z, C20 = (3 z^2 - r^2)/2, S21 = sqrt(3) y z, S22 = sqrt(3) x y).  Higher l is checked against an
independent implementation (scipy's complex spherical harmonics) and against harmonicity."""
import numpy as np
import pytest

import cases
import systems as S
from oracle import oracle as orc


def closed_forms(v):
    x, y, z = v
    r2 = x * x + y * y + z * z
    s3 = np.sqrt(3.0)
    return np.array([1.0, x, z, y, s3 / 2 * (x * x - y * y), s3 * x * z, 0.5 * (3 * z * z - r2), s3 * y * z, s3 * x * y])


def test_first_nine_match_the_reference_comment_block():
    rng = np.random.default_rng(1)
    for v in list(rng.normal(size=(50, 3))) + [np.array([0.0, 0.0, 1.3]), np.array([0.0, 0.0, -0.4]), np.zeros(3),
                                                 np.array([1.0, 0.0, 0.0]), np.array([0.0, -2.0, 0.0])]:
        got = orc.rlm_real(v, 2)
        assert np.abs(got - closed_forms(v)).max() <= 4e-15 * max(1.0, v @ v)


def test_higher_l_against_scipy_spherical_harmonics():
    sp = pytest.importorskip("scipy.special")
    rng = np.random.default_rng(2)
    lmax = 8
    for v in rng.normal(size=(20, 3)):
        r = np.linalg.norm(v)
        th, ph = np.arccos(v[2] / r), np.arctan2(v[1], v[0])
        got = orc.rlm_real(v, lmax)
        for l in range(lmax + 1):
            for m in range(0, l + 1):
                if hasattr(sp, "sph_harm_y"):
                    y = sp.sph_harm_y(l, m, th, ph)
                else:
                    y = sp.sph_harm(m, l, ph, th)
                R = np.sqrt(4 * np.pi / (2 * l + 1)) * r ** l * y
                if m == 0:
                    want = [(l * l + l, R.real)]
                else:  # C_lm = sqrt(2) (-1)^m Re R_lm, S_lm = sqrt(2) (-1)^m Im R_lm
                    want = [(l * l + l - m, np.sqrt(2) * (-1) ** m * R.real), (l * l + l + m, np.sqrt(2) * (-1) ** m * R.imag)]
                for idx, w in want:
                    assert abs(got[idx] - w) <= 1e-13 * max(1.0, r ** l), (l, m)


def test_shortest_and_multipoles_of_a_point_charge_grid():
    """A field that is 1 at one node and 0 elsewhere: the multipoles of its basin are the solid harmonics of the
    shortest vector from the attractor to that node, times omega/ntot -- orthogonal and triclinic cells."""
    for cellp, n in (((6.0, 7.0, 8.0, 90, 90, 90), (8, 9, 10)), ((6.0, 7.0, 8.0, 80, 95, 105), (8, 9, 10))):
        x2c = S.cell_x2c(*cellp)
        vec, _ = S.wscell(x2c)
        cell = orc.Cell(x2c, ws=(x2c @ vec.T.astype(float)))
        idg = np.ones(n, dtype=np.int32, order="F")
        xattr = np.array([[0.1], [0.95], [0.45]])
        for node in ((7, 0, 4), (1, 8, 9), (4, 4, 0)):
            f = np.zeros(n, order="F")
            f[node] = 1.0
            mp = orc.multipoles_bader(idg, xattr, 4, f, cell, S.omega(x2c))[:, 0]
            dx = np.array(node, dtype=float) / np.array(n) - xattr[:, 0]
            best = min((x2c @ (dx + np.array([a, b, c])) for a in range(-2, 3) for b in range(-2, 3) for c in range(-2, 3)),
                       key=np.linalg.norm)
            want = orc.rlm_real(best, 4) * S.omega(x2c) / f.size
            assert np.abs(mp - want).max() <= 1e-13 * np.abs(want).max()


def test_monopole_equals_the_population_and_yt_weights_partition():
    c = cases.make_case("cubic48")
    idg, nattr, xattr, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    om = S.omega(c["x2c"])
    mp = orc.multipoles_bader(idg, xattr, 3, c["f"], orc.Cell(c["x2c"]), om)
    _, ps = orc.integrate_bader(idg, [c["f"]], nattr, om)
    assert np.abs(mp[0] - ps[:, 0]).max() <= 1e-12 * np.abs(ps[:, 0]).max()
    # dipoles of a near-spherical atomic basin are small against r * population
    assert np.abs(mp[1:4]).max() < 0.5 * np.abs(mp[0]).max()
    # YT: the sum over basins of the weighted monopoles is the grid integral
    vec, area = S.wscell(c["x2c"] / np.array(c["n"], dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], c["x2c"], vec, area, atoms=c["atoms"])
    tot = 0.0
    for m in range(1, d.nattr + 1):
        w = orc.yt_weights(d, m, c["n"])
        tot += orc.multipoles_weighted(w, d.xattr[:, m - 1], 2, c["f"], orc.Cell(c["x2c"]), om)[0]
    assert abs(tot - c["f"].sum() * om / c["f"].size) <= 1e-11 * tot


def _rlm_device_formulation(v, lmax):
    """Python port of add_point() in critic2_b200/csrc/multipole.cu: the kernel takes sin/cos of the angles of
    tosphere directly from the vector (cos(theta) = z/r, sin(theta) = rho/r, cos/sin(phi) = x/rho, y/rho, angle
    addition for m phi) instead of acos / atan2 / sin / cos; the recursion is genylm's."""
    pi, eps, sh = np.pi, 1e-14, 1 / np.sqrt(2)
    r = np.sqrt(v @ v)
    out = np.zeros((lmax + 1) ** 2)
    out[0] = 0.28209479177387814347 * np.sqrt(4 * pi)
    sn, cs, c1, s1 = 0.0, 1.0, 1.0, 0.0
    if r > eps:
        t1 = v[2] / r
        rho = np.sqrt(v[0] ** 2 + v[1] ** 2)
        if t1 >= 1:
            cs, sn = 1.0, 0.0
        elif t1 <= -1:
            cs, sn = -1.0, 1.2246467991473532e-16
        else:
            cs, sn = t1, rho / r
        if abs(v[0]) > eps or abs(v[1]) > eps:
            c1, s1 = v[0] / rho, v[1] / rho
    zc, zs = [0.0, c1] + [0.0] * lmax, [0.0, s1] + [0.0] * lmax
    for m in range(2, lmax + 1):
        zc[m] = zc[m - 1] * c1 - zs[m - 1] * s1
        zs[m] = zs[m - 1] * c1 + zc[m - 1] * s1
    x = [0.0] * (lmax + 1)
    for l in range(1, lmax + 1):
        x[l] = -1.0 if l & 1 else 1.0
        dx = 0.0
        for m in range(l, 0, -1):
            t1 = np.sqrt((l + m) * (l - m + 1))
            x[m - 1] = -(sn * dx + 2 * m * cs * x[m]) / t1
            dx = sn * x[m] * t1
        t1, s = sn, 0.0
        for m in range(1, l + 1):
            x[m] = t1 * x[m]
            s += x[m] ** 2
            t1 *= sn
        s = 2 * s + x[0] ** 2
        t1 = np.sqrt((2 * l + 1) / (4 * pi * s))
        sc, rl = np.sqrt(4 * pi / (2 * l + 1)), r ** l
        out[l * l + l] = t1 * x[0] * sc * rl
        for m in range(1, l + 1):
            a = t1 * x[m]
            ph = -1.0 if m & 1 else 1.0
            out[l * l + l - m] = sh * 2 * ph * (a * zc[m] * sc * rl)
            out[l * l + l + m] = sh * 2 * ph * (a * zs[m] * sc * rl)
    return out


def test_device_formulation_of_the_harmonics_matches_the_angle_form_on_grid_vectors():
    """The kernel's transcendental-free formulation against the oracle's literal acos/atan2/sin/cos restatement on the
    vectors a grid produces (multiples of 1/N of the cell, incl. points on the axes and in the coordinate planes):
    <= 1e-12 of r^l.  Off the grid the two differ near the poles, where theta = acos(z/r) loses digits (reported)."""
    rng = np.random.default_rng(3)
    lmax, N = 10, 128
    vs = [rng.integers(-N // 2, N // 2, 3) / N * np.array([9.0, 10.0, 11.0]) for _ in range(1500)]
    vs += [np.array(a, float) for a in [(0, 0, 1), (0, 0, -2), (1, 0, 0), (0, -1, 0), (0, 0, 0), (3, 4, 0), (0, 2, 2), (1 / N, 0, 5)]]
    worst = 0.0
    for v in vs:
        r = np.sqrt(v @ v)
        sc = np.array([max(r, 1.0) ** l if r == 0.0 else r ** l for l in range(lmax + 1) for _ in range(2 * l + 1)])
        worst = max(worst, (np.abs(_rlm_device_formulation(v, lmax) - orc.rlm_real(v, lmax)) / sc).max())
    assert worst <= 1e-12, worst
    v = np.array([0.0, 1e-7, -1.0])                                    # not a grid vector: acos is ill-conditioned here
    d = np.abs(_rlm_device_formulation(v, 4) - orc.rlm_real(v, 4)).max()
    print(f"near-pole difference between the two formulations (angle form loses digits): {d:.2e}")
    assert d < 1e-7
