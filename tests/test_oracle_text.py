"""CPU tests of the oracle's formatted-text grid reader (read_cube / read_vasp numeric blocks) -- no GPU."""
import numpy as np

from oracle import oracle as orc


def test_fortran_float_forms():
    assert orc.fortran_float("1.5E-03") == 1.5e-3
    assert orc.fortran_float("1.5D-03") == 1.5e-3
    assert orc.fortran_float("1.5-03") == 1.5e-3
    assert orc.fortran_float("-.25q+2") == -25.0
    assert orc.fortran_float("+7") == 7.0
    assert orc.fortran_float("0.12345678901E+03") == 123.45678901


def test_cube_and_vasp_orders_round_trip():
    rng = np.random.default_rng(1)
    n = (3, 4, 5)
    f = np.asfortranarray(rng.standard_normal(n) * 10.0 ** rng.integers(-8, 3, n))
    # Gaussian cube: k fastest, 6 values per line, %13.5E
    cube = ""
    for i in range(n[0]):
        for j in range(n[1]):
            row = ["%13.5E" % f[i, j, k] for k in range(n[2])]
            cube += "".join(row) + "\n"
    g, end = orc.parse_text_grid(cube, n, order=1)
    assert end <= len(cube)
    want = np.array([[[float("%13.5E" % f[i, j, k]) for k in range(n[2])] for j in range(n[1])] for i in range(n[0])])
    assert np.array_equal(g, want)
    # CHGCAR: i fastest, E18.11, divided by the cell volume
    flat = f.ravel(order="F")
    chg = "\n".join(" ".join("%18.11E" % v for v in flat[q:q + 5]) for q in range(0, flat.size, 5)) + "\naugmentation occupancies 1 2\n"
    g2, end2 = orc.parse_text_grid(chg, n, order=0, divisor=123.456)
    want2 = np.array([float("%18.11E" % v) for v in flat]).reshape(n, order="F") / 123.456
    assert np.array_equal(g2, want2)
    assert chg[end2:].lstrip().startswith("augmentation")
