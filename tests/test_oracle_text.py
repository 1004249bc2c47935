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


def test_fortran_e_descriptor():
    assert orc.fortran_e(1.234565, 13, 5, 1) == " 1.23456E+000" or orc.fortran_e(1.234565, 13, 5, 1) == " 1.23457E+000"
    assert orc.fortran_e(1.0, 13, 5, 1) == " 1.00000E+000"
    assert orc.fortran_e(-1.0, 13, 5, 1) == "-1.00000E+000"
    assert orc.fortran_e(-1.0, 12, 5, 1) == "*" * 12            # 13 characters do not fit in 1p,E12.5E3
    assert orc.fortran_e(1.0, 12, 5, 1) == "1.00000E+000"
    assert orc.fortran_e(0.0, 12, 5, 1) == "0.00000E+000"
    assert orc.fortran_e(9.999996, 13, 5, 1) == " 1.00000E+001"   # the carry moves the exponent
    assert orc.fortran_e(0.5, 13, 0 + 5, 1) == " 5.00000E-001"
    assert orc.fortran_e(123.456, 22, 14, 0) == " 0.12345600000000E+003"
    assert orc.fortran_e(-123.456, 22, 14, 0) == "-0.12345600000000E+003"
    assert orc.fortran_e(2.5e-310, 22, 14, 0) == " 0.25000000000000E-309"
    assert orc.fortran_e(float("inf"), 13, 5, 1) == "     Infinity"
    assert orc.fortran_e(float("nan"), 13, 5, 1) == "          NaN"
    # exact tie: 2^-16 = 1.52587890625E-05 printed with 10 digits after the point -> ...789062|5 rounds to even
    assert orc.fortran_e(2.0 ** -16, 20, 10, 1) == "   1.5258789062E-005"


def test_format_text_grid_layouts():
    rng = np.random.default_rng(2)
    f = np.asfortranarray(rng.standard_normal((4, 3, 7)))
    t0 = orc.format_text_grid(f, 0, 13, 5, 1).decode()
    assert t0.count("\n") == 3 * 7 * 1 and len(t0) == 3 * 7 * (4 * 14 + 1 + 1)      # 4 < 6 values: blank + new line
    t1 = orc.format_text_grid(np.abs(f), 1, 12, 5, 1, ishift=(1, 0, 2)).decode()
    assert t1.count("\n") == 4 * 3 * 2 and len(t1) == 4 * 3 * (7 * 13 + 2 + 1)  # lines of 6 and 1 values: the short one ends with a blank
    first = t1.split()[0]
    assert first == orc.fortran_e(float(abs(f[1, 0, 2])), 12, 5, 1).strip()


def test_text_golden_fixture_matches_the_oracle():
    """tests/golden/text_golden.json (made by tests/golden/make_text_golden.py) against the oracle as it is now."""
    import json, os, struct
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_golden.json")))
    for t, b in zip(g["reader"]["tokens"], g["reader"]["bits"]):
        assert "%016x" % struct.unpack("<Q", struct.pack("<d", orc.fortran_float(t)))[0] == b, t
    vals = [struct.unpack("<d", struct.pack("<Q", int(b, 16)))[0] for b in g["writer"]["values_bits"]]
    for key, fields in g["writer"]["fields"].items():
        w, d, k = (int(x) for x in key.split(","))
        assert [orc.fortran_e(v, w, d, k) for v in vals] == fields


def _cube_gold():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cube_golden.json")))


def test_writer_against_the_reference_cube_files():
    """PINNED: tests/golden/cube_golden.json holds value fields and raw value blocks of cube files written by critic2
    itself (its nodata tests 005_plot/013_cube_simple and 029_cube_precise; extracted by tests/golden/make_cube_golden.py).
    The restated edit descriptors must reproduce them byte for byte: the E22.14E3 fields (incl. the sign and the leading
    zero), the 1P E12.5E3 fields, six values per line and the blank that ends a partial line."""
    g = _cube_gold()
    for fld in g["precise_fields"]:                       # 1200 fields: 14 digits -> double -> the same 14 digits
        assert orc.fortran_e(float(fld), 22, 14, 0) == fld
    for name, b in g["blocks"].items():                   # 4 rows of 10 values: lines of 6 + 4, cube order (layout 1)
        vals = np.array(b["text"].split(), dtype=np.float64).reshape(1, b["rows"], b["n3"])
        assert orc.format_text_grid(np.asfortranarray(vals), 1, 22, 14, 0).decode() == b["text"], name
    p = g["pairs"]                                        # the same 2x2x2 grid written by standardcube and precisecube
    vals = np.array(p["precise_text"].split(), dtype=np.float64).reshape(2, 2, 2)     # (ix, iy, iz), iz fastest in the file
    f = np.asfortranarray(vals)
    assert orc.format_text_grid(f, 1, 22, 14, 0).decode() == p["precise_text"]
    assert orc.format_text_grid(f, 1, 12, 5, 1).decode() == p["standard_text"]


def test_writer_shift_and_nci_layout_against_the_reference_files():
    """PINNED by critic2's outputs: (a) the same 10x10x10 grid written by `cube grid` plainly and with `shift 4 4 4`
    (tests/005_plot/016_cube_grid) -- the ishift indexing of writegrid_cube (crystalmod@write.f90:3556-3565);
    (b) the grad / dens cubes of NCIPLOT's write_cube_body, (6(" ",1p,e13.5e3)) (tests/015_grdplot/005_nciplot_basic)."""
    g = _cube_gold()
    s = g["shift"]
    f = np.asfortranarray(np.array(s["plain_text"].split(), dtype=np.float64).reshape(s["n"]))   # (ix, iy, iz), iz fastest
    assert orc.format_text_grid(f, 1, 22, 14, 0).decode() == s["plain_text"]
    assert orc.format_text_grid(f, 1, 22, 14, 0, ishift=tuple(s["ishift"])).decode() == s["shifted_text"]
    # and back: the reader on the reference's own text (read_cube's loop order: k fastest)
    back, used = orc.parse_text_grid(s["shifted_text"].encode(), tuple(s["n"]), 1, 1.0)
    sh = np.roll(f, shift=(-4, -4, -4), axis=(0, 1, 2))
    assert np.array_equal(back, sh) and used <= len(s["shifted_text"])
    for key in ("nci_grad", "nci_dens"):
        b = g[key]
        c = np.array(b["text"].split(), dtype=np.float64).reshape(b["n"][2], b["n"][1], b["n"][0], order="F")  # c(k,j,i)
        assert orc.format_text_grid(c, 0, 13, 5, 1).decode() == b["text"], key


def test_reader_orders_and_volume_scaling_against_the_reference_files():
    """PINNED by critic2's outputs: tests/005_plot/017_cube_files writes ONE 10x10x10 grid as a cube file (k fastest) and
    as a CHGCAR (values times the cell volume, i fastest).  Read back with read_cube's and read_vasp's conventions
    (grid3mod@proc.f90:559, :884-908: order and the division by det3(x2c)) the two must be the same field."""
    g = _cube_gold()
    n = tuple(g["chgcar"]["n"])
    a, b, c = g["chgcar"]["cell_bohr"]
    fc, _ = orc.parse_text_grid(g["shift"]["plain_text"].encode(), n, 1, 1.0)           # the cube block (016 = 017: same grid)
    fv, used = orc.parse_text_grid(g["chgcar"]["text"].encode(), n, 0, a * b * c)      # the CHGCAR block
    assert used <= len(g["chgcar"]["text"])
    assert np.abs(fv / fc - 1.0).max() <= 1e-12                                        # both files carry 14 digits
