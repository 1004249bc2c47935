"""GPU parity of the formatted-text grid reader (c2g_grid_parse_text) against the oracle: BIT-EXACT values (correct
rounding, the result of a list-directed READ), both file orders, the volume division of CHGCAR files, the tokens that
take the exact multi-word path on the device (rounding-boundary cases, subnormals, long mantissas; nothing is converted on the host) and the error paths."""
import numpy as np
import pytest

from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def check(ctx, text, n, order, divisor=1.0):
    ref, end = orc.parse_text_grid(text, n, order, divisor)
    h, used, nhost = ctx.parse_text(text.encode(), n, order, divisor)
    out = ctx.download(h, n)
    ctx.free(h)
    assert np.array_equal(bits(out.ravel(order="F")), bits(ref.ravel(order="F"))), "values differ in some bit"
    assert used == end
    return nhost


def test_cube_block_k_fastest(ctx):
    rng = np.random.default_rng(5)
    n = (17, 9, 23)
    f = rng.standard_normal(n) * 10.0 ** rng.integers(-40, 4, n)   # vacuum-like tiny values included
    rows = []
    for i in range(n[0]):
        for j in range(n[1]):
            vals = ["%13.5E" % f[i, j, k] for k in range(n[2])]
            rows += ["".join(vals[q:q + 6]) for q in range(0, n[2], 6)]   # cube files: 6 per line, new line per (i,j)
    nhost = check(ctx, "\n".join(rows) + "\n", n, 1)
    assert nhost == 0


def test_chgcar_block_i_fastest_with_volume_and_trailer(ctx):
    rng = np.random.default_rng(6)
    n = (20, 18, 16)
    f = np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-3, 4, n)
    flat = f.ravel(order="F")
    body = "\n".join(" " + " ".join("%.11E" % v for v in flat[q:q + 5]) for q in range(0, flat.size, 5))
    text = body + "\naugmentation occupancies   1  33\n  0.1234567E+00 0.7654321E-01\n"
    nhost = check(ctx, text, n, 0, divisor=987.654321)
    assert nhost == 0


def test_number_forms_and_separators(ctx):
    toks = ["1", "-2", "+3.", ".5", "-.25e1", "1.5D-03", "2.5d+02", "1.5-03", "7q2", "1E0", "0", "-0.0", "000123.4500E-2",
            "9007199254740993", "123456789012345678", "1.0E22", "1.0E23", "4.9E-324", "2.2250738585072011E-308", "1.7976931348623157E308",
            "0.1E-30", "3.14159265358979323846264338327950288", "1e-400", "8.5E-300", "6.02214076E23"]
    n = (len(toks), 1, 1)
    text = " , ".join(toks[:5]) + "\t" + "\r\n".join(toks[5:12]) + "\n" + "  ".join(toks[12:]) + "\n"
    nhost = check(ctx, text, n, 0)
    assert nhost >= 3   # the subnormal, the > 19 digit mantissa and the underflow take the exact multi-word path


def test_rounding_boundary_tokens(ctx):
    """Decimal strings of exact midpoints between two doubles (and their neighbours): the double-double path must either
    be right or hand them to the exact multi-word comparison."""
    rng = np.random.default_rng(7)
    toks = []
    from fractions import Fraction
    for _ in range(200):
        x = float(rng.standard_normal() * 10.0 ** rng.integers(-20, 20))
        y = np.nextafter(x, np.inf)
        mid = (Fraction(x) + Fraction(y)) / 2
        # 19 significant digits of the midpoint (truncated) and of a value just above it
        e = int(np.floor(np.log10(abs(x)))) if x != 0 else 0
        scaled = mid / Fraction(10) ** (e - 18)
        m = int(scaled)
        for d in (0, 1):
            toks.append(f"{m + d}E{e - 18}")
    # exact midpoints with all their digits (ties: round to even) and 40-digit strings just around them
    for _ in range(100):
        x = float(abs(rng.standard_normal()) * 10.0 ** rng.integers(-15, 15))
        y = np.nextafter(x, np.inf)
        mid = (Fraction(x) + Fraction(y)) / 2
        num, den = mid.numerator, mid.denominator        # den is a power of two: the decimal expansion is finite
        k = den.bit_length() - 1
        digits = str(num * 5 ** k)                         # mid = digits * 10^-k
        if len(digits) <= 58:
            toks.append(f"{digits}E-{k}")
            toks.append(f"{digits}1E-{k + 1}")
            toks.append(f"{int(digits) - 1}9E-{k + 1}")
    toks += ["2.4703282292062327208e-324", "2.4703282292062327209e-324", "4.9406564584124654e-324", "1.7976931348623158079e308",
             "1.7976931348623158080e308", "1e-330", "1e400", "0.000000000000000000000000000000000000001e-300"]
    n = (len(toks), 1, 1)
    nslow = check(ctx, " ".join(toks) + "\n", n, 0)
    assert nslow > 100


def test_large_block_throughput_and_parity(ctx):
    rng = np.random.default_rng(8)
    n = (96, 96, 96)
    f = np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-12, 3, n)
    flat = f.ravel(order="F")
    lines = [" ".join("%.11E" % v for v in flat[q:q + 5]) for q in range(0, flat.size, 5)]
    text = "\n".join(lines) + "\n"
    h, used, nhost = ctx.parse_text(text.encode(), n, 0, 1.0)
    out = ctx.download(h, n)
    ctx.free(h)
    want = np.array(text.split(), dtype=np.float64).reshape(n, order="F")   # numpy's strtod: correctly rounded
    assert np.array_equal(bits(out.ravel(order="F")), bits(want.ravel(order="F")))
    assert nhost == 0 and used == len(text) - 1


def test_error_paths(ctx):
    with pytest.raises(capi.C2GError):
        ctx.parse_text(b"1.0 2.0 abc 4.0\n", (4, 1, 1))          # not a number
    with pytest.raises(capi.C2GError):
        ctx.parse_text(b"1.0 2.0 3.0\n", (4, 1, 1))              # too few values
    with pytest.raises(capi.C2GError):
        ctx.parse_text(b"1.0 2.0\n", (2, 1, 1), order=7)         # bad order
    with pytest.raises(capi.C2GError):
        ctx.parse_text(b"1.0 2.0\n", (2, 1, 1), divisor=0.0)
    h, used, _ = ctx.parse_text(b"1.0 2.0 junk\n", (2, 1, 1))   # text after the block is not looked at
    assert np.array_equal(ctx.download(h, (2, 1, 1)).ravel(), [1.0, 2.0])
    ctx.free(h)


# ---------------------------------------------------------------------------------------------------------------
# formatted output (c2g_grid_format_text) against the oracle's restatement of the Fortran edit descriptors: same bytes
# ---------------------------------------------------------------------------------------------------------------
def test_format_nci_cube_body(ctx):
    """(6(" ",1p,e13.5e3)) over crho(k,j,i): rows along the stored fastest index, negative values fit."""
    rng = np.random.default_rng(11)
    f = np.asfortranarray(rng.standard_normal((13, 5, 4)) * 10.0 ** rng.integers(-30, 6, (13, 5, 4)))
    f[0, 0, 0] = 0.0; f[1, 0, 0] = -0.0; f[2, 0, 0] = 9.999996; f[3, 0, 0] = 9.9999949999; f[4, 0, 0] = 1e-310; f[5, 0, 0] = 1.7976931348623157e308
    h = ctx.upload(f)
    got = ctx.format_text(h, 0, 13, 5, 1)
    assert got == orc.format_text_grid(f, 0, 13, 5, 1)
    ctx.free(h)


def test_format_writegrid_cube_formats_and_shift(ctx):
    """(1p,6(" ",E12.5E3)) (negative values overflow the field: asterisks, as the reference prints them) and the
    precisecube format (6(" ",E22.14E3)), cube order with ishift."""
    rng = np.random.default_rng(12)
    f = np.asfortranarray(rng.standard_normal((6, 7, 11)) * 10.0 ** rng.integers(-12, 4, (6, 7, 11)))
    h = ctx.upload(f)
    for (w, d, k) in ((12, 5, 1), (22, 14, 0)):
        for sh in (None, (2, 5, 3)):
            got = ctx.format_text(h, 1, w, d, k, ishift=sh)
            ref = orc.format_text_grid(f, 1, w, d, k, ishift=sh or (0, 0, 0))
            assert got == ref, (w, d, k, sh)
    assert b"************" in ctx.format_text(h, 1, 12, 5, 1)
    ctx.free(h)


def test_format_rounding_ties_and_round_trip(ctx):
    """Exact ties (dyadic values whose next digit is exactly 5) round to even; a precisecube block written by the
    formatter and read back by the parser reproduces its 14 significant digits."""
    vals = [2.0 ** -16, 0.5, 1.5, 2.5, 0.125, 1.0009765625, 3.0517578125e-05, 152587890625.0, 7.62939453125e-06]
    f = np.asfortranarray(np.array(vals + [0.0] * (16 - len(vals))).reshape(16, 1, 1))
    h = ctx.upload(f)
    for (w, d, k) in ((20, 10, 1), (13, 5, 1), (12, 2, 1), (22, 14, 0), (10, 1, 1)):
        assert ctx.format_text(h, 0, w, d, k) == orc.format_text_grid(f, 0, w, d, k), (w, d, k)
    ctx.free(h)
    rng = np.random.default_rng(13)
    g = np.asfortranarray(np.abs(rng.standard_normal((8, 9, 10))) * 10.0 ** rng.integers(-8, 8, (8, 9, 10)))
    hg = ctx.upload(g)
    text = ctx.format_text(hg, 1, 22, 14, 0)
    hb, used, _ = ctx.parse_text(text, (8, 9, 10), 1, 1.0)
    back = ctx.download(hb, (8, 9, 10))
    assert np.abs(back / g - 1.0).max() <= 5.1e-14
    ctx.free(hg); ctx.free(hb)


def test_text_golden_fixture(ctx):
    """The committed golden vectors (tests/golden/text_golden.json): reader bit patterns and writer fields."""
    import json, os, struct
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_golden.json")))
    toks, want = g["reader"]["tokens"], g["reader"]["bits"]
    h, _, _ = ctx.parse_text((" ".join(toks) + "\n").encode(), (len(toks), 1, 1), 0, 1.0)
    got = ctx.download(h, (len(toks), 1, 1)).ravel()
    ctx.free(h)
    assert ["%016x" % v for v in got.view(np.uint64)] == want
    vals = np.array([struct.unpack("<d", struct.pack("<Q", int(b, 16)))[0] for b in g["writer"]["values_bits"]])
    hv = ctx.upload(np.asfortranarray(vals.reshape(-1, 1, 1)))
    for key, fields in g["writer"]["fields"].items():
        w, d, k = (int(x) for x in key.split(","))
        text = ctx.format_text(hv, 0, w, d, k).decode()
        got_fields = [text[q * (w + 1) + q // 6 + 1: q * (w + 1) + q // 6 + 1 + w] for q in range(len(vals))]
        assert got_fields == fields, key
    ctx.free(hv)


def test_writer_against_the_reference_cube_files(ctx):
    """PINNED by critic2's own outputs (tests/golden/cube_golden.json, from the reference's nodata tests
    005_plot/013_cube_simple and 029_cube_precise): the device writer reproduces the value blocks of those cube files
    byte for byte -- fields, six values per line, and the blank that ends a partial line."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cube_golden.json")))
    for name, b in g["blocks"].items():
        vals = np.array(b["text"].split(), dtype=np.float64).reshape(1, b["rows"], b["n3"])
        h = ctx.upload(np.asfortranarray(vals))
        assert ctx.format_text(h, 1, 22, 14, 0).decode() == b["text"], name
        ctx.free(h)
    p = g["pairs"]
    f = np.asfortranarray(np.array(p["precise_text"].split(), dtype=np.float64).reshape(2, 2, 2))
    h = ctx.upload(f)
    assert ctx.format_text(h, 1, 22, 14, 0).decode() == p["precise_text"]
    assert ctx.format_text(h, 1, 12, 5, 1).decode() == p["standard_text"]
    ctx.free(h)
    fields = g["precise_fields"]
    hv = ctx.upload(np.asfortranarray(np.array(fields, dtype=np.float64).reshape(-1, 1, 1)))
    text = ctx.format_text(hv, 0, 22, 14, 0).decode()
    assert [text[q * 23 + q // 6 + 1: q * 23 + q // 6 + 23] for q in range(len(fields))] == fields
    ctx.free(hv)
