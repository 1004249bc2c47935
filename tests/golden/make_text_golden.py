"""Golden vectors for the formatted-text reader / writer (tests/golden/text_golden.json).

Reader: decimal tokens -> IEEE-754 bit patterns, from Python's float() (correctly rounded, David Gay's algorithm),
i.e. independent of the oracle's restatement.  Writer: doubles (as bit patterns) -> Ew.dE3 fields, from
oracle.fortran_e (Python decimal + the Fortran edit-descriptor rules; not pinned by a Fortran compiler).
Run from the repository root:  python tests/golden/make_text_golden.py"""
import json, os, struct, sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc

rng = np.random.default_rng(20261017)
toks = ["0", "-0.0", "1", "-2.5", ".5", "1.5D-03", "1.5-03", "7q2", "9007199254740993", "123456789012345678901234567890",
        "4.9406564584124654e-324", "2.4703282292062327208e-324", "2.4703282292062327209e-324", "2.2250738585072011e-308",
        "1.7976931348623157e308", "1.7976931348623158079e308", "1.7976931348623158080e308", "1e-400", "1e400", "6.02214076E23"]
for _ in range(150):   # typical cube / CHGCAR fields
    toks.append("%13.5E" % float(rng.standard_normal() * 10.0 ** rng.integers(-40, 6)))
    toks.append("%.11E" % float(abs(rng.standard_normal()) * 10.0 ** rng.integers(-6, 6)))
for _ in range(60):    # full expansions of midpoints between adjacent doubles (ties) and their neighbours
    x = float(abs(rng.standard_normal()) * 10.0 ** rng.integers(-12, 12))
    mid = (Fraction(x) + Fraction(float(np.nextafter(x, np.inf)))) / 2
    k = mid.denominator.bit_length() - 1
    digits = str(mid.numerator * 5 ** k)
    if len(digits) <= 56:
        toks += [f"{digits}E-{k}", f"{digits}1E-{k + 1}", f"{int(digits) - 1}9E-{k + 1}"]
toks = [t.strip() for t in toks]
bits = ["%016x" % struct.unpack("<Q", struct.pack("<d", orc.fortran_float(t)))[0] for t in toks]

vals = [0.0, -0.0, 1.0, -1.0, 9.999996, 9.9999949999, 0.5, 2.0 ** -16, 1e-310, 4.9406564584124654e-324, 1.7976931348623157e308,
        123.456, -123.456, 3.0517578125e-05, float("inf"), float("-inf")]
vals += [float(rng.standard_normal() * 10.0 ** rng.integers(-30, 30)) for _ in range(120)]
fmt = {}
for (w, d, k) in ((13, 5, 1), (12, 5, 1), (22, 14, 0)):
    fmt[f"{w},{d},{k}"] = [orc.fortran_e(v, w, d, k) for v in vals]
out = {"reader": {"tokens": toks, "bits": bits},
       "writer": {"values_bits": ["%016x" % struct.unpack("<Q", struct.pack("<d", v))[0] for v in vals], "fields": fmt}}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "text_golden.json"), "w"), indent=0)
print(len(toks), "tokens,", len(vals), "values")
