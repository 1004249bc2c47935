"""Extracts golden vectors for the cube writer from the reference's OWN test outputs (tests/005_plot/ref/*.cube, the
`nodata` tests 013_cube_simple and 029_cube_precise, which critic2's CI runs): value fields written by
writegrid_cube with (6(" ",E22.14E3)) (precisecube) and (1p,6(" ",E12.5E3)) (standardcube), and the raw value blocks
incl. their line structure.  Run in the build container (needs /root/reference); writes tests/golden/cube_golden.json."""
import json
import os

REF = "/root/reference/tests/005_plot/ref"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cube_golden.json")


def body(path):
    lines = open(path).read().split("\n")
    nat = int(lines[2].split()[0])
    n = [int(lines[3 + k].split()[0]) for k in range(3)]
    return n, lines[6 + nat:]


gold = {"source": "critic2 tests/005_plot/ref (013_cube_simple_04/09/14.cube, 029_cube_precise_01/02.cube)", "precise_fields": [],
        "blocks": {}}
for name in ("013_cube_simple_04.cube", "013_cube_simple_09.cube", "013_cube_simple_14.cube"):
    n, lines = body(os.path.join(REF, name))
    fields = []
    for ln in lines:
        fields += [ln[k + 1:k + 23] for k in range(0, len(ln) - 22, 23)]
    gold["precise_fields"] += fields[:400]
    # the first 4 rows (n3 = 10 values each: a line of 6 and a line of 4 that ends with a blank) as raw text
    gold["blocks"][name] = {"n3": n[2], "rows": 4, "text": "\n".join(lines[:8]) + "\n"}
n1, l1 = body(os.path.join(REF, "029_cube_precise_01.cube"))
n2, l2 = body(os.path.join(REF, "029_cube_precise_02.cube"))
gold["pairs"] = {"n": n1, "standard_text": "\n".join(l1[:4]) + "\n", "precise_text": "\n".join(l2[:4]) + "\n"}
json.dump(gold, open(OUT, "w"), indent=0)
print("wrote", OUT, len(gold["precise_fields"]), "fields")
