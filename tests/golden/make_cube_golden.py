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


gold = {"source": "critic2 tests/005_plot/ref (013_cube_simple_04/09/14.cube, 029_cube_precise_01/02.cube, 016_cube_grid_01/02.cube), tests/015_grdplot/ref (005_nciplot_basic-gen-grad/dens.cube)", "precise_fields": [],
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
# the same 10x10x10 grid written plainly and with `shift 4 4 4` (005_plot/016_cube_grid): pins ishift of writegrid_cube
ng, lg1 = body(os.path.join(REF, "016_cube_grid_01.cube"))
_, lg2 = body(os.path.join(REF, "016_cube_grid_02.cube"))
gold["shift"] = {"n": ng, "ishift": [4, 4, 4], "plain_text": "\n".join(lg1[:2 * ng[0] * ng[1]]) + "\n",
                 "shifted_text": "\n".join(lg2[:2 * ng[0] * ng[1]]) + "\n"}
# NCIPLOT's write_cube_body, (6(" ",1p,e13.5e3)): 015_grdplot/005_nciplot_basic (nstep 2 2 2)
NCI = "/root/reference/tests/015_grdplot/ref"
for tag in ("grad", "dens"):
    nn, ln = body(os.path.join(NCI, "005_nciplot_basic-gen-%s.cube" % tag))
    gold["nci_" + tag] = {"n": nn, "text": "\n".join(ln[:nn[0] * nn[1]]) + "\n"}
# the same grid written as a CHGCAR (005_plot/017_cube_files): values times the cell volume, index 1 fastest, 5 per line
chg = open(os.path.join(REF, "017_cube_files.CHGCAR")).read().split("\n")
k0 = [i for i, l in enumerate(chg) if l.split() == ["10", "10", "10"]][0]
gold["chgcar"] = {"n": [10, 10, 10], "cell_bohr": [10.51632592951, 10.51632592951, 8.85147720644],
                  "text": "\n".join(chg[k0 + 1:k0 + 1 + 200]) + "\n"}
json.dump(gold, open(OUT, "w"), indent=0)
print("wrote", OUT, len(gold["precise_fields"]), "fields")
