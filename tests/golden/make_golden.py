"""Generates tests/golden/oracle_golden.json with the CPU oracle (oracle/oracle.cpp).

The reference's own golden files (tests/009_intgrid/ref/*.cro) cannot be used: their input grids
(tests/zz_source/...) are not shipped and critic2 cannot be compiled here (no Fortran compiler).
These fixtures therefore pin the ORACLE (and, through it, the CUDA path) on the seeded cases of
tests/cases.py.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import systems as S  # noqa: E402
from oracle import oracle as orc  # noqa: E402

out = {"bader": {}, "yt": {}, "nci": {}}
for name in ("cubic48", "triclinic", "odd_dims"):
    c = cases.make_case(name)
    idg, nattr, xattr, stats = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    vol, ps = orc.integrate_bader(idg, [c["f"]], nattr, S.omega(c["x2c"]))
    out["bader"][name] = {
        "nattr": int(nattr),
        "labels_sha256": hashlib.sha256(np.ascontiguousarray(idg.ravel(order="F")).tobytes()).hexdigest(),
        "pop": ps[:, 0].tolist(), "vol": vol.tolist(), "refine_iterations": int(stats[0]),
    }
for name in ("cubic48", "triclinic"):
    c = cases.make_case(name)
    vec, area = S.wscell(c["x2c"] / np.array(c["n"], dtype=float)[None, :])
    d = orc.yt_integrate(c["f"], c["x2c"], vec, area, atoms=c["atoms"])
    vol, ps = orc.integrate_yt(d, [c["f"]], S.omega(c["x2c"]))
    sb = d.spatial_basin(c["n"])
    out["yt"][name] = {
        "nattr": int(d.nattr), "nvec": int(len(area)),
        "labels_sha256": hashlib.sha256(np.ascontiguousarray(sb.ravel(order="F")).tobytes()).hexdigest(),
        "pop": ps[:, 0].tolist(), "vol": vol.tolist(),
    }
c = cases.make_case("triclinic")
crho, cgrad = orc.nci_rdg(c["f"], c["x2c"])
idx = [(0, 0, 0), (5, 7, 11), (71, 67, 63), (33, 20, 10), (1, 2, 3)]
out["nci"]["triclinic"] = {"points_kji": idx, "crho": [float(crho[i]) for i in idx], "cgrad": [float(cgrad[i]) for i in idx],
                           "sum_cgrad": float(cgrad.sum()), "n_negative": int((crho < 0).sum())}
json.dump(out, open(os.path.join(HERE, "oracle_golden.json"), "w"), indent=1)
print("written", os.path.join(HERE, "oracle_golden.json"))
