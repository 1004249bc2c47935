import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from critic2_b200 import capi
from oracle import oracle as orc
ctx = capi.Context(0)
for name in ("triclinic", "odd_dims", "cubic48"):
    c = cases.make_case(name)
    n = c["n"]; x2c = c["x2c"]
    crho_o, cgrad_o, lam2 = orc.nci_rdg(c["f"], x2c, want_lam2=True)
    h = ctx.upload(c["f"])
    crho, cgrad = ctx.nci_rdg(h, x2c, n)
    rel = np.abs(cgrad - cgrad_o) / np.maximum(np.abs(cgrad_o), 1e-300)
    print(name, "max rel", rel.max(), "frac>1e-12", (rel > 1e-12).mean(), "frac>1e-13", (rel > 1e-13).mean(), "frac==0", (rel == 0).mean(), "median s", np.median(cgrad_o))
    # is the lattice point exactly a node in the oracle's chain?
    c2x = np.linalg.inv(x2c)
    xmat = x2c / np.array(n, dtype=float)[None, :]
    worst = np.argsort(rel.ravel(order="F"))[-5:]
    for w in worst:
        k, j, i = np.unravel_index(w, rel.shape, order="F")
        x = ((0.0 + i * xmat[:, 0]) + j * xmat[:, 1]) + k * xmat[:, 2]
        wx = np.array([(c2x[d, 0] * x[0] + c2x[d, 1] * x[1]) + c2x[d, 2] * x[2] for d in range(3)])
        t = wx * np.array(n) - np.floor(wx * np.array(n))
        print("   kji", (k, j, i), "s_gpu", cgrad[k, j, i], "s_ref", cgrad_o[k, j, i], "rel", rel[k, j, i], "t", t, "rho", crho_o[k, j, i] / 100)
    offnode = 0
    ctx.free(h)
ctx.close()
