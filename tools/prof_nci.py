"""Small driver for ncu: one NCIPLOT RDG pass (tricubic, node-aligned) and the FOURIER variant on an N^3 grid field."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import systems as S
from critic2_b200 import capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = (N, N, N)
ctx = capi.Context(0)
x2c = S.cell_x2c(30.0, 30.0, 30.0)
at, z, al = S.random_atoms(24, 3, x2c, dmin=2.0)
h = ctx.alloc(n); ctx.promolecular(h, x2c, S.snap_to_grid(at, n), z, al, nimg=0, rc=0.0)
for rep in range(2):
    hr, hg = ctx.nci_rdg_resident(h, x2c, n)
    ctx.free(hr); ctx.free(hg)
hd = [ctx.fft_derivative(h, x2c, w) for w in ("grad", "xx", "yy", "zz")]
crho, cgrad = ctx.nci_rdg_fourier([h] + hd, x2c, n, nstep=(64, 64, 64))
print("ok", float(cgrad.max()))
ctx.close()
