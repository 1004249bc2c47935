"""FAST vs EXACT label comparison on a GPU-generated density (N^3, side^3 atoms); prints mismatches and timings."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import systems as S
from critic2_b200 import capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
side = int(sys.argv[2]) if len(sys.argv) > 2 else max(2, N // 128)
ctx = capi.Context(0)
n = (N, N, N)
x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
at, z, al = S.jittered_lattice(side, 5)
at = S.snap_to_grid(at, n)
h = ctx.alloc(n)
ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
lat2car = x2c / np.array(n, dtype=float)[None, :]
car2lat = np.linalg.inv(lat2car)
lid = np.zeros((3, 3, 3))
for i in (-1, 0, 1):
    for j in (-1, 0, 1):
        for k in (-1, 0, 1):
            if (i, j, k) != (0, 0, 0):
                lid[i + 1, j + 1, k + 1] = 1.0 / np.linalg.norm(lat2car @ np.array([i, j, k], float))
labs = []
for algo in (capi.BADER_EXACT, capi.BADER_FAST):
    t = time.time()
    b = ctx.bader_assign(h, car2lat, lid, algo=algo)
    ctx.synchronize()
    dt = time.time() - t
    b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
    labs.append(b.labels(n))
    print(f"N={N} algo={algo} nmax={b.nmax} wall={dt*1e3:.1f} ms stats={b.stats()[:7]}", flush=True)
    b.free()
mism = int(np.count_nonzero(labs[0] != labs[1]))
print(f"N={N}: FAST vs EXACT label mismatches: {mism} of {N**3}")
ctx.close()
sys.exit(1 if mism else 0)
