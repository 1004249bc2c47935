"""Small driver for ncu: one 256^3 (or N^3) BADER assign + integrate on a GPU-generated density."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import systems as S
from critic2_b200 import capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
side = int(sys.argv[2]) if len(sys.argv) > 2 else max(2, N // 64)
algo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ctx = capi.Context(0)
ctx.profile_enable(True)
n = (N, N, N)
x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
at, z, al = S.jittered_lattice(side, 5)
at = S.snap_to_grid(at, n)
h = ctx.alloc(n)
ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
h2 = ctx.alloc(n)   # second INTEGRABLE field, like bench.py (P_f = 2: 20 B/pt in the reduction)
ctx.promolecular(h2, x2c, at, z * 0.5, al * 1.3, nimg=1, rc=8.0)
lat2car = x2c / np.array(n, dtype=float)[None, :]
car2lat = np.linalg.inv(lat2car)
lid = np.zeros((3, 3, 3))
for i in (-1, 0, 1):
    for j in (-1, 0, 1):
        for k in (-1, 0, 1):
            if (i, j, k) != (0, 0, 0):
                lid[i + 1, j + 1, k + 1] = 1.0 / np.linalg.norm(lat2car @ np.array([i, j, k], float))
for rep in range(reps):
    ctx.profile_reset()
    t = time.time()
    b = ctx.bader_assign(h, car2lat, lid, algo=algo)
    b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
    vol, ps = ctx.integrate(b, [h, h2], abs(np.linalg.det(x2c)))
    dt = time.time() - t
    prof = ctx.profile()
    print(f"N={N} algo={algo} rep={rep} nmax={b.nmax} wall={dt*1e3:.2f} ms  sum(pop)={ps[:,0].sum():.6f} stats={b.stats()[:6]}")
    print("  ", {k: round(v[0], 3) for k, v in prof.items()}, " kernels total %.3f ms" % sum(v[0] for v in prof.values()))
    b.free()
ctx.close()
