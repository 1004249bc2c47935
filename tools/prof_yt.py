"""Small driver for ncu: one YT build + integrate on an N^3 many-basin density."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import systems as S
from critic2_b200 import capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
side = max(2, N // 64)
n = (N, N, N)
ctx = capi.Context(0)
x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
at, z, al = S.jittered_lattice(side, 4)
h = ctx.alloc(n); ctx.promolecular(h, x2c, S.snap_to_grid(at, n), z, al, nimg=1, rc=8.0)
vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
b = ctx.yt_build(h, vec, area)
b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
vol, ps = ctx.integrate(b, [h], S.omega(x2c))
print("ok", b.nmax, b.stats()[:3], float(vol.sum() / S.omega(x2c)))
ctx.close()
