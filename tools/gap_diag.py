"""Where do host-side gaps in the bench step come from?  usage: gap_diag.py [torch] [f2]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if "torch" in sys.argv:
    import torch
    torch.cuda.set_device(0)
import bench
from critic2_b200 import capi
ctx = capi.Context(0)
n, x2c, at, z, al, side = bench.workload(1024)
car2lat, lid = bench.bader_metrics(x2c, n)
omega = abs(np.linalg.det(x2c))
h = ctx.alloc(n); ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
h2 = h
if "f2" in sys.argv:
    h2 = ctx.alloc(n); ctx.promolecular(h2, x2c, at, z * 0.5, al * 1.3, nimg=1, rc=8.0)
if "prof" in sys.argv:
    ctx.profile_enable(True)
ident = None
ts = []
if "timer" in sys.argv:
    ctx.timer_start()
for rep in range(10):
    if "nosync" not in sys.argv: ctx.synchronize()
    t0 = time.perf_counter()
    b = ctx.bader_assign(h, car2lat, lid)
    t1 = time.perf_counter()
    if ident is None: ident = np.arange(1, b.nmax + 1, dtype=np.int32)
    b.set_map(b.nmax, ident)
    vol, ps = ctx.integrate(b, [h, h2], omega)
    t2 = time.perf_counter()
    b.free()
    if "nosync" not in sys.argv: ctx.synchronize()
    t3 = time.perf_counter()
    ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
print(sys.argv[1:], " | ".join(f"{a:.1f}+{b:.1f}+{c:.1f}" for a, b, c in ts))
ctx.close()
