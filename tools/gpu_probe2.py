"""Scratch GPU probe 2: YT parity + Bader timing after hot-loop rewrite."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import oracle as orc
import systems as S, helpers as H
from critic2_b200 import capi

ctx = capi.Context(0)
print(ctx.describe(), flush=True)
ctx.profile_enable(True)

def bader_case(n, nat, cellp, seed):
    x2c = S.cell_x2c(*cellp)
    at, z, al = S.random_atoms(nat, seed, x2c)
    at = S.snap_to_grid(at, n)
    f = orc.promolecular(n, x2c, at, z, al, nimg=1)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    idg, nattr, xattr, st = orc.bader_integrate(f, x2c, atoms=at)
    h = ctx.upload(f)
    for algo in (1, 0):
        b = ctx.bader_assign(h, car2lat, lid, algo=algo)
        mp, na, xa = H.assign_attractors(b.maxima(), n, x2c, at)
        b.set_map(na, mp)
        lab = b.labels(n)
        print(f"BADER {n} algo={algo} nmax={b.nmax} MISMATCH={int(np.count_nonzero(lab != idg))} counts_ok={int(b.counts().sum()) == f.size}", flush=True)
        b.free()
    ctx.free(h)

bader_case((48, 48, 48), 6, (9, 9, 9, 90, 90, 90), 1)
bader_case((64, 68, 72), 8, (10, 10.5, 11, 85, 95, 100), 2)

def yt_case(n, nat, cellp, seed):
    x2c = S.cell_x2c(*cellp)
    at, z, al = S.random_atoms(nat, seed, x2c)
    at = S.snap_to_grid(at, n)
    f = orc.promolecular(n, x2c, at, z, al, nimg=1)
    x2cg = x2c / np.array(n, dtype=float)[None, :]
    vec, area = S.wscell(x2cg)
    t = time.time(); d = orc.yt_integrate(f, x2c, vec, area, atoms=at); tref = time.time() - t
    lap = np.asfortranarray(np.gradient(np.gradient(f, axis=1), axis=1))
    t = time.time(); vref, pref = orc.integrate_yt(d, [f, lap], S.omega(x2c)); tint = time.time() - t
    sb = d.spatial_basin(n)
    h = ctx.upload(f); h2 = ctx.upload(lap)
    ctx.profile_reset()
    t = time.time(); b = ctx.yt_build(h, vec, area); tg = time.time() - t
    mp, na, xa = H.assign_attractors(b.maxima(), n, x2c, at)
    b.set_map(na, mp)
    lab = b.labels(n)
    t = time.time(); vol, ps = ctx.integrate(b, [h, h2], S.omega(x2c)); ti = time.time() - t
    w1 = b.yt_weights(1, n)
    w1o = orc.yt_weights(d, 1, n)
    print(f"YT {n} nvec={len(area)} nmax={b.nmax} nattr={na}/{d.nattr} label MISMATCH={int(np.count_nonzero(lab != sb))} ias={b.stats()[0]} ({b.stats()[0]/f.size:.3f}) levels bfs={b.stats()[1]} kahn={b.stats()[2]} "
          f"dvol_rel={np.abs(vol-vref).max()/np.abs(vref).max():.2e} dpop_rel={np.abs(ps[:,0]-pref[:,0]).max()/np.abs(pref[:,0]).max():.2e} dlap_abs={np.abs(ps[:,1]-pref[:,1]).max():.2e} (scale {np.abs(lap).sum()*S.omega(x2c)/f.size:.2e}) dw={np.abs(w1-w1o).max():.2e} "
          f"cpu build {tref:.1f}s int {tint:.1f}s gpu build {tg*1e3:.1f}ms int {ti*1e3:.1f}ms", flush=True)
    print("   ", {k: round(v[0], 3) for k, v in ctx.profile().items()}, flush=True)
    b.free(); ctx.free(h); ctx.free(h2)

yt_case((32, 32, 32), 4, (7, 7, 7, 90, 90, 90), 21)
yt_case((48, 52, 44), 6, (9, 9.5, 8.5, 80, 95, 105), 22)
yt_case((64, 64, 64), 8, (10, 10, 10, 90, 90, 90), 23)

for nn_, nat_side in ((256, 4), (512, 8)):
    n = (nn_,) * 3
    x2c = S.cell_x2c(5.0 * nat_side, 5.0 * nat_side, 5.0 * nat_side)
    at, z, al = S.jittered_lattice(nat_side, 5)
    at = S.snap_to_grid(at, n)
    h = ctx.alloc(n)
    ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    for algo in (1, 0):
        for rep in range(2):
            ctx.profile_reset()
            t = time.time(); b = ctx.bader_assign(h, car2lat, lid, algo=algo); tg = time.time() - t
            b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
            vol, ps = ctx.integrate(b, [h, h], S.omega(x2c))
            prof = ctx.profile()
            print(f"{nn_}^3 algo={algo} rep={rep} nmax={b.nmax} wall={tg*1e3:.1f}ms stats={b.stats()[:6]} kernels={sum(v[0] for v in prof.values()):.2f}ms", flush=True)
            print("   ", {k: round(v[0], 3) for k, v in prof.items()}, flush=True)
            b.free()
    if nn_ == 256:
        x2cg = x2c / np.array(n, dtype=float)[None, :]
        vec, area = S.wscell(x2cg)
        for rep in range(2):
            ctx.profile_reset()
            t = time.time(); b = ctx.yt_build(h, vec, area); tg = time.time() - t
            b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
            t = time.time(); vol, ps = ctx.integrate(b, [h, h], S.omega(x2c)); ti = time.time() - t
            print(f"YT {nn_}^3 rep={rep} nmax={b.nmax} build {tg*1e3:.1f}ms int {ti*1e3:.1f}ms stats={b.stats()[:5]} sumvol={vol.sum():.6f} omega={S.omega(x2c):.6f}", flush=True)
            print("   ", {k: round(v[0], 3) for k, v in ctx.profile().items()}, flush=True)
            b.free()
    ctx.free(h)
ctx.close()
