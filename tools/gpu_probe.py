"""Scratch GPU probe: parity + timing of the first CUDA path (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import oracle as orc
import systems as S, helpers as H
from critic2_b200 import capi

ctx = capi.Context(0)
print(ctx.describe(), flush=True)
ctx.profile_enable(True)

def run_case(n, nat, cellp, seed, algos=(capi.BADER_EXACT, capi.BADER_FAST), check=True):
    x2c = S.cell_x2c(*cellp)
    at, z, al = S.random_atoms(nat, seed, x2c)
    at = S.snap_to_grid(at, n)
    f = orc.promolecular(n, x2c, at, z, al, nimg=1)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    if check:
        t = time.time(); idg, nattr, xattr, st = orc.bader_integrate(f, x2c, atoms=at); tref = time.time() - t
    h = ctx.upload(f)
    lap = np.asfortranarray(np.gradient(np.gradient(f, axis=0), axis=0))
    h2 = ctx.upload(lap)
    for algo in algos:
        ctx.profile_reset()
        t = time.time(); b = ctx.bader_assign(h, car2lat, lid, algo=algo); tg = time.time() - t
        mp, na, xa = H.assign_attractors(b.maxima(), n, x2c, at)
        b.set_map(na, mp)
        lab = b.labels(n)
        vol, ps = ctx.integrate(b, [h, h2], S.omega(x2c))
        prof = ctx.profile()
        msg = f"{n} nat={nat} algo={algo} nmax={b.nmax} gpu_wall={tg*1e3:.1f}ms stats={b.stats()[:6]}"
        if check:
            mism = int(np.count_nonzero(lab != idg))
            vref, pref = orc.integrate_bader(idg, [f, lap], nattr, S.omega(x2c))
            msg += f" ref={tref:.1f}s MISMATCH={mism} nattr={na}/{nattr} dvol={np.abs(vol-vref).max():.2e} dpop_rel={np.abs(ps[:,0]-pref[:,0]).max()/np.abs(pref[:,0]).max():.2e}"
        print(msg, flush=True)
        print("   ", {k: round(v[0], 3) for k, v in prof.items()}, flush=True)
        b.free()
    ctx.free(h); ctx.free(h2)

run_case((48, 48, 48), 6, (9, 9, 9, 90, 90, 90), 1)
run_case((64, 68, 72), 8, (10, 10.5, 11, 85, 95, 100), 2)
run_case((96, 96, 96), 32, (16, 16, 16, 90, 90, 90), 3)
run_case((50, 61, 47), 5, (8, 9.5, 7.7, 90, 90, 90), 7)

# NCI parity
n = (40, 44, 48)
x2c = S.cell_x2c(9, 10, 11, 88, 93, 97)
at, z, al = S.random_atoms(5, 11, x2c)
f = orc.promolecular(n, x2c, at, z, al, nimg=1)
h = ctx.upload(f)
crho_o, cgrad_o, lam2 = orc.nci_rdg(f, x2c, want_lam2=True)
crho, cgrad = ctx.nci_rdg(h, x2c, n)
rel = np.abs(cgrad - cgrad_o) / np.maximum(np.abs(cgrad_o), 1e-300)
sgn = np.count_nonzero(np.sign(crho) != np.sign(crho_o))
print("NCI aligned: max rel dRDG %.3e  max rel |crho| %.3e sign mismatches %d of %d" % (rel.max(), (np.abs(np.abs(crho) - np.abs(crho_o)) / np.abs(crho_o)).max(), sgn, crho.size), flush=True)
# off-node lattice
nstep = (23, 19, 17)
x0 = x2c @ np.array([0.013, 0.021, 0.034])
xmat = x2c / np.array(nstep, dtype=float)[None, :] * 0.93
crho_o, cgrad_o, lam2 = orc.nci_rdg(f, x2c, nstep=nstep, x0=x0, xmat=xmat, want_lam2=True)
crho, cgrad = ctx.nci_rdg(h, x2c, n, nstep=nstep, x0=x0, xmat=xmat)
rel = np.abs(cgrad - cgrad_o) / np.maximum(np.abs(cgrad_o), 1e-300)
sgn = np.count_nonzero(np.sign(crho) != np.sign(crho_o))
print("NCI general: max rel dRDG %.3e  max rel |crho| %.3e sign mismatches %d of %d" % (rel.max(), (np.abs(np.abs(crho) - np.abs(crho_o)) / np.abs(crho_o)).max(), sgn, crho.size), flush=True)
ctx.free(h)

# timing at 256^3 (GPU-generated density), both algos, no CPU check
for nn_, nat_side in ((256, 4), (512, 8)):
    n = (nn_,) * 3
    x2c = S.cell_x2c(5.0 * nat_side, 5.0 * nat_side, 5.0 * nat_side)
    at, z, al = S.jittered_lattice(nat_side, 5)
    at = S.snap_to_grid(at, n)
    h = ctx.alloc(n)
    t = time.time(); ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0); print("gen %d^3: %.2fs" % (nn_, time.time() - t), flush=True)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    labs = {}
    for algo in (capi.BADER_EXACT, capi.BADER_FAST):
        for rep in range(2):
            ctx.profile_reset()
            t = time.time(); b = ctx.bader_assign(h, car2lat, lid, algo=algo); tg = time.time() - t
            prof = ctx.profile()
            if rep == 0:
                mp, na, xa = H.assign_attractors(b.maxima(), n, x2c, at)
                b.set_map(na, mp)
                labs[algo] = b.labels(n)
            print(f"{nn_}^3 algo={algo} rep={rep} nmax={b.nmax} wall={tg*1e3:.1f}ms stats={b.stats()[:6]}", flush=True)
            print("   ", {k: round(v[0], 3) for k, v in prof.items()}, flush=True)
            if rep == 1:
                ctx.profile_reset()
                t = time.time(); vol, ps = ctx.integrate(b, [h, h], S.omega(x2c)) if b.nattr else (None, None); ti = time.time() - t
                print("    integrate wall %.1fms" % (ti * 1e3), {k: round(v[0], 3) for k, v in ctx.profile().items()}, flush=True)
            b.free()
    print(f"{nn_}^3 FAST vs EXACT label mismatches:", int(np.count_nonzero(labs[0] != labs[1])), flush=True)
    ctx.free(h)
ctx.close()
