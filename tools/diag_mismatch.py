#!/usr/bin/env python
"""FAST vs EXACT label differences of a sized case (tests/sized_cases.py), with the environment switches that isolate
their cause.  For every switch set: number of mismatching points; for the default set also where they are, what
their 26 neighbours carry and how far the nearest point of another basin is.
usage: python tools/diag_mismatch.py case [case ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sized_cases as Z
from critic2_b200 import capi
from oracle import oracle as orc

ctx = capi.Context(0)
VARIANTS = [("default", {}), ("no early stop", {"C2G_NO_EARLY_STOP": "1"}), ("L0=4", {"C2G_BADER_L0": "4"}),
            ("certificates only at stride 2", {"C2G_SAFE_MAXS": "2"}), ("per-lane refill walkers (k_walk2)", {"C2G_WALK2": "1"}), ("5x5x5 cube certificates", {"C2G_CERT": "2"}), ("bare octets (3x3x3 at stride 2, round 1)", {"C2G_CERT": "3"}),
            ("block-cooperative walkers (k_walk3) for every launch", {"C2G_WALK3": "15"})]
for name in sys.argv[1:]:
    c = Z.CASES[name]()
    n, x2c, at = c["n"], c["x2c"], c["atoms"]
    h = ctx.alloc(n)
    ctx.promolecular(h, x2c, at, c["z"], c["alpha"], nimg=c["nimg"], rc=c["rc"])
    _, car2lat, lid = orc.bader_metrics(x2c, n)

    def labels(algo):
        b = ctx.bader_assign(h, car2lat, lid, algo=algo)
        mp, _ = Z.atom_map(b.maxima(), n, x2c, at)
        b.set_map(len(at), mp)
        lab = b.labels(n); st = b.stats(); b.free()
        return lab, st
    ex, _ = labels(capi.BADER_EXACT)
    for tag, env in VARIANTS:
        os.environ.update(env)
        try:
            fa, st = labels(capi.BADER_FAST)
        finally:
            for k in env:
                del os.environ[k]
        mm = np.argwhere(fa != ex)
        print(f"{name} [{tag}]: {len(mm)} of {fa.size} FAST labels differ from EXACT; stats {[int(v) for v in st]}", flush=True)
        if tag != "default":
            continue
        for p in mm[:12]:
            p = tuple(int(v) for v in p)
            nb = [fa[(p[0] + a) % n[0], (p[1] + b) % n[1], (p[2] + cc) % n[2]] for a in (-1, 0, 1) for b in (-1, 0, 1) for cc in (-1, 0, 1)]
            nbx = [ex[(p[0] + a) % n[0], (p[1] + b) % n[1], (p[2] + cc) % n[2]] for a in (-1, 0, 1) for b in (-1, 0, 1) for cc in (-1, 0, 1)]
            # Chebyshev distance to the nearest point whose EXACT label differs from the FAST label of p
            dist = None
            for r in range(1, 12):
                sl = np.ix_(*[[(p[d] + k) % n[d] for k in range(-r, r + 1)] for d in range(3)])
                if np.any(ex[sl] != fa[p]):
                    # ignore p itself
                    blk = ex[sl].copy(); blk[r, r, r] = fa[p]
                    if np.any(blk != fa[p]):
                        dist = r; break
            print(f"   point {p}: FAST {fa[p]} EXACT {ex[p]}; FAST labels of the 27-neighbourhood: {sorted(set(nb))}; EXACT: {sorted(set(nbx))}; "
                  f"nearest other-basin point (EXACT, Chebyshev): {dist}; parity of coords {tuple(v % 2 for v in p)} mod4 {tuple(v % 4 for v in p)} mod8 {tuple(v % 8 for v in p)}", flush=True)
    ctx.free(h)
ctx.close()
