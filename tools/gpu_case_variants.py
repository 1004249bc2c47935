"""Run one small parity case against the oracle under several env knobs of c2g_bader_assign."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, helpers as H
from critic2_b200 import capi
from oracle import oracle as orc
ctx = capi.Context(0)
for name in sys.argv[1:]:
    c = cases.make_case(name)
    idg, nattr, _, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    _, car2lat, lid = orc.bader_metrics(c["x2c"], c["n"])
    h = ctx.upload(c["f"])
    for env in ({}, {"C2G_SAFE_MAXS": "32"}, {"C2G_NO_EARLY_STOP": "1"}, {"C2G_BADER_L0": "4"}, {"C2G_BADER_L0": "8"},
                {"C2G_BADER_L0": "4", "C2G_SAFE_MAXS": "2"}):
        for k in ("C2G_SAFE_MAXS", "C2G_NO_EARLY_STOP", "C2G_BADER_L0"):
            os.environ.pop(k, None)
        os.environ.update(env)
        b = ctx.bader_assign(h, car2lat, lid, algo=capi.BADER_FAST)
        mp, na, _ = H.assign_attractors(b.maxima(), c["n"], c["x2c"], c["atoms"])
        b.set_map(na, mp)
        lab = b.labels(c["n"])
        print(name, env, "mismatches vs oracle:", int(np.count_nonzero(lab != idg)), "stats", b.stats()[:7], flush=True)
        b.free()
    ctx.free(h)
ctx.close()
