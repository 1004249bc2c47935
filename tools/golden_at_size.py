#!/usr/bin/env python
"""Runs the CPU oracle (faithful bader_integrate) at the BASELINE.json sizes next to the device path and records
the fixtures that tests/test_gpu_at_size.py asserts against (tests/golden/bader_at_size.json).

For every case of tests/sized_cases.py named on the command line: the density is generated in HBM, downloaded, the
oracle runs on it on the host (serial, like the reference), and the device labels (C2G_BADER_FAST and, with --exact,
the exact-walk referee) are compared with the oracle's on EVERY point.  Written per case: SHA-256 of the density
and of the oracle's labels (Fortran order, int32), attractor count, points per basin, mismatch counts, oracle time.

The 1024^3 oracle run needs ~25 GB of host memory and 15-25 min on one core; it is run once per change of the
generator and its fixture is committed.

usage: python tools/golden_at_size.py [--exact] [--out=gpurun_out/golden_at_size.json] case [case ...]
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import sized_cases as Z
from critic2_b200 import capi
from oracle import oracle as orc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a.ravel(order="F")).tobytes()).hexdigest()


def main():
    exact = "--exact" in sys.argv
    out = "gpurun_out/golden_at_size.json"
    names = []
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            out = a[6:]
        elif not a.startswith("--"):
            names.append(a)
    res = {}
    if os.path.exists(out):
        res = json.load(open(out))
    ctx = capi.Context(0)
    for name in names:
        c = Z.CASES[name]()
        n, x2c, at = c["n"], c["x2c"], c["atoms"]
        h = ctx.alloc(n)
        ctx.promolecular(h, x2c, at, c["z"], c["alpha"], nimg=c["nimg"], rc=c["rc"])
        f = ctx.download(h, n)
        rec = {"n": list(map(int, n)), "natoms": int(len(at)), "rho_sha256": sha(f)}
        _, car2lat, lid = orc.bader_metrics(x2c, n)
        dev = {}
        for algo, tag in ((capi.BADER_FAST, "fast"),) + (((capi.BADER_EXACT, "exact"),) if exact else ()):
            ctx.synchronize(); ctx.timer_start()
            b = ctx.bader_assign(h, car2lat, lid, algo=algo)
            ms = ctx.timer_stop()
            mp, dist = Z.atom_map(b.maxima(), n, x2c, at)
            b.set_map(len(at), mp)
            dev[tag] = b.labels(n)
            rec[tag + "_nmax"] = int(b.nmax); rec[tag + "_ms"] = round(ms, 3)
            rec[tag + "_max_dist_to_atom"] = float(dist.max())
            rec[tag + "_stats"] = [int(v) for v in b.stats()]
            b.free()
        ctx.free(h)
        t0 = time.perf_counter()
        idg, nattr, _, stats = orc.bader_integrate(f, x2c, atoms=at)
        rec["oracle_seconds"] = round(time.perf_counter() - t0, 2)
        rec["oracle_points_per_s"] = float(np.prod(n)) / (time.perf_counter() - t0)
        rec["nattr"] = int(nattr)
        rec["labels_sha256"] = sha(idg)
        rec["counts"] = np.bincount(idg.ravel(), minlength=nattr + 1).astype(int).tolist()
        rec["oracle_stats"] = [int(v) for v in stats]
        for tag, lab in dev.items():
            rec[tag + "_mismatches_vs_oracle"] = int(np.count_nonzero(lab != idg))
            rec[tag + "_labels_sha256"] = sha(lab)
        if exact:
            rec["fast_vs_exact_mismatches"] = int(np.count_nonzero(dev["fast"] != dev["exact"]))
        res[name] = rec
        print(name, json.dumps({k: v for k, v in rec.items() if k != "counts"}), flush=True)
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        json.dump(res, open(out, "w"), indent=1)
        del f, idg, dev
    ctx.close()


if __name__ == "__main__":
    main()
