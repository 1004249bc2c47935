#!/usr/bin/env python
"""Pins the CPU oracle (oracle/oracle.cpp) against REAL critic2 -- for whoever has a Fortran compiler.

critic2 cannot be built in the development image (no gfortran), so the oracle's BADER / YT restatement is pinned by
code review only.  This tool closes that gap in one command on any machine with a critic2 binary:

  python tools/pin_cases.py write  <dir>            # cube files + .cri inputs of the seeded parity cases
  (cd <dir> && for f in *.cri; do critic2 $f ${f%.cri}.cro; done)     # = tools/pin_with_critic2.sh
  python tools/pin_cases.py compare <dir>           # critic2's basins against the oracle's, point by point

write:   every case of tests/cases.py SMALL_CASES (+ a 4-molecule ammonia cell) is written as a Gaussian cube file with
         17 significant digits (so that critic2's list-directed READ recovers the very doubles the oracle sees) and two
         inputs: <case>_bader.cri (`bader wcube`) and <case>_yt.cri (`yt wcube`).
compare: BADER -- <case>_bader_wcube_all.cube holds bas%idg itself (int_cubew, integration@proc.f90:4465-4469): compared
         with the oracle's idg as a PARTITION (basin numbers are canonicalised by first appearance, so critic2's
         attractor reordering in int_reorder_gridout does not matter).  YT -- <case>_yt_wcube_NN.cube hold the weight of
         every basin with the 5 digits of writegrid_cube's default format: the interior points (w = 1) must coincide
         with the oracle's ibasin and the fractional weights agree to 2e-5.
Exit status 0 = every case pinned.
"""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import cases
import systems as S
from oracle import oracle as orc

NAMES = [c[0] for c in cases.SMALL_CASES if c[0] != "tiny"]


def write_cube(path, f, x2c, atoms, z):
    n = f.shape
    with open(path, "w") as o:
        o.write("critic2_b200 parity case\nvalues: %%24.16E, k fastest\n")
        o.write(f"{len(atoms):5d} {0.0:12.6f} {0.0:12.6f} {0.0:12.6f}\n")
        for i in range(3):
            v = x2c[:, i] / n[i]
            o.write(f"{n[i]:5d} {v[0]:22.16f} {v[1]:22.16f} {v[2]:22.16f}\n")
        for a, zz in zip(atoms, z):
            c = x2c @ a
            zi = int(min(36, max(1, round(zz))))
            o.write(f"{zi:5d} {float(zi):12.6f} {c[0]:22.16f} {c[1]:22.16f} {c[2]:22.16f}\n")
        for i in range(n[0]):
            for j in range(n[1]):
                row = f[i, j, :]
                for k0 in range(0, n[2], 6):
                    o.write(" ".join(f"{v:24.16E}" for v in row[k0:k0 + 6]) + "\n")


def read_cube(path):
    with open(path) as fh:
        lines = fh.readlines()
    nat = int(lines[2].split()[0])
    n = [int(lines[3 + i].split()[0]) for i in range(3)]
    vals = np.array(" ".join(lines[6 + abs(nat):]).split(), dtype=float)
    return np.asfortranarray(vals.reshape(n, order="C"))


def canonical(lab):
    """Basin numbers by first appearance in array element order (0 stays 0)."""
    flat = lab.ravel(order="F")
    _, first = np.unique(flat, return_index=True)
    order = flat[np.sort(first)]
    lut = {int(v): (0 if v == 0 else k + 1) for k, v in enumerate([u for u in order if u != 0])}
    lut[0] = 0
    return np.vectorize(lut.get)(lab)


def write(dirname):
    os.makedirs(dirname, exist_ok=True)
    for name in NAMES:
        c = cases.make_case(name)
        write_cube(os.path.join(dirname, f"{name}.cube"), c["f"], c["x2c"], c["atoms"], c["z"])
        for key in ("bader", "yt"):
            with open(os.path.join(dirname, f"{name}_{key}.cri"), "w") as o:
                o.write(f"crystal {name}.cube\nload {name}.cube\n{key} wcube\n")
    print(f"wrote {len(NAMES)} cases to {dirname}; run critic2 on every .cri there (tools/pin_with_critic2.sh), then `compare`")


def compare(dirname):
    bad = 0
    for name in NAMES:
        f = read_cube(os.path.join(dirname, f"{name}.cube"))       # exactly what critic2 read
        c = cases.make_case(name)
        assert np.array_equal(f, c["f"]), "the cube text does not round-trip"
        x2c, atoms = c["x2c"], c["atoms"]
        allc = os.path.join(dirname, f"{name}_bader_wcube_all.cube")
        if os.path.exists(allc):
            idg_ref = np.rint(read_cube(allc)).astype(np.int64)
            idg, nattr, _, _ = orc.bader_integrate(f, x2c, atoms=atoms)
            d = int(np.count_nonzero(canonical(idg_ref) != canonical(idg)))
            print(f"{name}: BADER {idg.size} points, {nattr} basins (critic2: {idg_ref.max()}), partition mismatches {d}")
            bad += d != 0
        else:
            print(f"{name}: {allc} missing (critic2 not run?)"); bad += 1
        wfiles = sorted(glob.glob(os.path.join(dirname, f"{name}_yt_wcube_[0-9]*.cube")))
        if wfiles:
            vec, area = S.wscell(x2c / np.array(f.shape, dtype=float)[None, :])
            dd = orc.yt_integrate(f, x2c, vec, area, atoms=atoms)
            wo = [orc.yt_weights(dd, i, f.shape) for i in range(1, dd.nattr + 1)]
            worst, nint = 0.0, 0
            for wf in wfiles:
                w = read_cube(wf)
                k = int(np.argmax([float((w * x).sum()) for x in wo]))      # the oracle basin this file belongs to
                worst = max(worst, float(np.abs(w - wo[k]).max()))
                nint += int(np.count_nonzero((w == 1.0) != (wo[k] == 1.0)))
            print(f"{name}: YT {len(wfiles)} basins (oracle {dd.nattr}), interior-point mismatches {nint}, max weight difference {worst:.1e}")
            bad += (nint != 0) or (worst > 2e-5) or (len(wfiles) != dd.nattr)
        else:
            print(f"{name}: no {name}_yt_wcube_NN.cube (critic2 not run?)"); bad += 1
    print("PINNED" if bad == 0 else f"{bad} check(s) failed")
    return 1 if bad else 0


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("write", "compare"):
        sys.exit(__doc__)
    sys.exit(write(sys.argv[2]) or 0 if sys.argv[1] == "write" else compare(sys.argv[2]))
