"""The C++ mirror of the reference's driver interface (critic2_b200/csrc/host: bader_integrate, intgrid_fields on
basindat / system objects, ferror-style failures) driven end to end on the GPU and compared with the oracle.
The mirror only marshals into the C ABI, so this also covers C2G_ORDER_SCAN + the host-side attractor
identification (identify_atom / are_lclose) a Fortran caller would do."""
import json
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

import cases
import systems as S
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "critic2_b200", "csrc", "host", "host_selftest")


def fnv1a(labels):
    h = 1469598103934665603
    for v in labels.ravel(order="F").tolist():
        h ^= v & 0xFFFFFFFF
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return format(h, "x")


@pytest.mark.parametrize("name", ["cubic48", "ortho_flat"])
def test_host_mirror_bader_and_integrable(name):
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    c = cases.make_case(name)
    n, x2c = c["n"], c["x2c"]
    idg, nattr, _, _ = orc.bader_integrate(c["f"], x2c, atoms=c["atoms"])
    vref, pref = orc.integrate_bader(idg, [c["f"]], nattr, S.omega(x2c))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "in.bin")
        with open(path, "wb") as fp:
            fp.write(struct.pack("4i", n[0], n[1], n[2], len(c["atoms"])))
            fp.write(np.asfortranarray(x2c).ravel(order="F").astype(np.float64).tobytes())
            fp.write(np.ascontiguousarray(c["atoms"], dtype=np.float64).tobytes())
            fp.write(np.asfortranarray(c["f"]).ravel(order="F").tobytes())
        out = subprocess.run([EXE, path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "ERROR : bader_integrate: inconsistent field size" in out.stderr   # the deliberate bad call
    r = json.loads(out.stdout)
    assert r["nattr"] == nattr
    assert r["labels_fnv"] == fnv1a(idg)
    # volumes are exact point counts times omega/ntot; the host mirror evaluates omega with det3's expression
    # (tools_math det3), numpy with an LU factorisation: the two differ by a few ulp (north_star allows 1e-10)
    assert np.abs(np.array(r["vol"]) - vref).max() <= 1e-14 * np.abs(vref).max()
    assert np.abs(np.array(r["pop"]) - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
    # INTEGRABLE ... MULTIPOLES 2 through intgrid_multipoles (attractors = atoms here)
    xattr = np.asfortranarray(np.asarray(c["atoms"], dtype=float).T)
    mref = orc.multipoles_bader(idg, xattr, 2, c["f"], orc.Cell(x2c), S.omega(x2c))
    got = np.array(r["mpole_lmax2"]).reshape(mref.shape, order="F")
    rmax = 0.5 * np.linalg.norm(x2c, axis=0).sum()
    assert np.abs(got - mref).max() <= 1e-10 * np.abs(pref[:, 0]).max() * rmax ** 2
