"""Multi-GPU (z-slab) BADER through the C ABI: one process per GPU, NCCL for the slab replication, the
label halo exchange, the candidate all-gather and the basin all-reduce.  Needs >= 2 GPUs (skipped on a
single-GPU box); the host-side logic is covered on CPU by tests/test_multi_gloo.py."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import cases
import systems as S
from critic2_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return len([l for l in out.splitlines() if l.startswith("GPU ")])
    except Exception:
        return 0


@pytest.mark.parametrize("nranks", [2, 4])
@pytest.mark.parametrize("name", ["triclinic", "odd_dims"])
def test_slab_sharded_bader_equals_oracle(nranks, name):
    if _ngpus() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    buf = ctypes.create_string_buffer(128)
    assert capi.load().c2g_nccl_unique_id(buf) == 0
    c = cases.make_case(name)
    idg, nattr, _, _ = orc.bader_integrate(c["f"], c["x2c"], atoms=c["atoms"])
    f2 = cases.second_field(c["f"])
    vref, pref = orc.integrate_bader(idg, [c["f"], f2], nattr, S.omega(c["x2c"]))
    with tempfile.TemporaryDirectory() as d:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "multi_worker.py"), str(r), str(nranks),
                                   buf.raw.hex(), name, d]) for r in range(nranks)]
        for p in procs:
            assert p.wait(timeout=600) == 0
        parts = [np.load(os.path.join(d, f"rank{r}.npz")) for r in range(nranks)]
    for algo in (capi.BADER_EXACT, capi.BADER_FAST):
        lab = np.concatenate([p[f"lab{algo}"] for p in parts], axis=2)
        assert np.array_equal(lab, idg)
        for p in parts:  # all-reduced results are identical on every rank
            assert np.array_equal(p[f"vol{algo}"], vref)
            assert np.abs(p[f"ps{algo}"][:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
            assert int(p[f"cnt{algo}"].sum()) == idg.size
    # multipoles: slab partials all-reduced, identical on every rank up to the order of the atomic adds
    if nattr == len(c["atoms"]):
        ortho = bool(np.all(c["x2c"] - np.diag(np.diag(c["x2c"])) == 0.0))
        cell = orc.Cell(c["x2c"]) if ortho else orc.Cell(c["x2c"], ws=c["x2c"] @ S.wscell(c["x2c"])[0].T.astype(float))
        xattr = np.asarray(c["atoms"], dtype=float).T
        mref = orc.multipoles_bader(idg, xattr, 3, c["f"], cell, S.omega(c["x2c"]))
        rmax = 0.5 * np.linalg.norm(c["x2c"], axis=0).sum()
        scale = np.abs(mref[0]).max() * rmax ** np.repeat(np.arange(4), 2 * np.arange(4) + 1)
        for p in parts:
            assert np.all(np.abs(p["mpole"] - mref) <= 1e-10 * scale[:, None])
        nattn_o, idg1_o, iatt_o, ilvec_o = orc.bader_remap(idg, xattr, cell)
        assert np.array_equal(np.concatenate([p["rm_idg1"] for p in parts], axis=2), idg1_o)
        for p in parts:
            assert np.array_equal(p["rm_iatt"], iatt_o) and np.array_equal(p["rm_ilvec"], ilvec_o)
    # NCIPLOT sharded along i: the ranks' pieces concatenate to the single-process result of the oracle
    crho_o, cgrad_o = orc.nci_rdg(c["f"], c["x2c"])
    assert [int(p["nci_ilo"]) for p in parts][1:] == [int(p["nci_ihi"]) for p in parts][:-1]
    cgrad = np.concatenate([p["cgrad"] for p in parts], axis=2)
    crho = np.concatenate([p["crho"] for p in parts], axis=2)
    assert cgrad.shape == cgrad_o.shape
    rel = np.abs(cgrad - cgrad_o) / np.abs(cgrad_o).max()
    assert (rel <= 1e-12).mean() >= 0.999 and rel.max() <= 1e-9
    assert np.abs(np.abs(crho) - np.abs(crho_o)).max() <= 1e-12 * np.abs(crho_o).max()
