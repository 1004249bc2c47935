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


@pytest.mark.parametrize("ngpus", [2, 4])
def test_single_process_multi_device_context(ngpus):
    """c2g_init_devices (SURVEY.md 8b `c2g_init(ngpus)`): ONE process -- what critic2 is -- drives `ngpus` devices; every
    call takes and returns whole arrays exactly like the one-GPU context, the z-slabs and NCCL live inside."""
    if _ngpus() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    import helpers as H
    ctx = capi.Context(ngpus=ngpus)
    assert f"x{ngpus}" in ctx.describe()
    for name in ("triclinic", "odd_dims"):
        c = cases.make_case(name)
        n, x2c = c["n"], c["x2c"]
        idg, nattr, _, _ = orc.bader_integrate(c["f"], x2c, atoms=c["atoms"])
        f2 = cases.second_field(c["f"])
        vref, pref = orc.integrate_bader(idg, [c["f"], f2], nattr, S.omega(x2c))
        _, car2lat, lid = orc.bader_metrics(x2c, n)
        h, h2 = ctx.upload(c["f"]), ctx.upload(f2)          # ONE host array each: scattered as slabs, replicated over NVLink
        assert np.array_equal(ctx.download(h, n), c["f"])
        for algo in (capi.BADER_FAST, capi.BADER_EXACT):
            b = ctx.bader_assign(h, car2lat, lid, algo=algo)
            mp, na, _ = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
            b.set_map(na, mp)
            assert na == nattr and np.array_equal(b.labels(n), idg)          # the whole idg(n1,n2,n3)
            vol, ps = ctx.integrate(b, [h, h2], S.omega(x2c))
            assert np.array_equal(vol, vref)
            assert np.abs(ps[:, 0] - pref[:, 0]).max() <= 1e-10 * np.abs(pref[:, 0]).max()
            assert int(b.counts().sum()) == idg.size
            if algo == capi.BADER_FAST and nattr == len(c["atoms"]):
                ortho = bool(np.all(x2c - np.diag(np.diag(x2c)) == 0.0))
                kw = {} if ortho else dict(ws=np.asfortranarray(x2c @ S.wscell(x2c)[0].T.astype(float)))
                cell = orc.Cell(x2c) if ortho else orc.Cell(x2c, ws=x2c @ S.wscell(x2c)[0].T.astype(float))
                xattr = np.asarray(c["atoms"], dtype=float).T
                mpole = ctx.integrate_multipoles(b, h, 3, xattr, x2c, S.omega(x2c), **kw)
                mref = orc.multipoles_bader(idg, xattr, 3, c["f"], cell, S.omega(x2c))
                rmax = 0.5 * np.linalg.norm(x2c, axis=0).sum()
                scale = np.abs(mref[0]).max() * rmax ** np.repeat(np.arange(4), 2 * np.arange(4) + 1)
                assert np.all(np.abs(mpole - mref) <= 1e-10 * scale[:, None])
                nattn, idg1, iatt, ilvec = ctx.basins_remap(b, xattr, x2c, shape=n, **kw)
                nattn_o, idg1_o, iatt_o, ilvec_o = orc.bader_remap(idg, xattr, cell)
                assert nattn == nattn_o and np.array_equal(idg1, idg1_o) and np.array_equal(iatt, iatt_o) and np.array_equal(ilvec, ilvec_o)
            b.free()
        # NCIPLOT: rows sharded inside, the caller gets the whole (k,j,i) arrays
        crho, cgrad = ctx.nci_rdg(h, x2c, n)
        crho_o, cgrad_o = orc.nci_rdg(c["f"], x2c)
        rel = np.abs(cgrad - cgrad_o) / np.abs(cgrad_o).max()
        assert cgrad.shape == cgrad_o.shape and (rel <= 1e-12).mean() >= 0.999 and rel.max() <= 1e-9
        # replicas-only paths: FFT field on every device, YT on the first one
        lap = ctx.fft_derivative(h, x2c, "lap")
        lap_o = orc.fft_derivative(c["f"], x2c, "lap")
        assert np.abs(ctx.download(lap, n) - lap_o).max() <= 1e-12 * np.abs(lap_o).max()
        ctx.free(lap)
        vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
        d = orc.yt_integrate(c["f"], x2c, vec, area, atoms=c["atoms"])
        y = ctx.yt_build(h, vec, area)
        mp, na, _ = H.assign_attractors(y.maxima(), n, x2c, c["atoms"])
        y.set_map(na, mp)
        assert np.array_equal(y.labels(n), d.spatial_basin(n))
        vol, ps = ctx.integrate(y, [h], S.omega(x2c))
        vr, pr = orc.integrate_yt(d, [c["f"]], S.omega(x2c))
        assert np.abs(ps[:, 0] - pr[:, 0]).max() <= 1e-10 * np.abs(pr[:, 0]).max()
        y.free()
        with pytest.raises(capi.C2GError):
            ctx.parse_text(b"1.0 2.0\n", (1, 1, 2), 0, 1.0)   # the text codec is single-device
        ctx.free(h); ctx.free(h2)
    ctx.close()
