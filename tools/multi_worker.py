"""Worker of tests/test_gpu_multi.py: one rank of a multi-GPU BADER run through the C ABI.
usage: multi_worker.py <rank> <nranks> <uid_hex> <case> <outdir>"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, helpers as H, systems as S
from critic2_b200 import capi
from oracle import oracle as orc

rank, nranks, uid_hex, name, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
uid = ctypes.create_string_buffer(bytes.fromhex(uid_hex), 128)
ctx = capi.Context(rank, rank=rank, nranks=nranks, nccl_uid=uid)
c = cases.make_case(name)
n, x2c = c["n"], c["x2c"]
zlo, zhi = ctx.slab_range(n[2])
_, car2lat, lid = orc.bader_metrics(x2c, n)
f2 = cases.second_field(c["f"])
h = ctx.upload_slab(c["f"][:, :, zlo:zhi], n)      # every rank uploads only its slab; NCCL replicates
h2 = ctx.upload_slab(f2[:, :, zlo:zhi], n)
full = ctx.download(h, n)
assert np.array_equal(full, c["f"]), "slab all-gather mismatch"
out = {}
for algo in (capi.BADER_EXACT, capi.BADER_FAST):
    b = ctx.bader_assign(h, car2lat, lid, algo=algo, order=capi.ORDER_SCAN)
    mp, na, _ = H.assign_attractors(b.maxima(), n, x2c, c["atoms"])
    b.set_map(na, mp)
    lab = b.labels((n[0], n[1], max(zhi - zlo, 0)))
    vol, ps = ctx.integrate(b, [h, h2], S.omega(x2c))
    out[f"lab{algo}"] = lab; out[f"vol{algo}"] = vol; out[f"ps{algo}"] = ps; out[f"cnt{algo}"] = b.counts()
    if algo == capi.BADER_FAST:  # multipoles: per-slab partial moments, all-reduced over NCCL
        ortho = bool(np.all(x2c - np.diag(np.diag(x2c)) == 0.0))
        kw = {} if ortho else dict(ws=np.asfortranarray(x2c @ S.wscell(x2c)[0].T.astype(float)))
        xattr = np.asarray(c["atoms"], dtype=float).T
        if xattr.shape[1] == na:
            out["mpole"] = ctx.integrate_multipoles(b, h, 3, xattr, x2c, S.omega(x2c), **kw)
            # DELOC attractor images: per-slab first-appearance tables min-reduced over NCCL, idg1 per slab
            nattn, idg1, iatt, ilvec = ctx.basins_remap(b, xattr, x2c, shape=(n[0], n[1], max(zhi - zlo, 0)), **kw)
            out["rm_idg1"] = idg1; out["rm_iatt"] = iatt; out["rm_ilvec"] = ilvec
    b.free()
# NCIPLOT: the output lattice is sharded along i, no collective
ilo, ihi = ctx.nci_range(n[0])
crho, cgrad = ctx.nci_rdg(h, x2c, n)
out["nci_ilo"] = ilo; out["nci_ihi"] = ihi; out["crho"] = crho; out["cgrad"] = cgrad
np.savez(os.path.join(outdir, f"rank{rank}.npz"), zlo=zlo, zhi=zhi, **out)
ctx.close()
print("rank", rank, "ok", zlo, zhi)
