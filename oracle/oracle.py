"""ctypes front end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/oracle.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Parity status: "parity unpinned" for BADER / YT / grid-field NCI
(no Fortran compiler and no reference input grids in this image); pinned by critic2's
own outputs for the cube writer (fortran_e / format_text_grid), the promolecular
density and the RDG formula (tests/golden/cube_golden.json, see oracle.cpp's header).

Arrays are Fortran-ordered numpy arrays f[n1,n2,n3] (index 1 fastest in memory).
3x3 matrices are numpy arrays M[i,j] passed in Fortran (column-major) order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.asfortranarray(a, dtype=np.float64)


def _i32(a):
    return np.asfortranarray(a, dtype=np.int32)


def _m33(m):
    """3x3 matrix -> 9 doubles, column-major."""
    return np.asfortranarray(np.asarray(m, dtype=np.float64)).ravel(order="F").copy()


def num_threads() -> int:
    return int(lib().orc_num_threads())


def bader_metrics(x2c, n):
    """lat2car, car2lat, lat_i_dist as computed in bader_integrate (bader@proc.f90:124-145).
    car2lat uses numpy.linalg.inv (LAPACK getrf/getri like the reference's matinv)."""
    x2c = np.asarray(x2c, dtype=np.float64)
    lat2car = x2c / np.asarray(n, dtype=np.float64)[None, :]
    car2lat = np.linalg.inv(lat2car)
    lid = np.zeros((3, 3, 3))
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                if i == j == k == 0:
                    continue
                d = lat2car @ np.array([i, j, k], dtype=np.float64)
                lid[i + 1, j + 1, k + 1] = 1.0 / np.sqrt(np.sum(d * d))
    return lat2car, car2lat, lid


def bader_integrate(f, x2c, atoms=None, ratom=1.0, atexist=True, maxattr=None):
    """Faithful bader_integrate.  Returns (idg[n1,n2,n3] int32, nattr, xattr[3,nattr], stats)."""
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    _, car2lat, lid = bader_metrics(x2c, n)
    xat = np.zeros((3, 0)) if atoms is None else _f64(np.asarray(atoms, dtype=np.float64).T)
    nat = xat.shape[1]
    if maxattr is None:
        maxattr = nat + 100000
    idg = np.zeros(f.shape, dtype=np.int32, order="F")
    nattr = C.c_int(0)
    xattr = np.zeros((3, maxattr), order="F")
    stats = np.zeros(4, dtype=np.int64)
    c2l = _m33(car2lat)
    lidf = np.ascontiguousarray(lid, dtype=np.float64).ravel()  # (d1+1)*9+(d2+1)*3+(d3+1)
    x2cf = _m33(x2c)
    rc = lib().orc_bader_integrate(
        _p(f, C.c_double), _p(n, C.c_int), _p(c2l, C.c_double), _p(lidf, C.c_double), _p(x2cf, C.c_double),
        C.c_int(1 if (atexist and nat > 0) else 0), C.c_int(nat), _p(xat, C.c_double), C.c_double(ratom),
        _p(idg, C.c_int), C.byref(nattr), _p(xattr, C.c_double), C.c_int(maxattr), _p(stats, C.c_long))
    if rc != 0:
        raise RuntimeError(f"orc_bader_integrate failed rc={rc}")
    return idg, nattr.value, xattr[:, : nattr.value].copy(), stats


def bader_canonical(f, x2c):
    """Own-trajectory terminal maximum (0-based linear id) of every point + stats."""
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    _, car2lat, lid = bader_metrics(x2c, n)
    term = np.zeros(f.shape, dtype=np.int32, order="F")
    stats = np.zeros(4, dtype=np.int64)
    c2l = _m33(car2lat)
    lidf = np.ascontiguousarray(lid, dtype=np.float64).ravel()
    lib().orc_bader_canonical(_p(f, C.c_double), _p(n, C.c_int), _p(c2l, C.c_double), _p(lidf, C.c_double),
                              _p(term, C.c_int), _p(stats, C.c_long))
    return term, stats


def integrate_bader(idg, fields, nattr, omega):
    idg = _i32(idg)
    n = np.array(idg.shape, dtype=np.int32)
    fl = [_f64(x) for x in fields]
    arr = (C.POINTER(C.c_double) * max(len(fl), 1))(*[_p(x, C.c_double) for x in fl])
    vol = np.zeros(nattr)
    psum = np.zeros((nattr, len(fl)), order="F")
    lib().orc_integrate_bader(_p(idg, C.c_int), _p(n, C.c_int), C.c_int(nattr), C.c_int(len(fl)), arr,
                              C.c_double(omega), _p(vol, C.c_double), _p(psum, C.c_double))
    return vol, psum


def qcksort(arr):
    """Returns the 1-based permutation iord produced by qcksort_r8 (tools@proc.f90:83-168)."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    iord = np.arange(1, arr.size + 1, dtype=np.int32)
    rc = lib().orc_qcksort_r8(_p(arr, C.c_double), _p(iord, C.c_int), C.c_int(1), C.c_int(arr.size))
    if rc != 0:
        raise RuntimeError("qcksort: Increase nstack")
    return iord


class YtData:
    """ytdata record (yt.f90:36-45), rank-indexed."""

    def __init__(self, nn, nvec):
        self.nn, self.nvec = nn, nvec
        self.nlo = np.zeros(nn, dtype=np.int32)
        self.ibasin = np.zeros(nn, dtype=np.int32)
        self.iio = np.zeros(nn, dtype=np.int32)
        self.inear = np.zeros((nvec, nn), dtype=np.int32, order="F")
        self.fnear = np.zeros((nvec, nn), dtype=np.float64, order="F")
        self.nattr = 0
        self.xattr = None

    def spatial_basin(self, shape):
        """ibasin looked up through iio: basin id of each spatial point (0 = IAS)."""
        return self.ibasin[self.iio - 1].reshape(shape, order="F")


def yt_integrate(f, x2c, vec, area, atoms=None, ratom=1.0, atexist=True, stable=False, maxattr=None):
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    nn = f.size
    vec = _i32(np.asarray(vec).reshape(-1, 3).T)  # (3,nvec) column-major
    nvec = vec.shape[1]
    area = np.ascontiguousarray(area, dtype=np.float64)
    xat = np.zeros((3, 0)) if atoms is None else _f64(np.asarray(atoms, dtype=np.float64).T)
    nat = xat.shape[1]
    if maxattr is None:
        maxattr = nat + 100000
    d = YtData(nn, nvec)
    nattr = C.c_int(0)
    xattr = np.zeros((3, maxattr), order="F")
    x2cf = _m33(x2c)
    rc = lib().orc_yt_integrate(
        _p(f, C.c_double), _p(n, C.c_int), C.c_int(nvec), _p(vec, C.c_int), _p(area, C.c_double),
        _p(x2cf, C.c_double), C.c_int(1 if (atexist and nat > 0) else 0), C.c_int(nat), _p(xat, C.c_double),
        C.c_double(ratom), C.c_int(1 if stable else 0),
        _p(d.nlo, C.c_int), _p(d.ibasin, C.c_int), _p(d.iio, C.c_int), _p(d.inear, C.c_int),
        _p(d.fnear, C.c_double), C.byref(nattr), _p(xattr, C.c_double), C.c_int(maxattr))
    if rc != 0:
        raise RuntimeError(f"orc_yt_integrate failed rc={rc}")
    d.nattr = nattr.value
    d.xattr = xattr[:, : d.nattr].copy()
    return d


def yt_reclassify(d, f, vec, mp, shape):
    """The classification loop of yt_integrate (yt@proc.f90:108-186) replayed in pure Python on the oracle's rank
    permutation d.iio with a GIVEN maximum -> basin map `mp` (1-based ids in the order in which the sweep meets the
    maxima; 0 = a maximum rejected by the DISCARD expression, :152-166, which the C++ oracle does not evaluate).
    Small grids only.  Returns the spatial basin ids (0 = IAS or unassigned)."""
    n1, n2, n3 = shape
    nn = n1 * n2 * n3
    vec = np.asarray(vec, dtype=np.int64).reshape(-1, 3)
    rank = np.asarray(d.iio, dtype=np.int64) - 1          # rank (0-based) of every spatial point
    order = np.empty(nn, dtype=np.int64)
    order[rank] = np.arange(nn)                            # spatial point of every rank
    idx = np.arange(nn)
    x, y, z = idx % n1, (idx // n1) % n2, idx // (n1 * n2)
    nb = np.stack([((x + v[0]) % n1) + n1 * (((y + v[1]) % n2) + n2 * ((z + v[2]) % n3)) for v in vec], axis=1)
    nbrank = rank[nb]
    ib = np.zeros(nn, dtype=np.int32)                      # by spatial point
    nfound = 0
    for r in range(nn - 1, -1, -1):
        i = order[r]
        hi = nb[i][nbrank[i] > r]
        if len(hi) == 0:
            ib[i] = mp[nfound]
            nfound += 1
        else:
            b0 = ib[hi[0]]
            ib[i] = b0 if (b0 != 0 and np.all(ib[hi] == b0)) else 0
    return ib.reshape(shape, order="F")


def yt_isosurface(f, vec, isov, stable=False, maxattr=None):
    """yt_isosurface (yt@proc.f90:233-390) without a DISCARD expression.
    Returns (idg[n1,n2,n3], nraw, nattr, xattr[3,nraw])."""
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    vec = _i32(np.asarray(vec, dtype=np.int32).reshape(-1, 3).T)
    if maxattr is None:
        maxattr = 100000
    idg = np.zeros(f.shape, dtype=np.int32, order="F")
    nraw, nattr = C.c_int(0), C.c_int(0)
    xattr = np.zeros((3, maxattr), order="F")
    rc = lib().orc_yt_isosurface(_p(f, C.c_double), _p(n, C.c_int), C.c_int(vec.shape[1]), _p(vec, C.c_int),
                                 C.c_double(isov), C.c_int(1 if stable else 0), _p(idg, C.c_int), C.byref(nraw),
                                 C.byref(nattr), _p(xattr, C.c_double), C.c_int(maxattr))
    if rc != 0:
        raise RuntimeError(f"orc_yt_isosurface failed rc={rc}")
    return idg, nraw.value, nattr.value, xattr[:, : nraw.value].copy()


def yt_weights(d: YtData, idb: int, shape):
    w = np.zeros(shape, order="F")
    lib().orc_yt_weights(C.c_long(d.nn), C.c_int(d.nvec), _p(d.nlo, C.c_int), _p(d.ibasin, C.c_int),
                         _p(d.iio, C.c_int), _p(d.inear, C.c_int), _p(d.fnear, C.c_double), C.c_int(idb),
                         _p(w, C.c_double))
    return w


def integrate_yt(d: YtData, fields, omega):
    fl = [_f64(x) for x in fields]
    arr = (C.POINTER(C.c_double) * max(len(fl), 1))(*[_p(x, C.c_double) for x in fl])
    vol = np.zeros(d.nattr)
    psum = np.zeros((d.nattr, len(fl)), order="F")
    lib().orc_integrate_yt(C.c_long(d.nn), C.c_int(d.nvec), _p(d.nlo, C.c_int), _p(d.ibasin, C.c_int),
                           _p(d.iio, C.c_int), _p(d.inear, C.c_int), _p(d.fnear, C.c_double),
                           C.c_int(d.nattr), C.c_int(len(fl)), arr, C.c_double(omega),
                           _p(vol, C.c_double), _p(psum, C.c_double))
    return vol, psum


class Cell:
    """What crystal%shortest reads (crystalmod@proc.f90:1056-1085): isortho, isortho_del, m_x2c, m_x2xr, m_xr2c
    and the Cartesian Wigner-Seitz neighbour vectors ws_ineighc(3, ws_nf)."""

    def __init__(self, x2c, x2xr=None, xr2c=None, ws=None, isortho=None, isortho_del=False):
        self.x2c = np.asarray(x2c, dtype=np.float64)
        off = self.x2c - np.diag(np.diag(self.x2c))
        self.isortho = bool(np.all(off == 0.0)) if isortho is None else bool(isortho)
        self.isortho_del = bool(isortho_del)
        self.x2xr = np.eye(3) if x2xr is None else np.asarray(x2xr, dtype=np.float64)
        self.xr2c = self.x2c.copy() if xr2c is None else np.asarray(xr2c, dtype=np.float64)
        self.ws = np.zeros((3, 0), order="F") if ws is None else np.asfortranarray(ws, dtype=np.float64)

    def args(self):
        self._keep = (_m33(self.x2c), _m33(self.x2xr), _m33(self.xr2c), self.ws)
        a, b, c, w = self._keep
        return (C.c_int(int(self.isortho)), C.c_int(int(self.isortho_del)), _p(a, C.c_double), _p(b, C.c_double),
                _p(c, C.c_double), C.c_int(w.shape[1]), _p(w, C.c_double))


def multipoles_bader(idg, xattr, lmax, fint, cell: Cell, omega):
    """mpole[(lmax+1)^2, nattr] of integration@proc.f90:1338-1360 (Bader / isosurface branch)."""
    idg = _i32(idg)
    fint = _f64(fint)
    n = np.array(idg.shape, dtype=np.int32)
    xattr = _f64(np.asarray(xattr, dtype=np.float64).reshape(3, -1))
    nattr = xattr.shape[1]
    mp = np.zeros(((lmax + 1) ** 2, nattr), order="F")
    lib().orc_multipoles_bader(_p(idg, C.c_int), _p(n, C.c_int), C.c_int(nattr), _p(xattr, C.c_double), C.c_int(lmax),
                               _p(fint, C.c_double), *cell.args(), C.c_double(omega), _p(mp, C.c_double))
    return mp


def multipoles_weighted(w, xattr_m, lmax, fint, cell: Cell, omega):
    """One column mpole(:, m) of the YT branch (integration@proc.f90:1316-1336, :1360) from the weights of basin m."""
    w = _f64(w)
    fint = _f64(fint)
    n = np.array(w.shape, dtype=np.int32)
    xm = np.ascontiguousarray(xattr_m, dtype=np.float64)
    mp = np.zeros((lmax + 1) ** 2)
    lib().orc_multipoles_weighted(_p(w, C.c_double), _p(n, C.c_int), _p(xm, C.c_double), C.c_int(lmax),
                                  _p(fint, C.c_double), *cell.args(), C.c_double(omega), _p(mp, C.c_double))
    return mp


def bader_remap(idg, xattr, cell: Cell, maxattn=None):
    """bader_remap (bader@proc.f90:237-296): (nattn, idg1, iatt[nattn], ilvec[3,nattn])."""
    idg = _i32(idg)
    n = np.array(idg.shape, dtype=np.int32)
    xattr = _f64(np.asarray(xattr, dtype=np.float64).reshape(3, -1))
    nattr = xattr.shape[1]
    maxattn = nattr * 125 if maxattn is None else maxattn
    c2x = _m33(np.linalg.inv(cell.x2c))
    iatt = np.zeros(maxattn, dtype=np.int32)
    ilvec = np.zeros((3, maxattn), dtype=np.int32, order="F")
    idg1 = np.zeros(idg.shape, dtype=np.int32, order="F")
    nattn = C.c_int(0)
    rc = lib().orc_bader_remap(_p(idg, C.c_int), _p(n, C.c_int), C.c_int(nattr), _p(xattr, C.c_double), _p(c2x, C.c_double),
                               *cell.args(), C.c_int(maxattn), C.byref(nattn), _p(iatt, C.c_int), _p(ilvec, C.c_int),
                               _p(idg1, C.c_int))
    if rc != 0:
        raise RuntimeError(f"orc_bader_remap failed rc={rc}")
    return nattn.value, idg1, iatt[: nattn.value].copy(), ilvec[:, : nattn.value].copy()


def yt_remap(d: "YtData", shape, xattr, cell: Cell, maxattn=None):
    """yt_remap (yt@proc.f90:533-594): (nattn, iatt[nattn], ilvec[3,nattn])."""
    xattr = _f64(np.asarray(xattr, dtype=np.float64).reshape(3, -1))
    nattr = xattr.shape[1]
    n = np.array(shape, dtype=np.int32)
    maxattn = nattr * 125 if maxattn is None else maxattn
    c2x = _m33(np.linalg.inv(cell.x2c))
    iatt = np.zeros(maxattn, dtype=np.int32)
    iatt[:nattr] = np.arange(1, nattr + 1)
    ilvec = np.zeros((3, maxattn), dtype=np.int32, order="F")
    nattn = C.c_int(nattr)
    for ib in range(1, nattr + 1):
        w = _f64(yt_weights(d, ib, tuple(int(v) for v in shape)))
        xi = np.ascontiguousarray(xattr[:, ib - 1])
        rc = lib().orc_yt_remap_basin(_p(w, C.c_double), _p(n, C.c_int), C.c_int(nattr), C.c_int(ib), _p(xi, C.c_double),
                                      _p(c2x, C.c_double), *cell.args(), C.c_int(maxattn), C.byref(nattn), _p(iatt, C.c_int),
                                      _p(ilvec, C.c_int))
        if rc != 0:
            raise RuntimeError(f"orc_yt_remap_basin failed rc={rc}")
    return nattn.value, iatt[: nattn.value].copy(), ilvec[:, : nattn.value].copy()


class AtomicGrids:
    """The grid1 objects of the species (grid1mod.f90): logarithmic radial grids r(i) = a exp(b (i-1)) with the atomic
    density f(i), g%rmax, and the cutoff min(cutrad(z), g%rmax) used by promolecular_atom (crystalmod@env.f90:671-684)."""

    def __init__(self, tables):
        """tables: list of dicts with keys a, b, ngrid, f (array), optional rcut."""
        self.nspc = len(tables)
        self.ngrid = np.array([t["ngrid"] for t in tables], dtype=np.int32)
        self.off = np.concatenate([[0], np.cumsum(self.ngrid)[:-1]]).astype(np.int32)
        self.a = np.array([t["a"] for t in tables], dtype=np.float64)
        self.b = np.array([t["b"] for t in tables], dtype=np.float64)
        rs = [t["a"] * np.exp(t["b"] * np.arange(t["ngrid"])) for t in tables]
        self.rtab = np.concatenate(rs).astype(np.float64)
        self.ftab = np.concatenate([np.asarray(t["f"], dtype=np.float64) for t in tables])
        self.rmax = np.array([r[-1] for r in rs], dtype=np.float64)           # g%rmax = r(ngrid)
        self.rcut = np.array([min(t.get("rcut", r[-1]), r[-1]) for t, r in zip(tables, rs)], dtype=np.float64)

    def args(self):
        return (C.c_int(self.nspc), _p(self.ngrid, C.c_int), _p(self.off, C.c_int), _p(self.a, C.c_double), _p(self.b, C.c_double),
                _p(self.rmax, C.c_double), _p(self.rcut, C.c_double), _p(self.rtab, C.c_double), _p(self.ftab, C.c_double))

    def interp(self, isp, r0):
        lib().orc_grid1_interp_value.restype = C.c_double
        o, ng = int(self.off[isp]), int(self.ngrid[isp])
        rt, ft = self.rtab[o:o + ng].copy(), self.ftab[o:o + ng].copy()
        return float(lib().orc_grid1_interp_value(C.c_int(ng), C.c_double(self.a[isp]), C.c_double(self.b[isp]),
                                                  C.c_double(self.rmax[isp]), _p(rt, C.c_double), _p(ft, C.c_double), C.c_double(r0)))


def promolecular_grid(n, x2c, atoms, ispc, grids: AtomicGrids, infrag=None):
    """promolecular_array3 (crystalmod@complex.f90:436-470): f[n1,n2,n3]; ispc 1-based species of every atom."""
    n = np.array(n, dtype=np.int32)
    xat = _f64(np.asarray(atoms, dtype=np.float64).T)
    isp = np.ascontiguousarray(ispc, dtype=np.int32)
    fr = None if infrag is None else np.ascontiguousarray(infrag, dtype=np.uint8)
    out = np.zeros(tuple(int(v) for v in n), order="F")
    x2cf = _m33(x2c)
    lib().orc_promolecular_grid(_p(n, C.c_int), _p(x2cf, C.c_double), C.c_int(xat.shape[1]), _p(xat, C.c_double), _p(isp, C.c_int),
                                *grids.args(), None if fr is None else _p(fr, C.c_ubyte), _p(out, C.c_double))
    return out


def hirshfeld_fields(promol, x2c, atoms, ispc, grids: AtomicGrids, fields, omega, domask=None):
    """intgrid_hirshfeld_fields (integration@proc.f90:1552-1596): (vol[nat], psum[nat, nprop])."""
    promol = _f64(promol)
    n = np.array(promol.shape, dtype=np.int32)
    xat = _f64(np.asarray(atoms, dtype=np.float64).T)
    nat = xat.shape[1]
    isp = np.ascontiguousarray(ispc, dtype=np.int32)
    dm = None if domask is None else np.ascontiguousarray(domask, dtype=np.uint8)
    fl = [_f64(x) for x in fields]
    arr = (C.POINTER(C.c_double) * max(len(fl), 1))(*[_p(x, C.c_double) for x in fl])
    vol = np.zeros(nat)
    psum = np.zeros((nat, len(fl)), order="F")
    x2cf = _m33(x2c)
    lib().orc_hirshfeld_fields(_p(n, C.c_int), _p(x2cf, C.c_double), C.c_int(nat), _p(xat, C.c_double), _p(isp, C.c_int),
                               *grids.args(), _p(promol, C.c_double), None if dm is None else _p(dm, C.c_ubyte),
                               C.c_int(len(fl)), arr, C.c_double(omega), _p(psum, C.c_double), _p(vol, C.c_double))
    return vol, psum


def rlm_real(v, lmax):
    """genrlm_real(lmax, tosphere(v)) (tools_math@proc.f90:273-306, :381-406)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros((lmax + 1) ** 2)
    lib().orc_rlm_real(_p(v, C.c_double), C.c_int(lmax), _p(out, C.c_double))
    return out


def tricubic_matrix():
    c = np.zeros((64, 64), order="F")
    lib().orc_tricubic_matrix(_p(c, C.c_double))
    return c


def grid_interp_tricubic(f, c2xl, xi):
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    y = C.c_double(0)
    yp = np.zeros(3)
    ypp = np.zeros((3, 3), order="F")
    m = _m33(c2xl)
    lib().orc_grid_interp_tricubic(_p(f, C.c_double), _p(n, C.c_int), _p(m, C.c_double), _p(xi, C.c_double),
                                   C.byref(y), _p(yp, C.c_double), _p(ypp, C.c_double))
    return y.value, yp, ypp


def nci_rdg(f, x2c, nstep=None, x0=None, xmat=None, nuclei_cart=None, want_lam2=False):
    """NCIPLOT grid-mode loop.  Default lattice = the periodic, node-aligned one
    (nci@proc.f90:412-427): nstep = grid n, xmat(:,i) = x2c(:,i)/nstep(i), x0 = 0."""
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    x2c = np.asarray(x2c, dtype=np.float64)
    c2x = np.linalg.inv(x2c)
    nstep = np.array(n if nstep is None else nstep, dtype=np.int32)
    if xmat is None:
        xmat = x2c / nstep.astype(np.float64)[None, :]
    x0 = np.zeros(3) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    nuc = np.zeros((0, 3)) if nuclei_cart is None else np.ascontiguousarray(nuclei_cart, dtype=np.float64)
    shape = (int(nstep[2]), int(nstep[1]), int(nstep[0]))  # (k,j,i), k fastest
    crho = np.zeros(shape, order="F")
    cgrad = np.zeros(shape, order="F")
    lam2 = np.zeros(shape, order="F") if want_lam2 else None
    lib().orc_nci_rdg(_p(f, C.c_double), _p(n, C.c_int), _p(x0, C.c_double), _p(_m33(xmat), C.c_double),
                      _p(nstep, C.c_int), _p(_m33(c2x), C.c_double), _p(_m33(x2c), C.c_double),
                      _p(_m33(c2x), C.c_double), C.c_int(nuc.shape[0]), _p(nuc, C.c_double),
                      _p(crho, C.c_double), _p(cgrad, C.c_double),
                      _p(lam2, C.c_double) if want_lam2 else None)
    return (crho, cgrad, lam2) if want_lam2 else (crho, cgrad)


def promolecular(n, x2c, atoms, z, alpha, nimg=1, rc=0.0):
    n = np.array(n, dtype=np.int32)
    xat = np.ascontiguousarray(atoms, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    f = np.zeros(tuple(int(x) for x in n), order="F")
    lib().orc_promolecular(_p(n, C.c_int), _p(_m33(x2c), C.c_double), C.c_int(xat.shape[0]),
                           _p(xat, C.c_double), _p(z, C.c_double), _p(alpha, C.c_double), C.c_int(nimg),
                           C.c_double(rc), _p(f, C.c_double))
    return f


# ifformat_as_ft_* codes of the reference (param.F90:225-236)
FT_CODES = {"x": 33, "y": 34, "z": 35, "xx": 36, "xy": 37, "xz": 38, "yy": 39, "yz": 40, "zz": 41,
            "grad": 42, "lap": 43, "pot": 44}


def fft_derivative(f, x2c, what):
    """grid3%fft (grid3mod@proc.f90:1757-1872); `what` is a key of FT_CODES or the integer code."""
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    iff = FT_CODES[what] if isinstance(what, str) else int(what)
    out = np.zeros(f.shape, order="F")
    rc = lib().orc_fft_derivative(_p(f, C.c_double), _p(n, C.c_int), _p(_m33(x2c), C.c_double), C.c_int(iff),
                                  _p(out, C.c_double))
    if rc:
        raise ValueError(f"orc_fft_derivative: bad code {iff}")
    return out


def grid_interp_trilinear(f, xi):
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    y = C.c_double(0)
    lib().orc_grid_interp_trilinear(_p(f, C.c_double), _p(n, C.c_int), _p(xi, C.c_double), C.byref(y))
    return y.value


def nci_rdg_fourier(f, x2c, nstep=None, x0=None, xmat=None, derived=None):
    """NCIPLOT loop with FOURIER interpolation (nci@proc.f90:527-565).  derived = (|grad|, Hxx, Hyy, Hzz)
    grids; computed with fft_derivative when omitted."""
    f = _f64(f)
    n = np.array(f.shape, dtype=np.int32)
    x2c = np.asarray(x2c, dtype=np.float64)
    c2x = np.linalg.inv(x2c)
    if derived is None:
        derived = tuple(fft_derivative(f, x2c, w) for w in ("grad", "xx", "yy", "zz"))
    fg, fxx, fyy, fzz = (_f64(d) for d in derived)
    nstep = np.array(n if nstep is None else nstep, dtype=np.int32)
    if xmat is None:
        xmat = x2c / nstep.astype(np.float64)[None, :]
    x0 = np.zeros(3) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    shape = (int(nstep[2]), int(nstep[1]), int(nstep[0]))
    crho = np.zeros(shape, order="F")
    cgrad = np.zeros(shape, order="F")
    lib().orc_nci_rdg_fourier(_p(f, C.c_double), _p(fg, C.c_double), _p(fxx, C.c_double), _p(fyy, C.c_double),
                              _p(fzz, C.c_double), _p(n, C.c_int), _p(x0, C.c_double), _p(_m33(xmat), C.c_double),
                              _p(nstep, C.c_int), _p(_m33(c2x), C.c_double), _p(_m33(c2x), C.c_double),
                              _p(crho, C.c_double), _p(cgrad, C.c_double))
    return crho, cgrad


# ---------------------------------------------------------------------------
# Formatted-text grids: the numeric block of read_cube (grid3mod@proc.f90:512-568) and read_vasp (:842-913).
# A list-directed READ converts every field with correct rounding; Python's float() does the same (David Gay's
# algorithm), so this restatement is exact.  Pure-Python loop: small cases only.
# ---------------------------------------------------------------------------
import re as _re

_EXP_LETTER = _re.compile(r"[dDqQ]")
_BARE_EXP = _re.compile(r"^([+-]?(?:\d+\.?\d*|\.\d+))([+-]\d+)$")


def fortran_float(tok: str) -> float:
    """One list-directed numeric field: E/D/Q exponent letters or a bare signed exponent ("1.5-03")."""
    t = _EXP_LETTER.sub("E", tok)
    m = _BARE_EXP.match(t)
    if m:
        t = m.group(1) + "E" + m.group(2)
    return float(t)


def parse_text_grid(text, n, order=0, divisor=1.0):
    """order 0: (((f(i,j,k),i=1,n1),j=1,n2),k=1,n3) (read_vasp); order 1: (((f(i,j,k),k=1,n3),j=1,n2),i=1,n1)
    (read_cube).  Returns f[n1,n2,n3] (Fortran order) = value / divisor, and the byte offset after the last field."""
    if isinstance(text, bytes):
        text = text.decode("ascii")
    nv = int(n[0]) * int(n[1]) * int(n[2])
    vals = np.empty(nv)
    end = 0
    it = _re.finditer(r"[^ \t\r\n,]+", text)
    for q in range(nv):
        m = next(it)
        vals[q] = fortran_float(m.group(0))
        end = m.end()
    if divisor != 1.0:
        vals = vals / divisor
    if order == 0:
        f = vals.reshape(tuple(int(x) for x in n), order="F")
    else:
        f = np.asfortranarray(vals.reshape(tuple(int(x) for x in n), order="C"))
    return f, end


# ---------------------------------------------------------------------------
# Formatted output of grid values: writegrid_cube (crystalmod@write.f90:3556-3565) and NCIPLOT's write_cube_body
# (nci@proc.f90:916-932).  Restates the Ew.dE3 edit descriptor with scale factor k (Fortran 2018 13.7.2.3.3) on top of
# Python's decimal module (exact binary -> decimal, round-half-even like the Fortran run-time library).
# NOT pinned against a Fortran compiler (none in the image): the asterisk rule for fields that do not fit and the
# optional leading zero follow the standard's text.
# ---------------------------------------------------------------------------
import decimal as _dec
import math as _math


def fortran_e(x: float, w: int, d: int, k: int) -> str:
    """One value in Ew.dE3 with scale factor k (0 or 1), right-justified; asterisks if it does not fit."""
    if _math.isnan(x):
        s = "NaN"
    elif _math.isinf(x):
        neg = x < 0
        s = ("-" if neg else "") + ("Infinity" if w >= 8 + (1 if neg else 0) else "Inf")
    else:
        S = d + (1 if k == 1 else 0)
        neg = _math.copysign(1.0, x) < 0
        if x == 0.0:
            N, e10 = 0, 0
            pexp = 0
        else:
            with _dec.localcontext() as c:
                c.prec = 1200
                D = _dec.Decimal(abs(x))
                e10 = D.adjusted()
                N = int(D.scaleb(-(e10 - S + 1)).quantize(_dec.Decimal(1), rounding=_dec.ROUND_HALF_EVEN))
            if N == 10 ** S:
                N //= 10
                e10 += 1
            pexp = e10 if k == 1 else e10 + 1
        dig = str(N).rjust(S, "0")
        body = (dig[0] + "." + dig[1:]) if k == 1 else ("0." + dig)
        s = ("-" if neg else "") + body + "E" + ("-" if pexp < 0 else "+") + "%03d" % abs(pexp)
        if len(s) > w and k == 0:
            s = s.replace("0.", ".", 1)   # the zero before the point is optional
    return "*" * w if len(s) > w else s.rjust(w)


def format_text_grid(f, layout, w, d, k, ishift=(0, 0, 0)):
    """layout 0: rows along index 1 of f as stored (NCI crho(k,j,i)); layout 1: cube order of f(i,j,k) with ishift.
    Every value " " + Ew.dE3, 6 per line, new line after each row.  A line with fewer than 6 values ends with a blank:
    when the data list ends inside the group 6(" ",E...) the literal that precedes the next data edit descriptor is still
    written (Fortran format control); pinned by the reference's own outputs, tests/005_plot/ref/029_cube_precise_0[12].cube."""
    f = _f64(f)
    n1, n2, n3 = f.shape
    out = []
    if layout == 0:
        for c in range(n3):
            for b in range(n2):
                row = [" " + fortran_e(float(f[a, b, c]), w, d, k) for a in range(n1)]
                out += ["".join(row[q:q + 6]) + (" \n" if len(row[q:q + 6]) < 6 else "\n") for q in range(0, n1, 6)]
    else:
        for iix in range(n1):
            ix = (iix + ishift[0]) % n1
            for iiy in range(n2):
                iy = (iiy + ishift[1]) % n2
                row = [" " + fortran_e(float(f[ix, iy, (iiz + ishift[2]) % n3]), w, d, k) for iiz in range(n3)]
                out += ["".join(row[q:q + 6]) + (" \n" if len(row[q:q + 6]) < 6 else "\n") for q in range(0, n3, 6)]
    return "".join(out).encode()


def voronoi_grid(n, x2c, atoms, reach_cells=3):
    """nearest_atom_grid (crystalmod@proc.f90:1138-1167; voronoi_grid, hirshfeld@proc.f90:93-122) by brute force in numpy:
    for every grid node x = ((i-1)/n1, (j-1)/n2, (k-1)/n3) the 1-based id of the nearest atom over the lattice
    translations -reach_cells..reach_cells, the distance as the reference computes it (Cartesian difference, norm2).
    Returns (idg[n1,n2,n3] int32 with ties resolved to the LOWER id, gap[n1,n2,n3] = relative distance gap to the nearest
    OTHER atom: the reference's choice at gap == 0 follows the traversal order of list_near_atoms, which is not restated)."""
    n = tuple(int(v) for v in n)
    x2c = np.asarray(x2c, dtype=np.float64)
    at = np.asarray(atoms, dtype=np.float64)
    at = at - np.floor(at)
    g = np.stack(np.meshgrid(*[np.arange(m) / m for m in n], indexing="ij"), -1).reshape(-1, 3)
    pc = g @ x2c.T
    r = range(-reach_cells, reach_cells + 1)
    shifts = np.array([[a, b, c] for a in r for b in r for c in r], dtype=np.float64)
    best = np.full((len(at), len(g)), np.inf)
    for ia, a in enumerate(at):
        img = (a[None, :] + shifts) @ x2c.T
        for k0 in range(0, len(img), 64):
            d = pc[None, :, :] - img[k0:k0 + 64, None, :]
            d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
            best[ia] = np.minimum(best[ia], d2.min(axis=0))
    order = np.argsort(best, axis=0, kind="stable")
    idg = (order[0] + 1).astype(np.int32)
    d1 = np.sqrt(best[order[0], np.arange(len(g))])
    d2 = np.sqrt(best[order[1], np.arange(len(g))]) if len(at) > 1 else np.full(len(g), np.inf)
    gap = np.ones(len(g)) if len(at) == 1 else (d2 - d1) / np.maximum(d2, 1e-300)   # a single atom has no competitor
    return idg.reshape(n), gap.reshape(n)
