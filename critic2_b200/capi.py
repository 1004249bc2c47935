"""ctypes binding of the C ABI declared in include/critic2_gpu.h (critic2_b200/libcritic2_gpu.so).

This is the reference-side binding a maintainer would write in Fortran (fortran/critic2_gpu.f90);
Python is used here only because the image has no Fortran compiler.  There is no CPU fallback: if the
shared library is missing or no CUDA device is usable every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# C2G_LIB_PATH: another build of the same library (kernel-variant experiments, tools/); the product is the in-tree file
LIB_PATH = os.environ.get("C2G_LIB_PATH") or os.path.join(_HERE, "libcritic2_gpu.so")
_lib = None

BADER_FAST, BADER_EXACT = 0, 1
# ifformat_as_ft_* codes of the reference (param.F90:225-236)
FT_CODES = {"x": 33, "y": 34, "z": 35, "xx": 36, "xy": 37, "xz": 38, "yy": 39, "yz": 40, "zz": 41,
            "grad": 42, "lap": 43, "pot": 44}
ORDER_INDEX, ORDER_SCAN = 0, 1

EXPORTS = [
    "c2g_init", "c2g_init_devices", "c2g_nccl_unique_id", "c2g_init_multi", "c2g_finalize", "c2g_last_error", "c2g_describe",
    "c2g_grid_upload", "c2g_grid_upload_slab", "c2g_slab_range", "c2g_slab_bounds_query", "c2g_grid_alloc", "c2g_grid_download", "c2g_grid_download_slab", "c2g_grid_free", "c2g_grid_promolecular",
    "c2g_bader_assign", "c2g_basins_maxima", "c2g_basins_counts", "c2g_basins_set_map", "c2g_basins_labels",
    "c2g_basins_relabel", "c2g_basins_nattr", "c2g_basins_free", "c2g_basins_stats", "c2g_integrate", "c2g_integrate_multipoles", "c2g_promolecular_grid", "c2g_hirshfeld_integrate", "c2g_basins_remap", "c2g_yt_build",
    "c2g_yt_weights", "c2g_yt_export", "c2g_voronoi_grid", "c2g_basins_weight_grid", "c2g_yt_isosurface", "c2g_nci_rdg", "c2g_nci_rdg_resident", "c2g_fft_derivative", "c2g_nci_rdg_fourier", "c2g_nci_range", "c2g_grid_upload_async", "c2g_basins_labels_async", "c2g_grid_parse_text", "c2g_grid_format_text", "c2g_profile_enable", "c2g_profile_count",
    "c2g_profile_get", "c2g_profile_reset", "c2g_launch_count", "c2g_flush_l2", "c2g_synchronize", "c2g_timer_start", "c2g_timer_stop",
]


class C2GError(RuntimeError):
    pass


def load():
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise C2GError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.c2g_last_error.restype = C.c_char_p
        lib.c2g_describe.restype = C.c_char_p
        lib.c2g_launch_count.restype = C.c_longlong
        lib.c2g_finalize.restype = None
        lib.c2g_basins_free.restype = None
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _m33(m):
    return np.asfortranarray(np.asarray(m, dtype=np.float64)).ravel(order="F").copy()


def slab_bounds(n3, nranks, rank):
    """z-slab [zlo, zhi) owned by `rank` (host arithmetic only, no GPU needed)."""
    a, b = C.c_int(0), C.c_int(0)
    rc = load().c2g_slab_bounds_query(C.c_int(int(n3)), C.c_int(int(nranks)), C.c_int(int(rank)), C.byref(a), C.byref(b))
    if rc != 0:
        raise C2GError(f"c2g_slab_bounds_query: status {rc}")
    return a.value, b.value


class Context:
    """One c2g_context (one GPU)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_uid=None, ngpus=None):
        """device: one GPU.  rank / nranks / nccl_uid: one process per GPU (c2g_init_multi).  ngpus: ONE process driving
        devices 0..ngpus-1 (c2g_init_devices) -- every method then takes and returns whole arrays."""
        self.lib = load()
        self.h = C.c_void_p()
        if ngpus is not None:
            rc = self.lib.c2g_init_devices(C.c_int(int(ngpus)), C.byref(self.h))
        elif nranks > 1:
            rc = self.lib.c2g_init_multi(C.c_int(device), C.c_int(rank), C.c_int(nranks), nccl_uid, C.byref(self.h))
        else:
            rc = self.lib.c2g_init(C.c_int(device), C.byref(self.h))
        if rc != 0:
            msg = self.lib.c2g_last_error(self.h).decode() if self.h else "no usable CUDA device"
            raise C2GError(f"c2g_init failed (status {rc}): {msg}")

    def _chk(self, rc):
        if rc != 0:
            raise C2GError(f"status {rc}: {self.lib.c2g_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.c2g_finalize(self.h)
            self.h = C.c_void_p()

    def describe(self):
        return self.lib.c2g_describe(self.h).decode()

    # ---- grids ----
    def upload(self, f):
        f = np.asfortranarray(f, dtype=np.float64)
        n = np.array(f.shape, dtype=np.int32)
        h = C.c_int(-1)
        self._chk(self.lib.c2g_grid_upload(self.h, _p(f, C.c_double), _p(n, C.c_int), C.byref(h)))
        return h.value

    def slab_range(self, n3):
        a, b = C.c_int(0), C.c_int(0)
        self._chk(self.lib.c2g_slab_range(self.h, C.c_int(int(n3)), C.byref(a), C.byref(b)))
        return a.value, b.value

    def upload_slab(self, fslab, n):
        """fslab: this rank's planes f[:, :, zlo:zhi] (Fortran order)."""
        fslab = np.asfortranarray(fslab, dtype=np.float64)
        n = np.array(n, dtype=np.int32)
        h = C.c_int(-1)
        self._chk(self.lib.c2g_grid_upload_slab(self.h, _p(fslab, C.c_double), _p(n, C.c_int), C.byref(h)))
        return h.value

    def alloc(self, n):
        n = np.array(n, dtype=np.int32)
        h = C.c_int(-1)
        self._chk(self.lib.c2g_grid_alloc(self.h, _p(n, C.c_int), C.byref(h)))
        return h.value

    def download(self, h, shape):
        f = np.zeros(tuple(int(x) for x in shape), order="F")
        self._chk(self.lib.c2g_grid_download(self.h, C.c_int(h), _p(f, C.c_double)))
        return f

    def free(self, h):
        self._chk(self.lib.c2g_grid_free(self.h, C.c_int(h)))

    def promolecular(self, h, x2c, atoms, z, alpha, nimg=1, rc=0.0):
        xat = np.ascontiguousarray(atoms, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        self._chk(self.lib.c2g_grid_promolecular(self.h, C.c_int(h), _p(_m33(x2c), C.c_double), C.c_int(xat.shape[0]),
                                                 _p(xat, C.c_double), _p(z, C.c_double), _p(alpha, C.c_double),
                                                 C.c_int(nimg), C.c_double(rc)))

    # ---- BADER ----
    def bader_assign(self, h, car2lat, lat_i_dist, algo=BADER_FAST, order=ORDER_INDEX):
        nmax = C.c_int(0)
        res = C.c_void_p()
        lid = np.ascontiguousarray(lat_i_dist, dtype=np.float64).ravel()
        self._chk(self.lib.c2g_bader_assign(self.h, C.c_int(h), _p(_m33(car2lat), C.c_double), _p(lid, C.c_double),
                                            C.c_int(algo), C.c_int(order), C.byref(nmax), C.byref(res)))
        return Basins(self, res, nmax.value)

    def yt_build(self, h, vec, area):
        vec = np.asfortranarray(np.asarray(vec, dtype=np.int32).reshape(-1, 3).T)
        area = np.ascontiguousarray(area, dtype=np.float64)
        nmax = C.c_int(0)
        res = C.c_void_p()
        self._chk(self.lib.c2g_yt_build(self.h, C.c_int(h), C.c_int(vec.shape[1]), _p(vec, C.c_int), _p(area, C.c_double),
                                        C.byref(nmax), C.byref(res)))
        return Basins(self, res, nmax.value)

    def integrate(self, basins, fieldhandles, omega):
        fh = np.array(list(fieldhandles), dtype=np.int32)
        nattr = basins.nattr
        psum = np.zeros((nattr, len(fh)), order="F")
        vol = np.zeros(nattr)
        self._chk(self.lib.c2g_integrate(self.h, basins.h, C.c_int(len(fh)), _p(fh, C.c_int), C.c_double(omega),
                                         _p(psum, C.c_double), _p(vol, C.c_double)))
        return vol, psum

    def integrate_multipoles(self, basins, fieldhandle, lmax, xattr, x2c, omega, x2xr=None, xr2c=None, ws=None,
                             isortho=None, isortho_del=False, domask=None):
        """mpole[(lmax+1)^2, nattr] (integration@proc.f90:1302-1361).  xattr (3, nattr) crystallographic; the cell
        arguments are those of crystal%shortest: an orthogonal m_x2c alone, or x2xr / xr2c / ws (3, ws_nf) Cartesian."""
        x2c = np.asarray(x2c, dtype=np.float64)
        if isortho is None:
            isortho = bool(np.all(x2c - np.diag(np.diag(x2c)) == 0.0))
        xa = np.asfortranarray(np.asarray(xattr, dtype=np.float64).reshape(3, -1))
        nattr = basins.nattr
        if xa.shape[1] != nattr:
            raise C2GError(f"xattr has {xa.shape[1]} columns, the basins have {nattr} attractors")
        a = _m33(x2c)
        b = _m33(np.eye(3) if x2xr is None else x2xr)
        c = _m33(x2c if xr2c is None else xr2c)
        w = np.zeros((3, 0), order="F") if ws is None else np.asfortranarray(ws, dtype=np.float64)
        dm = None if domask is None else np.ascontiguousarray(domask, dtype=np.uint8)
        mp = np.zeros(((lmax + 1) ** 2, nattr), order="F")
        self._chk(self.lib.c2g_integrate_multipoles(
            self.h, basins.h, C.c_int(int(fieldhandle)), C.c_int(int(lmax)), _p(xa, C.c_double),
            None if dm is None else _p(dm, C.c_ubyte), C.c_int(int(isortho)), C.c_int(int(isortho_del)), _p(a, C.c_double),
            _p(b, C.c_double), _p(c, C.c_double), C.c_int(w.shape[1]), _p(w, C.c_double) if w.shape[1] else None,
            C.c_double(omega), _p(mp, C.c_double)))
        return mp

    def basins_remap(self, basins, xattr, x2c, shape=None, x2xr=None, xr2c=None, ws=None, isortho=None, isortho_del=False,
                     maxattn=None, want_idg1=True):
        """bader_remap / yt_remap: (nattn, idg1 or None, iatt[nattn], ilvec[3,nattn]).  shape: (n1, n2, nz_owned) of idg1."""
        x2c = np.asarray(x2c, dtype=np.float64)
        if isortho is None:
            isortho = bool(np.all(x2c - np.diag(np.diag(x2c)) == 0.0))
        xa = np.asfortranarray(np.asarray(xattr, dtype=np.float64).reshape(3, -1))
        nattr = basins.nattr
        a, b, c = _m33(x2c), _m33(np.eye(3) if x2xr is None else x2xr), _m33(x2c if xr2c is None else xr2c)
        c2x = _m33(np.linalg.inv(x2c))
        w = np.zeros((3, 0), order="F") if ws is None else np.asfortranarray(ws, dtype=np.float64)
        cap = max(nattr * 27, 1) if maxattn is None else maxattn
        idg1 = np.zeros(tuple(int(v) for v in shape), dtype=np.int32, order="F") if (want_idg1 and shape is not None) else None
        while True:
            iatt = np.zeros(cap, dtype=np.int32)
            ilvec = np.zeros((3, cap), dtype=np.int32, order="F")
            nattn = C.c_int(0)
            rc = self.lib.c2g_basins_remap(
                self.h, basins.h, _p(xa, C.c_double), _p(c2x, C.c_double), C.c_int(int(isortho)), C.c_int(int(isortho_del)),
                _p(a, C.c_double), _p(b, C.c_double), _p(c, C.c_double), C.c_int(w.shape[1]),
                _p(w, C.c_double) if w.shape[1] else None, C.c_int(cap), C.byref(nattn), _p(iatt, C.c_int), _p(ilvec, C.c_int),
                None if idg1 is None else _p(idg1, C.c_int))
            if rc == 6 and nattn.value > cap and maxattn is None:   # C2G_ERR_OVERFLOW: retry with the size needed
                cap = nattn.value
                continue
            self._chk(rc)
            return nattn.value, idg1, iatt[: nattn.value].copy(), ilvec[:, : nattn.value].copy()

    # ---- HIRSHFELD ----
    @staticmethod
    def _species_args(tab):
        """tab: dict of arrays ngrid, off, a, b, rmax, rcut, rtab, ftab (see include/critic2_gpu.h)."""
        ng = np.ascontiguousarray(tab["ngrid"], dtype=np.int32)
        off = np.ascontiguousarray(tab["off"], dtype=np.int32)
        arrs = [np.ascontiguousarray(tab[k], dtype=np.float64) for k in ("a", "b", "rmax", "rcut", "rtab", "ftab")]
        keep = (ng, off, *arrs)
        return keep, (C.c_int(len(ng)), _p(ng, C.c_int), _p(off, C.c_int), *[_p(x, C.c_double) for x in arrs])

    def promolecular_grid(self, n, x2c, atoms, ispc, tab, infrag=None):
        """promolecular_array3 (crystalmod@complex.f90:436-470) as a resident grid; returns the handle."""
        n = np.array(n, dtype=np.int32)
        xat = np.asfortranarray(np.asarray(atoms, dtype=np.float64).T)
        isp = np.ascontiguousarray(ispc, dtype=np.int32)
        fr = None if infrag is None else np.ascontiguousarray(infrag, dtype=np.uint8)
        keep, sargs = self._species_args(tab)
        x2cf = _m33(x2c)
        h = C.c_int(-1)
        self._chk(self.lib.c2g_promolecular_grid(self.h, _p(n, C.c_int), _p(x2cf, C.c_double), C.c_int(xat.shape[1]),
                                                 _p(xat, C.c_double), _p(isp, C.c_int), *sargs,
                                                 None if fr is None else _p(fr, C.c_ubyte), C.byref(h)))
        return h.value

    def voronoi_grid(self, n, x2c, atoms):
        """voronoi_grid (hirshfeld@proc.f90:93-122): Basins whose labels are the nearest-atom ids; the map is already set."""
        nn = np.array(n, dtype=np.int32)
        xat = np.ascontiguousarray(atoms, dtype=np.float64)
        res = C.c_void_p()
        self._chk(self.lib.c2g_voronoi_grid(self.h, _p(nn, C.c_int), _p(_m33(x2c), C.c_double), C.c_int(len(xat)), _p(xat, C.c_double),
                                            C.byref(res)))
        b = Basins(self, res, len(xat))
        b.nattr = len(xat)
        return b

    def hirshfeld_integrate(self, hpromol, x2c, atoms, ispc, tab, fieldhandles, omega, domask=None):
        """intgrid_hirshfeld_fields (integration@proc.f90:1552-1596): (vol[nat], psum[nat, nprop])."""
        xat = np.asfortranarray(np.asarray(atoms, dtype=np.float64).T)
        nat = xat.shape[1]
        isp = np.ascontiguousarray(ispc, dtype=np.int32)
        dm = None if domask is None else np.ascontiguousarray(domask, dtype=np.uint8)
        fh = np.array(list(fieldhandles), dtype=np.int32)
        keep, sargs = self._species_args(tab)
        x2cf = _m33(x2c)
        psum = np.zeros((nat, len(fh)), order="F")
        vol = np.zeros(nat)
        self._chk(self.lib.c2g_hirshfeld_integrate(self.h, C.c_int(int(hpromol)), _p(x2cf, C.c_double), C.c_int(nat),
                                                   _p(xat, C.c_double), _p(isp, C.c_int), *sargs,
                                                   None if dm is None else _p(dm, C.c_ubyte), C.c_int(len(fh)), _p(fh, C.c_int),
                                                   C.c_double(omega), _p(psum, C.c_double), _p(vol, C.c_double)))
        return vol, psum

    # ---- NCIPLOT ----
    def nci_range(self, nstep1):
        """Lattice rows [ilo, ihi) computed by this rank (the whole range on a single GPU)."""
        a, b = C.c_int(0), C.c_int(0)
        self._chk(self.lib.c2g_nci_range(self.h, C.c_int(int(nstep1)), C.byref(a), C.byref(b)))
        return a.value, b.value

    def nci_rdg(self, h, x2c, n, nstep=None, x0=None, xmat=None, nuclei_cart=None, c2xl=None):
        x2c = np.asarray(x2c, dtype=np.float64)
        c2x = np.linalg.inv(x2c)
        nstep = np.array(n if nstep is None else nstep, dtype=np.int32)
        if xmat is None:
            xmat = x2c / nstep.astype(np.float64)[None, :]
        x0 = np.zeros(3) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        nuc = np.zeros((0, 3)) if nuclei_cart is None else np.ascontiguousarray(nuclei_cart, dtype=np.float64)
        ilo, ihi = self.nci_range(nstep[0])
        shape = (int(nstep[2]), int(nstep[1]), max(ihi - ilo, 0))
        crho = np.zeros(shape, order="F")
        cgrad = np.zeros(shape, order="F")
        self._chk(self.lib.c2g_nci_rdg(self.h, C.c_int(h), _p(x0, C.c_double), _p(_m33(xmat), C.c_double),
                                       _p(nstep, C.c_int), _p(_m33(c2x), C.c_double), _p(_m33(x2c), C.c_double),
                                       _p(_m33(c2x if c2xl is None else c2xl), C.c_double), C.c_int(nuc.shape[0]),
                                       _p(nuc, C.c_double), _p(crho, C.c_double), _p(cgrad, C.c_double)))
        return crho, cgrad

    def nci_rdg_resident(self, h, x2c, n, nstep=None):
        x2c = np.asarray(x2c, dtype=np.float64)
        c2x = np.linalg.inv(x2c)
        nstep = np.array(n if nstep is None else nstep, dtype=np.int32)
        xmat = x2c / nstep.astype(np.float64)[None, :]
        x0 = np.zeros(3)
        hr, hg = C.c_int(-1), C.c_int(-1)
        self._chk(self.lib.c2g_nci_rdg_resident(self.h, C.c_int(h), _p(x0, C.c_double), _p(_m33(xmat), C.c_double),
                                                _p(nstep, C.c_int), _p(_m33(c2x), C.c_double), _p(_m33(x2c), C.c_double),
                                                _p(_m33(c2x), C.c_double), C.c_int(0), None, C.byref(hr), C.byref(hg)))
        return hr.value, hg.value

    # ---- FFT-derived fields (grid3%fft) and NCIPLOT FOURIER mode ----
    def fft_derivative(self, h, x2c, what):
        """New resident grid = grid3%fft(h, ifformat_as_ft_<what>) (grid3mod@proc.f90:1757-1872)."""
        iff = FT_CODES[what] if isinstance(what, str) else int(what)
        out = C.c_int(-1)
        self._chk(self.lib.c2g_fft_derivative(self.h, C.c_int(h), C.c_int(iff), _p(_m33(x2c), C.c_double), C.byref(out)))
        return out.value

    def nci_rdg_fourier(self, handles, x2c, n, nstep=None, x0=None, xmat=None):
        """handles = (rho, |grad rho|, Hxx, Hyy, Hzz) resident grids (nci@proc.f90:527-565)."""
        x2c = np.asarray(x2c, dtype=np.float64)
        c2x = np.linalg.inv(x2c)
        nstep = np.array(n if nstep is None else nstep, dtype=np.int32)
        if xmat is None:
            xmat = x2c / nstep.astype(np.float64)[None, :]
        x0 = np.zeros(3) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        hh = np.ascontiguousarray(handles, dtype=np.int32)
        ilo, ihi = self.nci_range(nstep[0])
        shape = (int(nstep[2]), int(nstep[1]), max(ihi - ilo, 0))
        crho = np.zeros(shape, order="F")
        cgrad = np.zeros(shape, order="F")
        self._chk(self.lib.c2g_nci_rdg_fourier(self.h, _p(hh, C.c_int), _p(x0, C.c_double), _p(_m33(xmat), C.c_double),
                                               _p(nstep, C.c_int), _p(_m33(c2x), C.c_double), _p(_m33(c2x), C.c_double),
                                               _p(crho, C.c_double), _p(cgrad, C.c_double)))
        return crho, cgrad

    # ---- formatted-text grids (cube / CHGCAR numeric blocks) ----
    def parse_text(self, text: bytes, n, order=0, divisor=1.0):
        """Numbers of a cube (order=1: k fastest) or CHGCAR (order=0: i fastest) block -> new resident grid.
        Returns (handle, bytes consumed, values that took the exact multi-word path on the device)."""
        nn = np.array(n, dtype=np.int32)
        h = C.c_int(-1)
        used = C.c_size_t(0)
        nhost = C.c_longlong(0)
        self._chk(self.lib.c2g_grid_parse_text(self.h, C.c_char_p(text), C.c_size_t(len(text)), _p(nn, C.c_int), C.c_int(order),
                                               C.c_double(divisor), C.byref(h), C.byref(used), C.byref(nhost)))
        return h.value, used.value, nhost.value

    def format_text_size(self, h, layout, width, digits, scale):
        nb = C.c_size_t(0)
        self._chk(self.lib.c2g_grid_format_text(self.h, C.c_int(h), C.c_int(layout), None, C.c_int(width), C.c_int(digits),
                                                C.c_int(scale), None, C.c_size_t(0), C.byref(nb)))
        return nb.value

    def format_text_into(self, h, layout, width, digits, scale, ptr, cap, ishift=None):
        """Step 2 into a caller-owned host buffer (raw pointer, e.g. pinned); returns the number of bytes written."""
        nb = C.c_size_t(0)
        sh = None if ishift is None else _p(np.ascontiguousarray(ishift, dtype=np.int32), C.c_int)
        self._chk(self.lib.c2g_grid_format_text(self.h, C.c_int(h), C.c_int(layout), sh, C.c_int(width), C.c_int(digits),
                                                C.c_int(scale), C.c_void_p(ptr), C.c_size_t(cap), C.byref(nb)))
        return nb.value

    def format_text(self, h, layout, width, digits, scale, ishift=None):
        """Value block of a cube file from a resident grid (see c2g_grid_format_text); returns bytes."""
        n = self.format_text_size(h, layout, width, digits, scale)
        buf = np.empty(n, dtype=np.uint8)
        self.format_text_into(h, layout, width, digits, scale, buf.ctypes.data, n, ishift)
        return buf.tobytes()

    # ---- profiling ----
    def profile_enable(self, on=True):
        self._chk(self.lib.c2g_profile_enable(self.h, C.c_int(1 if on else 0)))

    def profile_reset(self):
        self._chk(self.lib.c2g_profile_reset(self.h))

    def profile(self):
        self.synchronize()
        out = {}
        name = C.create_string_buffer(64)
        ms = C.c_double(0)
        nl = C.c_int(0)
        for i in range(self.lib.c2g_profile_count(self.h)):
            self.lib.c2g_profile_get(self.h, C.c_int(i), name, C.byref(ms), C.byref(nl))
            out[name.value.decode()] = (ms.value, nl.value)
        return out

    def launch_count(self):
        return int(self.lib.c2g_launch_count(self.h))

    def flush_l2(self):
        self._chk(self.lib.c2g_flush_l2(self.h))

    def timer_start(self):
        self._chk(self.lib.c2g_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double(0)
        self._chk(self.lib.c2g_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def upload_ptr(self, ptr, n):
        """Upload from a raw host pointer (e.g. pinned memory)."""
        n = np.array(n, dtype=np.int32)
        h = C.c_int(-1)
        self._chk(self.lib.c2g_grid_upload(self.h, C.c_void_p(ptr), _p(n, C.c_int), C.byref(h)))
        return h.value

    def upload_ptr_async(self, ptr, n):
        """Asynchronous upload from a (pinned) host buffer; see c2g_grid_upload_async."""
        nn = np.array(n, dtype=np.int32)
        h = C.c_int(-1)
        self._chk(self.lib.c2g_grid_upload_async(self.h, C.c_void_p(ptr), _p(nn, C.c_int), C.byref(h)))
        return h.value

    def upload_slab_ptr(self, ptr, n):
        n = np.array(n, dtype=np.int32)
        h = C.c_int(-1)
        self._chk(self.lib.c2g_grid_upload_slab(self.h, C.c_void_p(ptr), _p(n, C.c_int), C.byref(h)))
        return h.value

    def download_slab_ptr(self, h, ptr):
        self._chk(self.lib.c2g_grid_download_slab(self.h, C.c_int(h), C.c_void_p(ptr)))

    def download_ptr(self, h, ptr):
        self._chk(self.lib.c2g_grid_download(self.h, C.c_int(h), C.c_void_p(ptr)))

    def synchronize(self):
        self._chk(self.lib.c2g_synchronize(self.h))


class Basins:
    """c2g_basins: device-resident result of a BADER/YT assignment."""

    def __init__(self, ctx, h, nmax):
        self.ctx, self.h, self.nmax = ctx, h, nmax
        self.nattr = 0

    def maxima(self):
        """(nmax,3) 1-based grid coordinates."""
        p = np.zeros((self.nmax, 3), dtype=np.int32)
        self.ctx._chk(self.ctx.lib.c2g_basins_maxima(self.h, _p(p, C.c_int)))
        return p

    def counts(self):
        c = np.zeros(self.nmax, dtype=np.int64)
        self.ctx._chk(self.ctx.lib.c2g_basins_counts(self.h, _p(c, C.c_longlong)))
        return c

    def set_map(self, nattr, mp):
        mp = np.ascontiguousarray(mp, dtype=np.int32)
        self.ctx._chk(self.ctx.lib.c2g_basins_set_map(self.h, C.c_int(nattr), _p(mp, C.c_int)))
        self.nattr = nattr

    def relabel(self, assigned, nattr_new):
        a = np.ascontiguousarray(assigned, dtype=np.int32)
        self.ctx._chk(self.ctx.lib.c2g_basins_relabel(self.h, C.c_int(len(a)), _p(a, C.c_int), C.c_int(nattr_new)))
        self.nattr = nattr_new

    def labels_ptr_async(self, ptr):
        """Asynchronous download of bas%idg into a (pinned) host buffer; complete after Context.synchronize()."""
        self.ctx._chk(self.ctx.lib.c2g_basins_labels_async(self.h, C.c_void_p(ptr)))

    def labels_ptr(self, ptr):
        """idg (this rank's slab) into a raw host pointer."""
        self.ctx._chk(self.ctx.lib.c2g_basins_labels(self.h, C.c_void_p(ptr)))

    def labels(self, shape):
        idg = np.zeros(tuple(int(x) for x in shape), dtype=np.int32, order="F")
        self.ctx._chk(self.ctx.lib.c2g_basins_labels(self.h, _p(idg, C.c_int)))
        return idg

    def stats(self):
        s = np.zeros(8, dtype=np.int64)
        self.ctx._chk(self.ctx.lib.c2g_basins_stats(self.h, _p(s, C.c_longlong)))
        return s

    def yt_export(self, shape, nvec, full=True):
        """ytdata record (yt.f90:36-45): (nlo, ibasin, iio, inear, fnear), position-indexed like the reference's."""
        nn = int(np.prod(shape))
        nlo = np.zeros(nn, dtype=np.int32); ibasin = np.zeros(nn, dtype=np.int32); iio = np.zeros(nn, dtype=np.int32)
        inear = np.zeros((nvec, nn), dtype=np.int32, order="F") if full else None
        fnear = np.zeros((nvec, nn), dtype=np.float64, order="F") if full else None
        self.ctx._chk(self.ctx.lib.c2g_yt_export(self.h, _p(nlo, C.c_int), _p(ibasin, C.c_int), _p(iio, C.c_int),
                                                 _p(inear, C.c_int) if full else None, _p(fnear, C.c_double) if full else None))
        return nlo, ibasin, iio, inear, fnear

    def yt_weights(self, idb, shape):
        w = np.zeros(tuple(int(x) for x in shape), order="F")
        self.ctx._chk(self.ctx.lib.c2g_yt_weights(self.h, C.c_int(idb), _p(w, C.c_double)))
        return w

    def weight_grid(self, idb):
        """Resident weight field of basin idb (int_cubew, integration@proc.f90:4449-4459); returns a grid handle."""
        h = C.c_int(-1)
        self.ctx._chk(self.ctx.lib.c2g_basins_weight_grid(self.h, C.c_int(idb), C.byref(h)))
        return h.value

    def isosurface(self, isov):
        """yt_isosurface (yt@proc.f90:233-390) on a YT result: (regions Basins, nraw, nattr)."""
        nraw, nattr = C.c_int(0), C.c_int(0)
        res = C.c_void_p()
        self.ctx._chk(self.ctx.lib.c2g_yt_isosurface(self.h, C.c_double(isov), C.byref(nraw), C.byref(nattr), C.byref(res)))
        b = Basins(self.ctx, res, nraw.value)
        b.nattr = nraw.value
        return b, nraw.value, nattr.value

    def free(self):
        if self.h:
            self.ctx.lib.c2g_basins_free(self.h)
            self.h = C.c_void_p()
