#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs through the C ABI, inputs resident in HBM.

One JSON line per path (stdout), each with the algorithmic bytes per grid point of SURVEY.md 8(d), the CUDA-event
time of the call on the library's stream (best of `reps` after a warm-up), achieved GB/s against the measured HBM
peak (MEASURED_PEAKS.json, else the 6650 GB/s fallback) and a size-independent or sub-sampled parity check
against the oracle.  bench.py stays the headline (configs[4]); this file is evidence for the rows next to it.

usage: python tools/bench_paths.py [--quick] [--only=name,name]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import systems as S
import bench
from critic2_b200 import capi
from oracle import oracle as orc

quick = "--quick" in sys.argv
PEAK, PEAK_SRC = bench.measured_peaks()
ctx = capi.Context(0)
ctx.profile_enable(True)


def timed(fn, reps=3, cleanup=None):
    """Best CUDA-event time of fn() on the library stream; cleanup(out) releases the result of every call but the last."""
    out = fn()  # warm-up
    best, prof = 1e30, {}
    for _ in range(reps):
        if cleanup:
            cleanup(out)
        ctx.profile_reset(); ctx.synchronize(); ctx.timer_start()
        out = fn()
        ms = ctx.timer_stop()
        if ms < best:
            best, prof = ms, {k: round(v[0], 3) for k, v in ctx.profile().items()}
    return best, prof, out


def free_all(hs):
    for x in hs:
        ctx.free(x)


def emit(path, config, n, ms, bytes_per_pt, prof, check, extra=None):
    npts = float(np.prod(n))
    gbs = bytes_per_pt * npts / (ms * 1e-3) / 1e9
    rec = {"path": path, "config": config, "grid": list(map(int, n)), "ms": round(ms, 3), "points_per_s": npts / (ms * 1e-3),
           "algorithmic_bytes_per_point": bytes_per_pt, "achieved_GBps": round(gbs, 1), "peak_GBps": PEAK, "peak_source": PEAK_SRC,
           "frac": round(gbs / PEAK, 4), "kernels_ms": prof, "check": check}
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)


# ---- config 2: urea-like tetragonal cell, 256^3: BADER + INTEGRABLE rho and FFT Laplacian ----
def urea():
    N = 128 if quick else 256
    n = (N, N, N)
    x2c = S.cell_x2c(10.52, 10.52, 8.85)
    at, z, al = S.random_atoms(16, 2, x2c, dmin=2.0)
    at = S.snap_to_grid(at, n)
    h = ctx.alloc(n); ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=0.0)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    om = S.omega(x2c)

    def run():
        hl = ctx.fft_derivative(h, x2c, "lap")
        b = ctx.bader_assign(h, car2lat, lid)
        b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
        vol, ps = ctx.integrate(b, [h, hl], om)
        b.free(); ctx.free(hl)
        return vol, ps
    ms, prof, (vol, ps) = timed(run)
    f = ctx.download(h, n)
    chk = {"volume_sum_over_omega": float(vol.sum() / om), "population_sum_vs_grid_sum": float(ps[:, 0].sum() / (f.sum() * om / f.size)),
           "laplacian_sum_over_scale": float(abs(ps[:, 1].sum()) / (np.abs(f).sum() * om / f.size))}
    if N <= 128:  # full parity against the oracle at the quick size
        idg, nattr, _, _ = orc.bader_integrate(f, x2c, atoms=at)
        chk["oracle_basins"] = int(nattr)
    emit("BADER assign + INTEGRABLE (rho, FFT lap) incl. the FFT", "configs[1] urea-like 16 atoms, tetragonal", n, ms, 32.0 + 16.0, prof, chk)
    ctx.free(h)


# ---- config 3: NCIPLOT on a 512^3 grid field, benzene-dimer-like 24 atoms in a 30 bohr box ----
def nci():
    N = 128 if quick else 512
    n = (N, N, N)
    x2c = S.cell_x2c(30.0, 30.0, 30.0)
    ang = np.arange(6) * np.pi / 3
    ring = np.stack([2.64 * np.cos(ang), 2.64 * np.sin(ang), np.zeros(6)], 1)      # C6 ring, bohr
    hyd = np.stack([4.69 * np.cos(ang), 4.69 * np.sin(ang), np.zeros(6)], 1)
    mono = np.concatenate([ring, hyd])
    dimer = np.concatenate([mono + [15.0, 15.0, 11.7], mono + [15.0 + 3.0, 15.0, 11.7 + 6.6]])  # parallel displaced, 3.5 A
    at = S.snap_to_grid(dimer / 30.0, n)
    z = np.array(([6.0] * 6 + [1.0] * 6) * 2); al = np.array(([2.2] * 6 + [1.9] * 6) * 2)
    h = ctx.alloc(n); ctx.promolecular(h, x2c, at, z, al, nimg=0, rc=0.0)

    ms, prof, (hr, hg) = timed(lambda: ctx.nci_rdg_resident(h, x2c, n), cleanup=free_all)
    # sub-sampled parity at the full size: a coarse node-aligned lattice evaluated by the oracle
    f = ctx.download(h, n)
    cg = ctx.download(hg, (n[2], n[1], n[0])); cr = ctx.download(hr, (n[2], n[1], n[0]))
    st = N // 16
    sub = (16, 16, 16)
    xm = x2c / np.array(n, dtype=float)[None, :] * st
    cro, cgo = orc.nci_rdg(f, x2c, nstep=sub, x0=np.zeros(3), xmat=xm)
    sel_g = cg[::st, ::st, ::st]; sel_r = cr[::st, ::st, ::st]
    chk = {"subsampled_points": int(np.prod(sub)), "max_rel_rdg_err_vs_oracle": float(np.abs(sel_g - cgo).max() / np.abs(cgo).max()),
           "max_rel_abs_crho_err": float(np.abs(np.abs(sel_r) - np.abs(cro)).max() / np.abs(cro).max())}
    emit("NCIPLOT RDG, tricubic, node-aligned", "configs[2] benzene-dimer-like 24 atoms, 30 bohr box", n, ms, 24.0, prof, chk)
    ctx.free(hr); ctx.free(hg)

    # FOURIER mode: 4 derived grids (one forward + one backward transform each, |grad| three backward) + the loop
    ms_d, prof_d, hd = timed(lambda: [ctx.fft_derivative(h, x2c, w) for w in ("grad", "xx", "yy", "zz")], reps=2, cleanup=free_all)
    emit("FFT-derived grids for NCIPLOT FOURIER (|grad|, Hxx, Hyy, Hzz)", "configs[2]", n, ms_d, 4 * 16.0, prof_d,
         {"note": "16 B/pt minimal per output grid (SURVEY.md 8d); cuFFT D2Z/Z2D do the transforms"})
    for x in hd:
        ctx.free(x)
    ctx.free(h)


# ---- config 4: YT on a 512^3 periodic density with many basins ----
def yt():
    N = 128 if quick else 512
    side = 4 if quick else 8
    n = (N, N, N)
    x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
    at, z, al = S.jittered_lattice(side, 4)
    at = S.snap_to_grid(at, n)
    h = ctx.alloc(n); ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
    vec, area = S.wscell(x2c / np.array(n, dtype=float)[None, :])
    om = S.omega(x2c)
    state = {}

    def build():
        if "b" in state:
            state["b"].free()
        state["b"] = ctx.yt_build(h, vec, area)
        return state["b"]
    ms_b, prof_b, b = timed(build, reps=2)
    b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
    ms_i, prof_i, (vol, ps) = timed(lambda: ctx.integrate(b, [h], om))
    st = b.stats()
    phi = st[0] / float(np.prod(n))
    nhi = 3.0  # mean number of higher neighbours of an IAS point (SURVEY.md 8d probe); the library does not count it
    f = ctx.download(h, n)
    chk = {"volume_sum_over_omega": float(vol.sum() / om), "population_sum_vs_grid_sum": float(ps[:, 0].sum() / (f.sum() * om / f.size)),
           "phi_ias": phi, "nhi_assumed": nhi, "bfs_levels": int(st[1]), "kahn_levels": int(st[2]), "nvec": int(len(area)), "basins": int(b.nmax)}
    emit("YT build (yt_integrate)", "configs[3] 512 atoms cubic, tie-free", n, ms_b, 24.0 + 12.0 * nhi * phi, prof_b, chk)
    emit("YT integrate (adjoint sweep, Volume + rho)", "configs[3]", n, ms_i, 4.0 + 8.0 + 12.0 * nhi * phi + 16.0 * phi, prof_i, chk)
    # ISOSURFACE regions on the same YT levels (yt_isosurface): median contour value -> merged regions
    isov = float(np.quantile(f[::8, ::8, ::8], 0.5))
    ms_s, prof_s, (reg, nraw, nsurv) = timed(lambda: b.isosurface(isov), reps=2, cleanup=lambda o: o[0].free())
    lab = reg.labels(n)
    emit("ISOSURFACE regions (yt_isosurface) from the resident YT levels", "configs[3], contour at the median density", n, ms_s,
         8.0 + 4.0 + 4.0, prof_s,
         {"regions_before_merging": int(nraw), "surviving_regions": int(nsurv), "labels_positive_iff_above_contour":
          bool(np.array_equal(lab > 0, f >= isov)), "note": "16 B/pt = read rho + read YT label + write region"})
    reg.free()
    b.free(); ctx.free(h)


# ---- FFT derivative alone at the headline size ----
def fft_big():
    N = 256 if quick else 1024
    n = (N, N, N)
    x2c = S.cell_x2c(40.0, 40.0, 40.0)
    at, z, al = S.jittered_lattice(8, 5)
    h = ctx.alloc(n); ctx.promolecular(h, x2c, S.snap_to_grid(at, n), z, al, nimg=1, rc=8.0)
    ms, prof, hl = timed(lambda: [ctx.fft_derivative(h, x2c, "lap")], reps=2, cleanup=free_all)
    free_all(hl)
    emit("FFT Laplacian (grid3%fft, iff = lap)", "1024^3 headline grid", n, ms, 16.0, prof,
         {"note": "16 B/pt = read f + write lap; the half-spectrum D2Z + multiply + Z2D moves ~64 B/pt through cuFFT's passes"})
    ctx.free(h)


# ---- formatted-text reader: a CHGCAR-like E18.11 block ----
def text_reader():
    rep = 8 if quick else 512            # 64^3 values of text, repeated: 128^3 or 512^3 values
    rng = np.random.default_rng(9)
    base = np.abs(rng.standard_normal(64 ** 3)) * 10.0 ** rng.integers(-6, 4, 64 ** 3)
    block = ("\n".join(" " + " ".join("%.11E" % v for v in base[q:q + 5]) for q in range(0, base.size, 5)) + "\n").encode()
    text = block * rep
    N = round((64 ** 3 * rep) ** (1 / 3))
    n = (N, N, N)
    t0 = time.perf_counter()
    ref = np.array(block.split(), dtype=np.float64)      # numpy's strtod loop, one thread: the CPU side
    t_cpu = time.perf_counter() - t0
    state = {}

    def run():
        if "h" in state:
            ctx.free(state["h"])
        state["h"], state["used"], state["nhost"] = ctx.parse_text(text, n, 0, 1.0)
        return state["h"]
    ms, prof, h = timed(run, reps=2)
    out = ctx.download(h, n).ravel(order="F")
    same = bool(np.array_equal(out[:ref.size].view(np.uint64), ref.view(np.uint64)) and
                np.array_equal(out[-ref.size:].view(np.uint64), ref.view(np.uint64)))
    kern = sum(v for k, v in prof.items() if k != "text_h2d")
    emit("formatted-text grid reader (CHGCAR-like E18.11 block), host text in, resident grid out", "SURVEY 8(f)-1", n, ms,
         len(text) / float(np.prod(n)) + 8.0, prof,
         {"bit_exact_vs_numpy_strtod": same, "values_on_exact_slow_path": int(state["nhost"]), "text_bytes": len(text),
          "kernels_only_ms": round(kern, 3), "kernels_only_GBps": round((2 * len(text) + 8.0 * np.prod(n)) / (kern * 1e-3) / 1e9, 1),
          "cpu_numpy_values_per_s_1thread": ref.size / t_cpu, "gpu_values_per_s_incl_h2d": float(np.prod(n)) / (ms * 1e-3)})
    ctx.free(h)


# ---- formatted-text writer: NCIPLOT's cube body for a 512^3 RDG grid ((6(" ",1p,e13.5e3))) ----
def text_writer():
    N = 128 if quick else 512
    n = (N, N, N)
    x2c = S.cell_x2c(30.0, 30.0, 30.0)
    at, z, al = S.random_atoms(24, 3, x2c, dmin=2.0)
    h = ctx.alloc(n); ctx.promolecular(h, x2c, S.snap_to_grid(at, n), z, al, nimg=0, rc=0.0)
    hr, hg = ctx.nci_rdg_resident(h, x2c, n)
    import torch
    nbytes = ctx.format_text_size(hg, 0, 13, 5, 1)
    pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)       # the caller's buffer (page-locked)
    ms, prof, _ = timed(lambda: ctx.format_text_into(hg, 0, 13, 5, 1, pinned.data_ptr(), nbytes), reps=2)
    text = pinned.numpy()[:200000].tobytes()
    # parity on the first rows against the oracle (pure Python: a few thousand values)
    cg = ctx.download(hg, n)
    nrow = 8
    ref = orc.format_text_grid(cg[:, :nrow, :1], 0, 13, 5, 1)
    t0 = time.perf_counter()
    cpu = "".join(" %13.5E" % v for v in cg.ravel(order="F")[:2000000])     # C printf on one host thread, for scale
    t_cpu = time.perf_counter() - t0
    kern = prof.get("text_format", 0.0)
    emit("formatted-text writer (NCIPLOT cube body, 1p,e13.5e3), resident grid in, host text out", "SURVEY 8(f)-1", n, ms,
         8.0 + nbytes / float(np.prod(n)), prof,
         {"first_rows_identical_to_oracle": bool(text[:len(ref)] == ref), "text_bytes": int(nbytes), "kernel_only_ms": kern,
          "kernel_only_GBps": round((8.0 * np.prod(n) + nbytes) / (max(kern, 1e-9) * 1e-3) / 1e9, 1),
          "cpu_python_printf_values_per_s_1thread": 2000000 / t_cpu, "gpu_values_per_s_incl_d2h": float(np.prod(n)) / (ms * 1e-3)})
    for x in (h, hr, hg):
        ctx.free(x)


# ---- INTEGRABLE ... MULTIPOLES 5 (the reference default lmax) on the urea-like 256^3 Bader basins ----
def multipoles():
    N = 128 if quick else 256
    n = (N, N, N)
    x2c = S.cell_x2c(10.52, 10.52, 8.85)
    at, z, al = S.random_atoms(16, 2, x2c, dmin=2.0)
    at = S.snap_to_grid(at, n)
    h = ctx.alloc(n); ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=0.0)
    _, car2lat, lid = orc.bader_metrics(x2c, n)
    om = S.omega(x2c)
    b = ctx.bader_assign(h, car2lat, lid)
    pm = b.maxima()
    b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
    xattr = ((pm - 1) / np.array(n, dtype=float)).T
    for lmax in (5, 2):
        ms, prof, mp = timed(lambda: ctx.integrate_multipoles(b, h, lmax, xattr, x2c, om))
        _, ps = ctx.integrate(b, [h], om)
        chk = {"lmax": lmax, "moments_per_basin": (lmax + 1) ** 2, "basins": int(b.nmax),
               "monopole_vs_population_max_rel": float(np.abs(mp[0] - ps[:, 0]).max() / np.abs(ps[:, 0]).max()),
               "note": "FP64-pipe bound (acos, atan2, lmax+1 sincos, the genylm recursion per point), 12 algorithmic B/pt"}
        if lmax == 5:
            # sub-volume parity + CPU side: the oracle on the first 16 planes of the same labels and field
            f = ctx.download(h, n); idg = b.labels(n)
            sub_f = np.asfortranarray(f[:, :, :16]); sub_l = np.asfortranarray(idg[:, :, :16])
            t0 = time.perf_counter()
            # same point coordinates need the same n3: evaluate the oracle on the full grid with labels zeroed elsewhere
            lab0 = np.zeros_like(idg); lab0[:, :, :16] = sub_l
            ref = orc.multipoles_bader(lab0, xattr, lmax, f, orc.Cell(x2c), om)
            t_cpu = time.perf_counter() - t0
            hz = ctx.upload(np.asfortranarray(np.where(lab0 > 0, f, 0.0)))
            got = ctx.integrate_multipoles(b, hz, lmax, xattr, x2c, om)
            ctx.free(hz)
            scale = np.abs(ref[0]).max() * (0.5 * np.linalg.norm(x2c, axis=0).sum()) ** np.repeat(np.arange(lmax + 1), 2 * np.arange(lmax + 1) + 1)
            chk["subvolume_max_err_over_scale"] = float((np.abs(got - ref) / scale[:, None]).max())
            chk["cpu_oracle_points_per_s_1thread"] = float(sub_f.size / t_cpu)
        emit("INTEGRABLE MULTIPOLES (Bader basins, lmax = %d)" % lmax, "configs[1] urea-like 16 atoms, tetragonal", n, ms, 12.0, prof, chk)
    b.free(); ctx.free(h)


# ---- HIRSHFELD on the urea-like 256^3 grid: promolecular density + atomic integrals (synthetic Slater-type radial grids) ----
def hirshfeld():
    from test_oracle_hirshfeld import slater_tables
    N = 64 if quick else 256
    n = (N, N, N)
    x2c = S.cell_x2c(10.52, 10.52, 8.85)
    at, z, al = S.random_atoms(16, 2, x2c, dmin=2.0)
    ispc = np.arange(1, 17, dtype=np.int32)                     # one species per atom (its own Z, alpha)
    g = slater_tables(z, al, ngrid=600, a=1e-4, rmax=12.0)
    tab = dict(ngrid=g.ngrid, off=g.off, a=g.a, b=g.b, rmax=g.rmax, rcut=g.rcut, rtab=g.rtab, ftab=g.ftab)
    ms_p, prof_p, hp = timed(lambda: [ctx.promolecular_grid(n, x2c, at, ispc, tab)], reps=2, cleanup=free_all)
    hp = hp[0]
    om = S.omega(x2c)
    ms_h, prof_h, (vol, ps) = timed(lambda: ctx.hirshfeld_integrate(hp, x2c, at, ispc, tab, [hp], om), reps=2)
    chk = {"atoms": 16, "cutoff_bohr": 12.0, "volume_sum_over_omega": float(vol.sum() / om),
           "population_sum_vs_Z": float(ps[:, 0].sum() / np.sum(z)), "note": "FP64-pipe bound (log + 12 divisions per image and point)"}
    if quick:   # full parity against the oracle at the quick size
        ref = orc.promolecular_grid(n, x2c, at, ispc, g)
        chk["promolecular_max_rel_err_vs_oracle"] = float(np.abs(ctx.download(hp, n) - ref).max() / ref.max())
    emit("HIRSHFELD: promolecular density on the grid (promolecular_array3)", "configs[1] cell, 16 atoms", n, ms_p, 8.0, prof_p, chk)
    emit("HIRSHFELD: atomic integrals (intgrid_hirshfeld_fields loop, Volume + 1 field)", "configs[1] cell, 16 atoms", n, ms_h, 16.0, prof_h, chk)
    ctx.free(hp)


only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--only=")]
for fn in (urea, nci, yt, fft_big, text_reader, text_writer, multipoles, hirshfeld):
    if only and fn.__name__ not in only[0]:
        continue

    try:
        fn()
    except Exception as e:  # keep going: one JSON line per failure
        print(json.dumps({"path": fn.__name__, "error": repr(e)}), flush=True)
ctx.close()
