#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native critic2 on-grid QTAIM hot path.

Metric (BASELINE.json): grid points/s for BADER assign + integrate.  One "step" = one pass of the hot
path over one synthetic grid: c2g_bader_assign (near-grid basin assignment) + c2g_integrate (Volume,
rho and a second INTEGRABLE grid; P_f = 2, 32 algorithmic bytes per grid point).

Workload (config.workload): BASELINE.json configs[4] -- a 1024^3 synthetic periodic promolecular-like
density, cubic 40 bohr cell, 512 atoms (jittered 8x8x8 lattice, seed 5), generated directly in HBM.  It
fits one B200 (rho 8.6 GB + second field 8.6 GB + labels 4.3 GB + work lists), so it is also the N=1
workload; at N>1 the same grid is sharded as z-slabs (strong scaling).

Arms:
  default            this repository's CUDA path through the C ABI (critic2_b200/libcritic2_gpu.so)
  --impl reference   critic2's own CPU algorithm.  critic2 is Fortran and cannot be compiled in this
                     image (no Fortran compiler), so this arm times the C++ restatement in oracle/
                     (kind "port") on a bounded sample of the same density model with all host threads
                     the reference would use (its Bader assignment is serial; the integration is
                     OpenMP over attractors).

JSON keys follow the driver's contract; see DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALG_BYTES_ASSIGN = 12.0      # read rho (8) + write label (4)            SURVEY.md 8(d)
ALG_BYTES_INTEGRATE = 20.0   # read label (4) + 2 fp64 integrand grids   SURVEY.md 8(d)
ALG_BYTES_TOTAL = ALG_BYTES_ASSIGN + ALG_BYTES_INTEGRATE


def max_over_ranks(x: float) -> float:
    """Max over ranks of a host float (identity when torch.distributed is not initialised)."""
    try:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
    except ImportError:
        pass
    return float(x)


def barrier():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
    except ImportError:
        pass


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML (nvidia_ml_py).
    An nvidia-smi subprocess polling in a loop was measured to stall kernel launches (it holds driver locks: a
    34 ms step took 124 ms under `nvidia-smi -lms 100`), so it is only the fallback, queried once after the region."""

    def __init__(self, index=0, period=0.05):
        self.index = index
        self.period = period
        self.rows = []   # (sm_mhz, sm_max_mhz, reasons bitmask)
        self._stop = threading.Event()
        self.thread = None
        self.nvml = None
        self.handle = None

    def start(self):
        if os.environ.get("C2G_BENCH_NOCLOCKS"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    idx = self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _loop(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((sm, self.smax, rs))
            except Exception:
                pass
            self._stop.wait(self.period)

    def mark(self):
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        if self.nvml is None:
            return self._smi_once()
        self._stop.set()
        self.thread.join(timeout=1.0)
        rows = self.rows[lo:hi] if hi is not None and hi > lo else self.rows[lo:]
        if not rows:
            rows = self.rows
        nv = self.nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(nm for nm, bit in names.items() if any(r[2] & bit for r in rows))
        sm = [r[0] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax if sm else None,
                "samples": len(sm), "reasons": reasons, "how": "NVML in-process, every 50 ms inside the timed region"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}",
                                  "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "samples": 1,
                    "reasons": sorted(nm for nm, v in zip(names, f[2:6]) if v.lower().startswith("active")),
                    "how": "nvidia-smi, once right after the timed region (NVML python binding unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


def workload(size: int):
    """Density model of configs[4] scaled to `size`^3: 5 bohr atom spacing, 128 grid points per atom
    spacing at 1024^3 (side = size/128 atoms per axis, at least 2)."""
    import systems as S
    side = max(2, size // 128)
    n = (size, size, size)
    x2c = S.cell_x2c(5.0 * side, 5.0 * side, 5.0 * side)
    at, z, al = S.jittered_lattice(side, 5)
    at = S.snap_to_grid(at, n)
    return n, x2c, at, z, al, side


def bader_metrics(x2c, n):
    lat2car = x2c / np.asarray(n, dtype=np.float64)[None, :]
    car2lat = np.linalg.inv(lat2car)
    lid = np.zeros((3, 3, 3))
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                if (i, j, k) != (0, 0, 0):
                    lid[i + 1, j + 1, k + 1] = 1.0 / np.linalg.norm(lat2car @ np.array([i, j, k], dtype=float))
    return car2lat, lid


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle): bounded sample of the same density model
# ------------------------------------------------------------------------------------------------
def cpu_sample_system(n):
    """Atoms of the density model (5 bohr spacing, 128 grid points per spacing, seed 5) for a grid of n points."""
    import systems as S
    sides = [max(1, x // 128) for x in n]
    x2c = S.cell_x2c(5.0 * sides[0], 5.0 * sides[1], 5.0 * sides[2])
    if n[0] == n[1] == n[2]:   # a full config: exactly bench.workload(size)
        _, x2c, at, z, al, _ = workload(n[0])
        return x2c, at, z, al
    rng = np.random.default_rng(5)
    g = np.stack(np.meshgrid(*[np.arange(s) for s in sides], indexing="ij"), -1).reshape(-1, 3).astype(float)
    at = (g + 0.5) / np.array(sides)[None, :] + rng.uniform(-0.15, 0.15, g.shape) / np.array(sides)[None, :]
    at = S.snap_to_grid(at % 1.0, n)
    z = rng.uniform(1.0, 8.0, len(at)); al = rng.uniform(1.2, 2.7, len(at))
    return x2c, at, z, al


def cpu_baseline_run(sample_n=(256, 256, 128), steps=1, budget_s=None):
    """Times the oracle's faithful bader_integrate (serial, like the reference) + integrate_bader (OpenMP over
    attractors, like the reference) on a grid of the same density model.  Runs `steps` passes, fewer if the next one
    would exceed `budget_s` seconds; returns the MEAN points/s over the passes actually run."""
    import systems as S
    from oracle import oracle as orc
    n = tuple(int(x) for x in sample_n)
    x2c, at, z, al = cpu_sample_system(n)
    f = orc.promolecular(n, x2c, at, z, al, nimg=1, rc=8.0)
    f2 = np.asfortranarray(np.roll(f, 3, 0) * 0.5)
    times = []
    t_begin = time.perf_counter()
    for _ in range(max(1, steps)):
        if times and budget_s is not None and (time.perf_counter() - t_begin) + max(t[0] for t in times) > budget_s:
            break
        t0 = time.perf_counter()
        idg, nattr, _, stats = orc.bader_integrate(f, x2c, atoms=at)
        t1 = time.perf_counter()
        orc.integrate_bader(idg, [f, f2], nattr, S.omega(x2c))
        t2 = time.perf_counter()
        times.append((t2 - t0, t1 - t0, t2 - t1))
    npts = float(np.prod(n))
    mean = [float(np.mean([t[i] for t in times])) for i in range(3)]
    return {
        "value": npts / mean[0], "unit": "grid points/s", "cores": int(orc.num_threads()), "kind": "port",
        "sample": f"{n[0]}x{n[1]}x{n[2]} grid of the same density model ({len(at)} atoms, 128 points per atom spacing), "
                  f"mean of {len(times)} pass(es); oracle bader_integrate (serial like bader@proc.f90) {mean[1]:.1f} s + "
                  f"integrate (OpenMP over attractors) {mean[2]:.2f} s per pass",
        "seconds": mean[0], "passes": len(times), "grid": list(n),
    }


def config_dict(size, side, world, nmax=None):
    """The `config` object of both arms (identical for the same --size and --gpus)."""
    return {"workload": f"BADER assign + INTEGRABLE (Volume, rho, second grid; P_f=2) on a synthetic {size}^3 "
                        f"promolecular-like periodic density, cubic {5.0 * side:.0f} bohr cell, {side**3} atoms "
                        "(BASELINE.json configs[4])",
            "grid": [size, size, size], "atoms": side ** 3, "maxima": side ** 3 if nmax is None else nmax,
            "parallelism": f"z-slabs x{world}" if world > 1 else "single GPU",
            "l2": "inputs (>= 1 GB per field) are larger than the 126 MB L2; no flush needed",
            "algo": "hierarchical exact near-grid walks + edge refinement (C2G_BADER_FAST)"}


def run_reference(args):
    """critic2's own CPU algorithm for the path (the oracle port: critic2 is Fortran and no Fortran compiler exists in
    this image).  --size <= 512: every step is one pass over the FULL grid of our arm's config (same_config).  The
    1024^3 default would take ~20 min per pass: every step is then one FULL pass over the 512^3 grid of the same density
    model (the other size the metric names; the grid of our arm's `also_512` line), and the JSON line says so.  The
    run stops early when the next pass would take the whole run beyond ~150 s; `steps` is the number of passes timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t_all = time.perf_counter()
    size = args.size
    side = max(2, size // 128)
    full = size <= 512
    # 1024^3 would take ~20 min per pass on one host core: the reference arm then times the FULL 512^3 grid of the same
    # density model -- the other size BASELINE.json's metric names, and exactly the grid of our arm's `also_512` line
    # (C2G_REF_SAMPLE=1: the bounded 256x256x128 sample instead, ~10 s per pass)
    if full:
        sample_n = (size, size, size)
    elif os.environ.get("C2G_REF_SAMPLE"):
        sample_n = (256, 256, 128)
    else:
        sample_n = (512, 512, 512)
    res = cpu_baseline_run(sample_n=sample_n, steps=max(1, args.steps), budget_s=float(os.environ.get("C2G_REF_BUDGET_S", "150")))
    out = {
        "impl": "reference", "metric": "grid points/s, BADER assign+integrate", "value": res["value"], "unit": "grid points/s",
        "n_gpus": args.gpus, "steps": res["passes"], "steps_requested": args.steps, "warmup": 0,
        "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(size, side, max(1, args.gpus)),
        "same_config_as_gpu_arm": bool(full),
        "same_grid_as_gpu_arm_also_512": bool(list(res["grid"]) == [512, 512, 512]),
        "timed_grid": res["grid"],
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "grid points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("critic2 (Fortran) cannot be compiled in this image; this is the C++ restatement in oracle/ of the reference's own "
                 "CPU algorithm. " + ("Timed on the full grid of the config." if full else
                 "ms_per_step and value are those of ONE FULL pass over the grid named in timed_grid / cpu_baseline.sample (the 512^3 "
                 "workload of the same density model = the grid of the GPU arm's also_512 line), NOT of a full "
                 f"{size}^3 pass, which takes ~20 min on one host core (measured once: tests/golden/bader_at_size.json, head1024, "
                 "oracle_seconds); points/s of this serial algorithm is size-independent to ~15 % (0.81e6 at 1024^3, 0.92e6 at 512^3).")),
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(out), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    from critic2_b200 import capi

    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            rc = capi.load().c2g_nccl_unique_id(buf)
            if rc != 0:
                raise SystemExit("c2g_nccl_unique_id failed")
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        uid = ctypes.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
    ctx = capi.Context(local, rank=rank, nranks=world, nccl_uid=uid)

    size = args.size
    n, x2c, at, z, al, side = workload(size)
    nn = float(np.prod(n))
    car2lat, lid = bader_metrics(x2c, n)
    omega = abs(np.linalg.det(x2c))

    # inputs resident in HBM (generated on the device; every rank holds the replicated read-only field)
    h_rho = ctx.alloc(n)
    ctx.promolecular(h_rho, x2c, at, z, al, nimg=1, rc=8.0)
    h_f2 = ctx.alloc(n)
    ctx.promolecular(h_f2, x2c, at, z * 0.5, al * 1.3, nimg=1, rc=8.0)

    ident = None

    def step_resident():
        nonlocal ident
        b = ctx.bader_assign(h_rho, car2lat, lid, algo=capi.BADER_FAST)
        if ident is None or len(ident) != b.nmax:
            ident = np.arange(1, b.nmax + 1, dtype=np.int32)
        b.set_map(b.nmax, ident)
        vol, ps = ctx.integrate(b, [h_rho, h_f2], omega)
        return b, vol, ps

    # ---- warm-up (the clock sampler starts here: its start-up must not perturb the timed region) ----
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        b, vol, ps = step_resident()
        nmax = b.nmax
        b.free()
    pop_sum = float(ps[:, 0].sum())

    # ---- timed region: device-resident inputs ----
    ctx.profile_enable(not os.environ.get("C2G_BENCH_NOPROF"))
    ctx.profile_reset()
    l0 = ctx.launch_count()
    barrier(); ctx.synchronize()
    m0 = clocks.mark()
    ctx.timer_start()
    trace = []
    for _ in range(args.steps):
        tw = time.perf_counter()
        b, vol, ps = step_resident()
        b.free()
        trace.append(round((time.perf_counter() - tw) * 1e3, 2))
    ms = ctx.timer_stop()
    if os.environ.get("C2G_BENCH_TRACE") and rank == 0:
        print("host wall per step (ms):", trace, file=sys.stderr, flush=True)
    m1 = clocks.mark()
    barrier()
    launches = ctx.launch_count() - l0
    clk = clocks.stop(m0, m1) if rank == 0 else None
    ms_step = max_over_ranks(ms / args.steps)
    prof = ctx.profile()
    ctx.profile_enable(False)
    value = nn / (ms_step * 1e-3)

    # ---- roofline (measured live: CUDA events around every kernel of the timed region) ----
    peak, peak_src = measured_peaks()
    assign_ms = sum(v[0] for k, v in prof.items() if k.startswith("bader_")) / args.steps
    integ_ms = sum(v[0] for k, v in prof.items() if k.startswith("basin_")) / args.steps
    walk_ms = sum(v[0] for k, v in prof.items() if k.startswith("bader_walk")) / args.steps
    zlo, zhi = ctx.slab_range(n[2])
    nloc = float(n[0] * n[1] * (zhi - zlo))
    kern_ms = max_over_ranks(assign_ms + integ_ms)
    achieved = ALG_BYTES_TOTAL * nloc / (max(assign_ms + integ_ms, 1e-9) * 1e-3) / 1e9
    # measured DRAM traffic of the same kernel group (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per
    # step), from the committed capture of this workload; null when there is none for this size / GPU count
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"{size}^3x{world}"
        if key in tj:
            traffic, traffic_src = tj[key]["dram_bytes_per_step"], tj[key]["source"]
    except (OSError, ValueError, KeyError):
        pass
    kms = {k: v[0] / args.steps for k, v in prof.items()}
    dk = max(kms, key=kms.get) if kms else None
    dominant = None if dk is None else {
        "name": dk + (" (k_walk3, last-level launch)" if dk == "bader_walk_l1" else ""), "ms": round(kms[dk], 4),
        "share_of_kernel_group": round(kms[dk] / max(assign_ms + integ_ms, 1e-9), 4),
        "note": "the walkers are issue-bound, not HBM-bound (ncu, profiles/r02m_ncu_step_1024_table.txt: 67 % of the issue slots busy, "
                "24 of 32 lanes active, 227 warp instructions per step, L1 / L2 hit 75 %, 5.7 GB of DRAM traffic for 7.4e8 steps): an HBM "
                "fraction of this launch alone would be meaningless, so `achieved` is the algorithmic bytes of the whole step over the "
                "time of the whole kernel group"}
    roofline = {
        "bound": "hbm", "kernel": "BADER assign+integrate kernel group (k_maxima2, k_walk3, k_walk_big2, k_classify, k_vsafe, k_pack5, k_items_*, "
                                  "k_fill_edge_v, k_requeue, k_basin_reduce); dominant: k_walk3 (last-level launch, bader_walk_l1)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "algorithmic_bytes_per_point": ALG_BYTES_TOTAL, "points_per_launch_group": nloc, "traffic": traffic,
        "traffic_source": traffic_src,
        "stages": {
            "assign": {"ms": assign_ms, "GBps": ALG_BYTES_ASSIGN * nloc / max(assign_ms, 1e-9) / 1e6, "of_which_walk_ms": walk_ms},
            "integrate": {"ms": integ_ms, "GBps": ALG_BYTES_INTEGRATE * nloc / max(integ_ms, 1e-9) / 1e6},
        },
        "kernels_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in prof.items()},
        "dominant_kernel": dominant,
        "kernel_group_ms_max_over_ranks": kern_ms,
    }

    # ---- e2e: through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region ----
    plane = n[0] * n[1]
    nzl = zhi - zlo
    host_rho = torch.empty(max(1, plane * nzl), dtype=torch.float64, pin_memory=True)
    host_f2 = torch.empty(max(1, plane * nzl), dtype=torch.float64, pin_memory=True)
    host_idg = torch.empty(max(1, plane * nzl), dtype=torch.int32, pin_memory=True)
    ctx.download_slab_ptr(h_rho, host_rho.data_ptr())
    ctx.download_slab_ptr(h_f2, host_f2.data_ptr())

    def step_e2e():
        """Host buffers in, host results out, through the public C ABI.  Single GPU: the asynchronous entry points
        let the upload of the second field overlap the assignment on the first, and the download of the labels
        overlap that upload (PCIe is full duplex); everything is complete at the final synchronize."""
        nonlocal ident
        if world > 1:
            ha = ctx.upload_slab_ptr(host_rho.data_ptr(), n)
            hb = ctx.upload_slab_ptr(host_f2.data_ptr(), n)
        else:
            ha = ctx.upload_ptr_async(host_rho.data_ptr(), n)
            hb = ctx.upload_ptr_async(host_f2.data_ptr(), n)
        bb = ctx.bader_assign(ha, car2lat, lid, algo=capi.BADER_FAST)
        if ident is None or len(ident) != bb.nmax:
            ident = np.arange(1, bb.nmax + 1, dtype=np.int32)
        bb.set_map(bb.nmax, ident)
        if world > 1:
            v, p = ctx.integrate(bb, [ha, hb], omega)
            bb.labels_ptr(host_idg.data_ptr())
        else:
            bb.labels_ptr_async(host_idg.data_ptr())
            v, p = ctx.integrate(bb, [ha, hb], omega)
            ctx.synchronize()
        bb.free(); ctx.free(ha); ctx.free(hb)
        return p

    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()  # warm-up
    barrier(); ctx.synchronize()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pe = step_e2e()
    ms_e = ctx.timer_stop()
    wall_e = (time.perf_counter() - t0) * 1e3
    barrier()
    ms_e2e = max_over_ranks(max(ms_e, wall_e) / e2e_steps)
    e2e = {"value": nn / (ms_e2e * 1e-3), "unit": "grid points/s", "ms_per_step": ms_e2e, "steps": e2e_steps,
           "h2d_bytes_per_step": int(2 * 8 * nn), "d2h_bytes_per_step": int(4 * nn + 8 * 3 * nmax),
           "pop_sum_matches_resident": bool(abs(float(pe[:, 0].sum()) - pop_sum) <= 1e-9 * abs(pop_sum))}

    # ---- the other size BASELINE.json's metric names (512^3), device-resident, same step, a few passes ----
    also = None
    if world == 1 and size != 512 and not os.environ.get("C2G_BENCH_NO512"):
        n5, x5, at5, z5, al5, side5 = workload(512)
        c5, l5 = bader_metrics(x5, n5)
        om5 = abs(np.linalg.det(x5))
        g1 = ctx.alloc(n5); ctx.promolecular(g1, x5, at5, z5, al5, nimg=1, rc=8.0)
        g2 = ctx.alloc(n5); ctx.promolecular(g2, x5, at5, z5 * 0.5, al5 * 1.3, nimg=1, rc=8.0)
        id5 = np.arange(1, side5 ** 3 + 1, dtype=np.int32)

        def step512():
            b5 = ctx.bader_assign(g1, c5, l5, algo=capi.BADER_FAST)
            b5.set_map(b5.nmax, id5[: b5.nmax])
            r = ctx.integrate(b5, [g1, g2], om5)
            b5.free()
            return r
        for _ in range(3):
            step512()
        ctx.synchronize(); ctx.timer_start()
        k5 = max(3, min(args.steps, 10))
        for _ in range(k5):
            v5, p5 = step512()
        ms5 = ctx.timer_stop() / k5
        also = {"grid": [512, 512, 512], "atoms": side5 ** 3, "steps": k5, "ms_per_step": ms5, "value": float(np.prod(n5)) / (ms5 * 1e-3),
                "unit": "grid points/s", "roofline_frac_of_step": ALG_BYTES_TOTAL * float(np.prod(n5)) / (ms5 * 1e-3) / 1e9 / peak,
                "volume_sum_over_omega": float(v5.sum() / om5)}
        # the same grid end to end (pinned host buffers in, labels + sums out): the line the reference arm's default run
        # (one full 512^3 pass on the host) is the same-grid counterpart of
        nn5 = int(np.prod(n5))
        hr5 = torch.empty(nn5, dtype=torch.float64, pin_memory=True)
        hf5 = torch.empty(nn5, dtype=torch.float64, pin_memory=True)
        hi5 = torch.empty(nn5, dtype=torch.int32, pin_memory=True)
        ctx.download_slab_ptr(g1, hr5.data_ptr()); ctx.download_slab_ptr(g2, hf5.data_ptr())

        def e2e512():
            ha = ctx.upload_ptr_async(hr5.data_ptr(), n5)
            hb = ctx.upload_ptr_async(hf5.data_ptr(), n5)
            b5 = ctx.bader_assign(ha, c5, l5, algo=capi.BADER_FAST)
            b5.set_map(b5.nmax, id5[: b5.nmax])
            b5.labels_ptr_async(hi5.data_ptr())
            r = ctx.integrate(b5, [ha, hb], om5)
            ctx.synchronize()
            b5.free(); ctx.free(ha); ctx.free(hb)
            return r
        e2e512()
        ctx.synchronize(); t5 = time.perf_counter()
        for _ in range(3):
            e2e512()
        ms5e = (time.perf_counter() - t5) * 1e3 / 3
        also["e2e"] = {"value": float(nn5) / (ms5e * 1e-3), "unit": "grid points/s", "ms_per_step": ms5e, "steps": 3,
                       "h2d_bytes_per_step": 2 * 8 * nn5, "d2h_bytes_per_step": 4 * nn5 + 8 * 3 * side5 ** 3}
        del hr5, hf5, hi5
        ctx.free(g1); ctx.free(g2)

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline_run(steps=1)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the oracle is optional test infrastructure
            cpu = {"value": None, "unit": "grid points/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}

    if rank == 0:
        out = {
            "metric": "grid points/s, BADER assign+integrate", "value": value, "unit": "grid points/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(size, side, world, nmax),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "also_512": also, "gpu_launches": int(launches), "clocks": clk,
            "device": ctx.describe(), "check": {"population_sum": pop_sum, "volume_sum_over_omega": float(vol.sum() / omega)},
        }
        print(json.dumps(out), flush=True)
    ctx.free(h_rho); ctx.free(h_f2)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def run_other_config(args):
    """The BASELINE.json configs that are not the headline, as driver-visible bench lines: `--config urea256 | nci512 |
    yt512 | fft1024 | hirshfeld | multipoles` runs the matching path of tools/bench_paths.py (same library, same C ABI,
    inputs resident in HBM, CUDA-event timing, a parity check per path) and prints one contract-shaped JSON line per
    path it measures."""
    import subprocess
    fn = {"urea256": "urea", "nci512": "nci", "yt512": "yt", "fft1024": "fft_big", "hirshfeld": "hirshfeld",
          "multipoles": "multipoles"}[args.config]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_paths.py"), f"--only={fn}"], capture_output=True, text=True)
    n = 0
    for line in out.stdout.splitlines():
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if "error" in d:
            print(json.dumps({"metric": "grid points/s", "config": {"workload": args.config}, "error": d["error"]}), flush=True)
            continue
        n += 1
        print(json.dumps({
            "metric": "grid points/s, " + d["path"], "value": d["points_per_s"], "unit": "grid points/s", "n_gpus": 1, "steps": 3,
            "warmup": 1, "ms_per_step": d["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": d["path"] + " -- " + d["config"], "grid": d["grid"],
                                            "l2": "inputs larger than the 126 MB L2 (>= 1 GB per field at 512^3)"},
            "roofline": {"bound": "hbm", "achieved": d["achieved_GBps"], "peak": d["peak_GBps"], "unit": "GB/s", "frac": d["frac"],
                         "peak_source": d["peak_source"], "algorithmic_bytes_per_point": d["algorithmic_bytes_per_point"],
                         "traffic": None, "kernels_ms_per_step": d["kernels_ms"]},
            "check": d["check"], "gpu_launches": None}), flush=True)
    if n == 0:
        sys.stderr.write(out.stderr[-2000:])
        return 1
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="grid points per axis (default: the 1024^3 headline config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--config", default="bader1024", choices=["bader1024", "urea256", "nci512", "yt512", "fft1024", "hirshfeld", "multipoles"],
                    help="bader1024 = the headline (BASELINE.json configs[4], default); the others run one path of tools/bench_paths.py")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "bader1024":
        return run_other_config(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
