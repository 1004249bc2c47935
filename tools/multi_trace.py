"""Per-call wall time vs kernel time of the multi-GPU BADER step (run under torchrun; rank 0 prints)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
import bench
from critic2_b200 import capi
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
uid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = ctypes.create_string_buffer(128)
    if rank == 0: capi.load().c2g_nccl_unique_id(buf)
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda(); dist.broadcast(t, 0)
    uid = ctypes.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
ctx = capi.Context(local, rank=rank, nranks=world, nccl_uid=uid)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, x2c, at, z, al, side = bench.workload(size)
car2lat, lid = bench.bader_metrics(x2c, n)
h = ctx.alloc(n); ctx.promolecular(h, x2c, at, z, al, nimg=1, rc=8.0)
ctx.profile_enable(True)
for rep in range(4):
    ctx.profile_reset(); ctx.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    b = ctx.bader_assign(h, car2lat, lid)
    t1 = time.perf_counter()
    b.set_map(b.nmax, np.arange(1, b.nmax + 1, dtype=np.int32))
    t2 = time.perf_counter()
    vol, ps = ctx.integrate(b, [h, h], abs(np.linalg.det(x2c)))
    t3 = time.perf_counter()
    b.free(); ctx.synchronize()
    t4 = time.perf_counter()
    prof = ctx.profile()
    if rank == 0:
        print(f"rep {rep}: assign {1e3*(t1-t0):.2f} ms, set_map {1e3*(t2-t1):.2f}, integrate {1e3*(t3-t2):.2f}, free {1e3*(t4-t3):.2f}; kernels {sum(v[0] for v in prof.values()):.2f} ms", flush=True)
        if rep == 3: print({k: round(v[0], 3) for k, v in prof.items()})
ctx.close()
if world > 1: dist.destroy_process_group()
