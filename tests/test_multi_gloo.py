"""World-size-2 CPU test (gloo) of the host-side multi-GPU logic: the z-slab decomposition of the C ABI
(c2g_slab_bounds_query), the sharded integration + all-reduce of bench.py, and its max-over-ranks timing.
The device kernels themselves are covered by the -m gpu tests (tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    import systems as S
    from critic2_b200 import capi
    from oracle import oracle as orc
    import bench

    c = cases.make_case("cubic48")
    n = c["n"]
    zlo, zhi = capi.slab_bounds(n[2], world, rank)
    term, _ = orc.bader_canonical(c["f"], c["x2c"])            # stands in for the device labels
    uniq = np.unique(term)
    lab = (np.searchsorted(uniq, term) + 1).astype(np.int32)
    # each rank integrates its own slab, partial sums are all-reduced (what c2g_integrate does with NCCL)
    sums = np.zeros(len(uniq))
    cnts = np.zeros(len(uniq))
    for i in range(len(uniq)):
        m = lab[:, :, zlo:zhi] == i + 1
        sums[i] = c["f"][:, :, zlo:zhi][m].sum()
        cnts[i] = m.sum()
    t = torch.from_numpy(np.concatenate([sums, cnts]))
    dist.all_reduce(t)
    tot = t.numpy()
    full = np.array([c["f"][lab == i + 1].sum() for i in range(len(uniq))])
    ok = np.allclose(tot[: len(uniq)], full, rtol=1e-12) and tot[len(uniq):].sum() == c["f"].size
    # bench.py's timing reduction: max over ranks
    tmax = bench.max_over_ranks(1.0 + rank)
    ok = ok and tmax == float(world)
    q.put((rank, zlo, zhi, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_integration_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 48
    assert all(r[3] for r in res)
