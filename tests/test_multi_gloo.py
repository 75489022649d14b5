"""N>1 host logic on CPU: world_size-2 gloo run of the Morton-slice scheme with the oracle standing in
for the GPU engine.  Checks the slice arithmetic, the in-place all-gather layout and that the sliced
run reproduces the single-rank run bit for bit (slices are whole 32-body chunks)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, steps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import nbody_barnes_hut_cuda_b200 as bh
    import oracle_lib as O
    from nbody_barnes_hut_cuda_b200.sliced import allgather_slices, slice_bounds

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    soa = bh.ic_uniform_cube(n, 7, 1000.0)
    posm, vel, ids = O.soa_to_internal(soa)
    first, count, per = slice_bounds(n, rank, world)
    total = per * world
    full = [np.zeros((total, 4), np.float32), np.zeros((total, 4), np.float32), np.zeros(total, np.int32)]
    full[0][:n], full[1][:n], full[2][:n] = posm, vel, ids
    tens = [torch.from_numpy(a) for a in full]
    for _ in range(steps):
        r = O.engine_step(full[0][:n], full[1][:n], full[2][:n], 1, slice_first=first, slice_count=count)
        full[0][:n], full[1][:n], full[2][:n] = r["posm"], r["vel"], r["ids"]
        allgather_slices(dist, tens, rank, per)
    if rank == 0:
        q.put((full[0][:n].copy(), full[1][:n].copy(), full[2][:n].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_slice_bounds_cover_everything_in_whole_chunks():
    sys.path.insert(0, ROOT)
    from nbody_barnes_hut_cuda_b200.sliced import GROUP, slice_bounds

    for n in (1, 31, 32, 33, 1000, 16384, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                first, count, per = slice_bounds(n, r, world)
                assert (first % GROUP == 0 or count == 0) and per % GROUP == 0 and first == min(n, r * per)
                covered += count
            assert covered == n and per * world >= n


@pytest.mark.parametrize("n,steps", [(3000, 3), (4099, 2)])
def test_two_rank_sliced_run_equals_single_rank(n, steps):
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nbody_barnes_hut_cuda_b200 as bh
    import oracle_lib as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    soa = bh.ic_uniform_cube(n, 7, 1000.0)
    posm, vel, ids = O.soa_to_internal(soa)
    want = O.engine_step(posm, vel, ids, steps)
    assert (got[2] == want["ids"]).all()
    assert got[0].tobytes() == want["posm"].tobytes() and got[1].tobytes() == want["vel"].tobytes()
