"""Multi-GPU Morton-range slices (BASELINE.json north_star, SURVEY §8e "replicated tree").

One process per GPU (torchrun).  Every rank holds the full state; each step every rank
  1. runs the step graph: bounds, keys, sort, reorder and tree on ALL bodies (deterministic, so the
     trees are identical without a broadcast), traversal + kick-drift only on its own slice of the
     freshly sorted order, written into the current-state arrays;
  2. all-gathers the updated slices (posm 16 B, vel 16 B, ids 4 B per body) in place over
     NCCL/NVLink so every rank again holds the full state.
Slices are whole 32-body chunks, so chunk boundaries — and therefore every acceptance decision and
every float — are identical to the single-GPU run.

The step loop itself — slice bookkeeping, the in-place NCCL all-gathers on their own stream, the overlap of the
next step's keys + sort with the gathers of velocities and ids — is C++ behind the C ABI (bh_mg_*, csrc/bh_mg.cu).
This module is the test / bench harness around it: torch.distributed only carries the 128-byte NCCL id to the
other ranks and reduces the timings; `step(overlap=False)` keeps a torch all-gather loop as an independent
cross-check of the C++ driver.
"""
from __future__ import annotations

import time

import numpy as np

from .engine import lib as _lib

GROUP = int(_lib().bh_group_size())   # bodies per traversal chunk (BH_GROUP in csrc/bh_common.cuh)


# kernels of one sliced step of one rank: the three-part step (2 keys + 6 sort, reorder of positions + 5 build + 3 centre of
# mass + 2 force, reorder of the rest + update = 21) + the cube re-reduction over the gathered positions (2)
SLICED_LAUNCHES_PER_STEP = 23

def slice_bounds(n: int, rank: int, world: int):
    """Mirror of default_slice() in csrc/bh_engine.cu: (first, count, padded_per_rank)."""
    groups = (n + GROUP - 1) // GROUP
    per = (groups + world - 1) // world * GROUP
    first = min(n, rank * per)
    last = min(n, (rank + 1) * per)
    return first, last - first, per


def C_void(t):
    import ctypes as C

    return C.c_void_p(t.data_ptr())


class _DevView:
    """Zero-copy torch view of a raw device pointer via the CUDA array interface."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_views(torch, state: dict, per: int, world: int, device):
    total = per * world
    posm = torch.as_tensor(_DevView(state["posm"], (total, 4), "<f4"), device=device)
    vel = torch.as_tensor(_DevView(state["vel"], (total, 4), "<f4"), device=device)
    ids = torch.as_tensor(_DevView(state["ids"], (total,), "<i4"), device=device)
    return posm, vel, ids


def allgather_slices(dist, tensors, rank: int, per: int):
    """In-place all-gather: rank r's rows [r*per, (r+1)*per) of each full tensor."""
    for full in tensors:
        dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per])


class SlicedSimulation:
    def __init__(self, bh, soa, rank: int, world: int, local: int, dist, flags: int = 0, **params):
        import torch

        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.n = len(soa[0])
        self.first, self.count, self.per = slice_bounds(self.n, rank, world)
        self.device = torch.device(f"cuda:{local}")
        self.eng = bh.BHEngine(self.n, device=local, flags=flags, **params)
        self.eng.set_slice(rank, world)
        self.eng.load_soa(*soa)
        st = self.eng.state_ptrs()
        assert (st["first"], st["count"]) == (self.first, self.count), "slice arithmetic differs from the C side"
        self.views = device_views(torch, st, self.per, world, self.device)
        # the C++/NCCL driver: rank 0 makes the id, torch.distributed only carries it to the others
        idt = torch.zeros(128, dtype=torch.uint8, device=self.device)   # BH_MG_ID_BYTES
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(bh.mg_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        self.mg = bh.MultiGpu(self.eng, bytes(idt.cpu().numpy().tobytes()), rank, world, local)
        info = self.mg.info()
        assert (info["first"], info["count"], info["per"]) == (self.first, self.count, self.per)

    def step(self, nsteps: int = 1, overlap: bool = True):
        """nsteps sliced steps.  With overlap the three all-gathers are issued asynchronously in the order
        posm, vel, ids; the next step's head (bounds, keys, radix sort — positions only) starts as soon as
        the positions have landed, while velocities and ids are still crossing NVLink."""
        stream = self.torch.cuda.current_stream().cuda_stream
        if not overlap:
            for _ in range(nsteps):
                self.eng.simulation_step(1, stream)
                allgather_slices(self.dist, self.views, self.rank, self.per)
            return
        self.mg.step(nsteps, stream)   # bh_mg_step: asynchronous, the gathers of the last step stay in flight

    def finish(self):
        """The current stream waits for the all-gathers a previous step(overlap=True) left in flight."""
        self.mg.finish(self.torch.cuda.current_stream().cuda_stream)

    def step_host(self, host_in, host_out, nsteps: int = 1):
        """End-to-end step with HOST state, sharded over the ranks' PCIe links: every rank uploads 1/world of
        the 7 SoA arrays, an all-gather completes them on every device, bh_import_soa + sliced step(s) +
        slice all-gather run as usual, and every rank reads back 1/world of the 6 result arrays (original body
        order) into host_out — the host ends up with the full state, spread over the ranks.
        host_in: 7 pinned float32 torch tensors [n]; host_out: 6 pinned float32 torch tensors [n]."""
        torch, dist = self.torch, self.dist
        n, world, rank = self.n, self.world, self.rank
        chunk = (n + world - 1) // world
        lo, hi = min(n, rank * chunk), min(n, (rank + 1) * chunk)
        if not hasattr(self, "_soa_dev"):
            self._soa_dev = [torch.empty(chunk * world, dtype=torch.float32, device=self.device) for _ in range(9)]
        stream = torch.cuda.current_stream().cuda_stream
        for k in range(7):
            self._soa_dev[k][lo:hi].copy_(host_in[k][lo:hi], non_blocking=True)
            dist.all_gather_into_tensor(self._soa_dev[k], self._soa_dev[k][rank * chunk:(rank + 1) * chunk])
        self.finish()
        self.eng.load_soa_device(self._soa_dev[:7], n, stream)
        self.step(nsteps)
        self.finish()
        ptrs = [C_void(t) for t in self._soa_dev[:6]] + [None] * 3
        from .engine import _check, lib
        import ctypes as C
        _check(lib().bh_export_soa(self.eng._ctx, *ptrs, C.c_void_p(stream)), "bh_export_soa")
        for k in range(6):
            host_out[k][lo:hi].copy_(self._soa_dev[k][lo:hi], non_blocking=True)
        torch.cuda.synchronize()

    def close(self):
        self.finish()
        self.torch.cuda.synchronize()
        self.mg.close()
        self.eng.close()


def run_sliced_bench(args, w, bh, dist, rank, world, local):
    import torch

    import bench

    soa = bench.make_ic(bh, w)
    n = w["n"]
    sim = SlicedSimulation(bh, soa, rank, world, local, dist)
    dev = sim.device

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    sim.step(args.warmup)
    sim.finish()
    barrier()
    sim.eng.check_device_error()
    clocks = bench.ClockSampler(local)
    clocks.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    sim.step(args.steps)
    sim.finish()
    b.record()
    barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ck = clocks.stop()
    inter = torch.tensor([float(sim.eng.stat(bh.STAT.INTERACTIONS_CELL) + sim.eng.stat(bh.STAT.INTERACTIONS_BODY))],
                         device=dev, dtype=torch.float64)
    dist.all_reduce(inter)

    # per-phase split on this rank (phase-timer flag switched on for a few steps of the same context)
    # — shows the replicated (Amdahl) part
    stream = torch.cuda.current_stream().cuda_stream
    sim.eng.set_flags(2)
    psteps = 5
    acc = {}
    for _ in range(psteps):
        sim.eng.simulation_step(1, stream)
        for k, v in sim.eng.phase_ms().items():
            acc[k] = acc.get(k, 0.0) + v / psteps
        allgather_slices(dist, sim.views, rank, sim.per)
    sim.eng.set_flags(0)
    force_all = torch.zeros(world, device=dev, dtype=torch.float32)
    dist.all_gather_into_tensor(force_all, torch.tensor([acc["force"]], device=dev, dtype=torch.float32))
    force_per_rank = [round(float(x), 3) for x in force_all.tolist()]
    ag0, ag1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ag0.record()
    for _ in range(5):
        allgather_slices(dist, sim.views, rank, sim.per)
    ag1.record()
    barrier()
    allgather_ms = ag0.elapsed_time(ag1) / 5

    # e2e: host SoA in and out every step, sharded over the ranks' PCIe links (SlicedSimulation.step_host).
    # Skipped above 64M bodies (13 pinned host arrays per rank).
    e2e = None
    if n <= 64_000_000:
        host_in = [torch.from_numpy(x).pin_memory() for x in soa]
        host_out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(6)]
        esteps = 5
        sim.step_host(host_in, host_out, 1)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            sim.step_host(host_in, host_out, 1)
        barrier()
        e2e_s = torch.tensor([(time.perf_counter() - t0) / esteps], device=dev, dtype=torch.float64)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e = {"value": n / float(e2e_s.item()), "unit": "body-steps/s", "h2d_bytes_per_step": 28 * n,
               "d2h_bytes_per_step": 24 * n, "ms_per_step": float(e2e_s.item()) * 1e3,
               "api": "SlicedSimulation.step_host: each rank uploads 1/N of the host SoA, all-gather, bh_import_soa, "
                      "1 sliced step, slice all-gather, bh_export_soa, each rank reads back 1/N"}
    cells = sim.eng.stat(bh.STAT.CELLS)
    sim.close()
    total_ms = float(ms.item())
    # the SAME workload on ONE GPU (rank 0, 10 graph-replayed steps): the strong-scaling curve reads off this line alone
    same1 = None
    if rank == 0:
        with bh.BHEngine(n, device=local) as one:
            one.load_soa(*soa)
            stream1 = torch.cuda.current_stream().cuda_stream
            one.simulation_step(3, stream1)
            torch.cuda.synchronize()
            a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a1.record()
            one.simulation_step(10, stream1)
            b1.record()
            torch.cuda.synchronize()
            one.check_device_error()
            ms1 = a1.elapsed_time(b1) / 10
            same1 = {"workload": args.workload, "n_bodies": n, "ms_per_step": ms1, "value": n / (ms1 * 1e-3), "unit": "body-steps/s",
                     "steps": 10, "warmup": 3, "note": "bh_step on rank 0's GPU alone, same input, measured in this run"}
    dist.barrier()
    roofline = bench.force_roofline(bh, local, float(inter.item()) / world, max(force_per_rank))
    line = {
        "metric": "body-steps/s", "value": n * args.steps / (total_ms * 1e-3), "unit": "body-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench.workload_config(args, w),
        "engine": {"group": 32, "key_bits": 30, "launches_per_step": SLICED_LAUNCHES_PER_STEP,
                   "parallelism": f"morton-slices x{world}: replicated sort+tree, sliced traversal, in-place NCCL all-gather of 36 B/body",
                   "driver": "bh_mg_step (C++/NCCL behind the C ABI, csrc/bh_mg.cu); torch.distributed carries the NCCL id and reduces the timings",
                   "l2": "state far larger than L2 (>= 5 GB context at 16M bodies); no flush between steps"},
        "same_workload_1gpu": same1,
        "efficiency_same_workload": (same1["ms_per_step"] / (world * total_ms / args.steps)) if same1 else None,
        "interactions_per_body": float(inter.item()) / n,
        "interactions_per_s": float(inter.item()) * args.steps / (total_ms * 1e-3),
        "phase_ms_rank0": {k: round(v, 4) for k, v in acc.items()}, "force_ms_per_rank": force_per_rank,
        "allgather_ms": allgather_ms,
        "cells": cells, "e2e": e2e, "roofline": roofline,
        "gpu_launches": SLICED_LAUNCHES_PER_STEP * args.steps * world, "clocks": ck,
    }
    dist.destroy_process_group()
    return line if rank == 0 else None
