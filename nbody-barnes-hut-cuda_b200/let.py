"""Locally-essential-tree mode (BASELINE.json north_star "beyond ~100M bodies"; SURVEY §8e, second bullet).

Each rank owns a subset of the bodies (a Morton-key range at start-up) and never sees the others' bodies.
Per step (include/bh.h "locally-essential-tree exchange"):
  boxes      every rank's body AABB is all-gathered; their union gives the global cube (reference formula,
             nbody_v5_bench.cu:148-154) that all ranks fix, so Morton keys share one grid;
  local tree keys, sort, tree, centre of mass over the rank's own bodies;
  export     one walk of that tree per peer against the peer's box (csrc/bh_let.cu) -> point masses;
  exchange   all-to-all of the point lists (variable sizes);
  union step own bodies + received points (ids = -1) go through the ordinary step; received points are
             dropped afterwards, own bodies stay in Morton order for the next step.
`LetRank` is the per-rank logic; `let_step_emulated` drives several ranks on ONE device (tests — the
guide forbids emulating ranks with kernels that wait on each other, this path has no such kernels);
`LetSimulation` is the torch.distributed driver (NCCL all-gather + all-to-all).
"""
from __future__ import annotations

import numpy as np

from .engine import DBG, FLAG_NO_GRAPH, PHASE, BHEngine
from .sliced import _DevView

f32 = np.float32


def global_cube(boxes_lohi: np.ndarray) -> np.ndarray:
    """d_bounds of the union of the ranks' boxes, with the reference's float arithmetic (bench:148-154)."""
    b = np.asarray(boxes_lohi, f32).reshape(-1, 6)
    b = b[b[:, 0] <= b[:, 3]]
    lo = b[:, :3].min(0).astype(f32)
    hi = b[:, 3:].max(0).astype(f32)
    ext = (hi - lo).astype(f32)
    size = f32(max(ext[0], max(ext[1], ext[2])))
    return np.array([lo[0], lo[1], lo[2], f32(lo[0] + size), f32(lo[1] + size), f32(lo[2] + size)], f32)


def split_by_keys(keys: np.ndarray, world: int, sample: int = 1 << 20, seed: int = 0):
    """Sampled key splitters: returns `world` boolean masks, one per rank."""
    rng = np.random.default_rng(seed)
    samp = np.sort(keys if len(keys) <= sample else rng.choice(keys, sample, replace=False))
    cuts = [samp[(len(samp) * r) // world] for r in range(1, world)]
    edges = [0] + [int(c) for c in cuts] + [1 << 32]
    return [(keys >= edges[r]) & (keys < edges[r + 1]) for r in range(world)]


class LetRank:
    def __init__(self, bh, torch, device, posm, vel, ids, capacity: int, cap_per_peer: int, npeers: int, **params):
        self.bh, self.torch, self.device = bh, torch, device
        self.posm, self.vel, self.ids = posm, vel, ids          # device tensors: [n,4] f32, [n,4] f32, [n] i32
        self.capacity, self.cap, self.npeers = capacity, cap_per_peer, npeers
        dev_index = device.index if device.index is not None else 0
        self.eng = BHEngine(capacity, device=dev_index, flags=FLAG_NO_GRAPH, **params)
        self.out = torch.empty((npeers, cap_per_peer, 4), dtype=torch.float32, device=device)
        self.last = {}

    @property
    def n(self) -> int:
        return int(self.posm.shape[0])

    def _stream(self) -> int:
        return self.torch.cuda.current_stream().cuda_stream

    def local_box(self) -> np.ndarray:
        if self.n == 0:
            return np.array([1, 1, 1, -1, -1, -1], f32)     # lo > hi: "no bodies here"
        self.eng.import_state(self.posm, self.vel, self.ids, self.n, self._stream())
        return self.eng.local_bounds()

    def build_local_tree(self, cube: np.ndarray):
        self.eng.set_fixed_bounds(cube)
        if self.n == 0:
            return
        st = self._stream()
        for ph in (PHASE.KEYS, PHASE.SORT, PHASE.BUILD, PHASE.COM):
            self.eng.run_phase(ph, st)

    def export(self, boxes_lohi: np.ndarray, me: int) -> np.ndarray:
        boxes = np.asarray(boxes_lohi, f32).reshape(-1, 6).copy()
        boxes[me] = [1, 1, 1, -1, -1, -1]                    # nothing is exported to oneself
        if self.n == 0:
            return np.zeros(len(boxes), np.int32)
        return self.eng.let_export(boxes, self.out, self.cap, self._stream())

    def union_step(self, received):
        """received: list of [k,4] device tensors (point masses from the peers)."""
        torch = self.torch
        imp = [r for r in received if r.shape[0] > 0]
        n_imp = sum(int(r.shape[0]) for r in imp)
        nl = self.n
        nu = nl + n_imp
        if nu == 0:
            return
        if nu > self.capacity:
            raise self.bh.BHError(f"LET union of {nu} bodies exceeds the context capacity {self.capacity}")
        posm_u = torch.cat([self.posm] + imp) if imp else self.posm
        vel_u = torch.cat([self.vel, torch.zeros((n_imp, 4), dtype=torch.float32, device=self.device)]) if n_imp else self.vel
        ids_u = torch.cat([self.ids, torch.full((n_imp,), -1, dtype=torch.int32, device=self.device)]) if n_imp else self.ids
        st = self._stream()
        self.eng.import_state(posm_u.contiguous(), vel_u.contiguous(), ids_u.contiguous(), nu, st)
        self.eng.simulation_step(1, st)
        ptr = self.eng.state_ptrs()
        pu = torch.as_tensor(_DevView(ptr["posm"], (nu, 4), "<f4"), device=self.device)
        vu = torch.as_tensor(_DevView(ptr["vel"], (nu, 4), "<f4"), device=self.device)
        iu = torch.as_tensor(_DevView(ptr["ids"], (nu,), "<i4"), device=self.device)
        keep = iu >= 0
        self.posm, self.vel, self.ids = pu[keep].clone(), vu[keep].clone(), iu[keep].clone()
        self.last = {"n_union": nu, "n_import": n_imp}

    def last_accelerations(self):
        """(ids, acc[k,3]) of the own bodies for the step just taken (host arrays; test/diagnostic path)."""
        ids_s = self.eng.debug_get(DBG.IDS_SORTED)
        acc = self.eng.debug_get(DBG.ACC)
        keep = ids_s >= 0
        return ids_s[keep], acc[keep, :3]

    def close(self):
        self.eng.close()


def let_step_emulated(ranks):
    """One LET step of several LetRank objects living on one device; the all-to-all is a list shuffle."""
    boxes = np.stack([r.local_box() for r in ranks])
    cube = global_cube(boxes)
    for r in ranks:
        r.build_local_tree(cube)
    counts = [r.export(boxes, i) for i, r in enumerate(ranks)]
    sent = [[r.out[p, : int(counts[i][p])].clone() for p in range(len(ranks))] for i, r in enumerate(ranks)]
    for i, r in enumerate(ranks):
        r.union_step([sent[s][i] for s in range(len(ranks)) if s != i])
    return counts


class LetSimulation:
    """torch.distributed driver: one process per GPU, NCCL all-gather of boxes and all-to-all of point lists."""

    def __init__(self, bh, local_soa, local_ids, rank: int, world: int, local: int, dist, capacity: int,
                 cap_per_peer: int, **params):
        import torch

        self.torch, self.dist, self.rankno, self.world = torch, dist, rank, world
        self.device = torch.device(f"cuda:{local}")
        px, py, pz, vx, vy, vz, m = local_soa
        n = len(px)
        posm = torch.from_numpy(np.stack([px, py, pz, m], 1).astype(f32)).to(self.device)
        vel = torch.from_numpy(np.stack([vx, vy, vz, np.zeros(n, f32)], 1).astype(f32)).to(self.device)
        ids = torch.from_numpy(np.asarray(local_ids, np.int32)).to(self.device)
        self.rank = LetRank(bh, torch, self.device, posm, vel, ids, capacity, cap_per_peer, world, **params)
        self.stats = {}

    def step(self, nsteps: int = 1):
        torch, dist, w = self.torch, self.dist, self.world
        for _ in range(nsteps):
            box = torch.from_numpy(self.rank.local_box()).to(self.device)
            boxes = torch.empty((w, 6), dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(boxes, box)
            boxes_h = boxes.cpu().numpy()
            self.rank.build_local_tree(global_cube(boxes_h))
            counts = self.rank.export(boxes_h, self.rankno)
            send_counts = torch.from_numpy(counts.astype(np.int64)).to(self.device)
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts)
            sc, rc = counts.astype(np.int64).tolist(), recv_counts.cpu().numpy().tolist()
            send = torch.cat([self.rank.out[p, : sc[p]] for p in range(w)]) if sum(sc) else torch.empty((0, 4), dtype=torch.float32, device=self.device)
            recv = torch.empty((int(sum(rc)), 4), dtype=torch.float32, device=self.device)
            dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=rc, input_split_sizes=sc)
            self.rank.union_step([recv])
            self.stats = {"exported": int(sum(sc)), "imported": int(sum(rc)), "n_local": self.rank.n}

    def close(self):
        self.rank.close()


def _morton30_numpy(px, py, pz, cube):
    """Host-side copy of the reference key (bench:57-61) for the start-up partition only."""
    size = max(f32(cube[3] - cube[0]), f32(1.0))

    def q(p, lo):
        t = ((p - f32(lo)) / size * f32(1023.0)).astype(f32)
        return np.clip(t, 0, 1023).astype(np.uint32)

    def spread(v):
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v

    return (spread(q(px, cube[0])) << np.uint32(2)) | (spread(q(py, cube[1])) << np.uint32(1)) | spread(q(pz, cube[2]))


def run_let_bench(args, w, bh, dist, rank, world, local):
    """bench.py leg for body counts that are not replicated (default above 100M bodies)."""
    import time

    import torch

    import bench

    n = w["n"]
    soa = bench.make_ic(bh, w)                     # every rank generates the same bodies, keeps its key range
    lo = np.array([soa[a].min() for a in range(3)], f32)
    hi = np.array([soa[a].max() for a in range(3)], f32)
    cube = global_cube(np.concatenate([lo, hi])[None, :])
    keys = _morton30_numpy(soa[0], soa[1], soa[2], cube)
    mask = split_by_keys(keys, world)[rank]
    sel = np.nonzero(mask)[0]
    del keys, mask
    local_soa = [a[sel] for a in soa]
    del soa
    nl = len(sel)
    cap_peer = max(1 << 20, int(0.15 * n / world))
    sim = LetSimulation(bh, local_soa, sel.astype(np.int32), rank, world, local, dist,
                        capacity=int(1.5 * n / world) + (world - 1) * cap_peer // 2 + 4096, cap_per_peer=cap_peer)
    dev = sim.device

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    sim.step(args.warmup)
    barrier()
    sim.rank.eng.check_device_error()
    clocks = bench.ClockSampler(local)
    clocks.start()
    barrier()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    sim.step(args.steps)
    b.record()
    barrier()
    wall = time.perf_counter() - t0
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ck = clocks.stop()
    eng = sim.rank.eng
    inter = torch.tensor([float(eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY))], device=dev,
                         dtype=torch.float64)
    dist.all_reduce(inter)
    st = torch.tensor([float(sim.stats["exported"]), float(sim.stats["imported"]), float(sim.stats["n_local"])], device=dev,
                      dtype=torch.float64)
    stats_all = torch.empty((world, 3), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(stats_all, st)
    sim.close()
    total_ms = float(ms.item())
    line = {
        "metric": "body-steps/s", "value": n * args.steps / (total_ms * 1e-3), "unit": "body-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "n_bodies": n, "theta": 0.5, "G": 0.5, "dt": 0.02,
                   "softening": 50.0, "max_speed": 500.0, "group": 32,
                   "parallelism": f"locally-essential-tree x{world}: sampled key-range ownership, per-peer export walk, "
                                  "NCCL all-to-all of point masses, ordinary step on own + imported bodies",
                   "l2": "state far larger than L2; no flush between steps"},
        "interactions_per_body": float(inter.item()) / n, "interactions_per_s": float(inter.item()) * args.steps / (total_ms * 1e-3),
        "let_points_exported_imported_local_per_rank": [[int(x) for x in row] for row in stats_all.tolist()],
        "wall_s_timed_loop": wall, "e2e": None, "gpu_launches": None, "clocks": ck,
    }
    dist.destroy_process_group()
    return line if rank == 0 else None
