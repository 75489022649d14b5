"""Locally-essential-tree mode (BASELINE.json north_star "beyond ~100M bodies"; SURVEY §8e, second bullet).

Each rank owns the bodies of one Morton-key range and never sees the others' bodies.  Per step
(include/bh.h "locally-essential-tree exchange"):
  cube       every rank's body AABB is all-gathered; their union gives the global cube (reference formula,
             nbody_v5_bench.cu:148-154) that all ranks fix, so Morton keys share one grid;
  splitters  keys + sort of the own bodies; every body carries the work its traversal chunk cost in the
             previous step (interaction-list entries, acc.w -> vel.w); every rank contributes SAMPLE keys
             taken at equal increments of its cumulative work; the weighted quantiles of the pooled sample
             are the new key ranges (SURVEY §8e "sampled key splitters") — equal WORK, not equal count;
  migrate    the bodies that left the rank's range are runs of its sorted order: one all-to-all;
             (cube, election and migration run every INTERVAL steps, on a cube enlarged by the distance a body
             can travel in that time so that the key grid — and with it the meaning of a key range — stays put;
             in between a rank keeps its bodies, and the few that crossed out of its key range ("strays", at
             the two ends of its sorted order) are added to the NEAREST of its boxes.  Putting them into boxes
             of their own key intervals does not work: those are no longer runs of whole cells, they span the
             domain faces and the export lists overflow — measured.)
  local tree keys, sort, tree, centre of mass over the rank's own bodies;
  domain     the key range is cut at octree-cell boundaries (domain_cuts) and described by the tight body
             AABB of every interval (bh_let_domain_boxes) — a key range is not convex, whole cells are;
  export     one walk of the local tree per peer against the peer's boxes (csrc/bh_let.cu) -> point masses;
  exchange   all-to-all of the point lists (variable sizes);
  forces     traversal of the local tree; the received points get a small tree of their own (second context,
             same cube) that the rank's body groups traverse as well (bh_force_from, accumulating) — the own
             tree is never rebuilt for them; then the ordinary kick-drift-clamp.
`LetRank` is the per-rank logic; `let_step_emulated` drives several ranks on ONE device (tests — the
guide forbids emulating ranks with kernels that wait on each other, this path has no such kernels);
`LetSimulation` is the torch.distributed driver (NCCL all-gather + all-to-all).
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from .engine import DBG, FLAG_NO_GRAPH, PHASE, BHEngine
from .engine import lib as _lib
from .sliced import _DevView

f32 = np.float32
KEY_END = 1 << 30      # one past the largest 30-bit key
SAMPLE = 4096          # keys every rank contributes to the splitter election
WORK_FLOOR = 300.0     # added to every body's measured work: the per-body cost of the non-traversal phases in units of
                       # list entries (256M two-disc on 8 GPUs: ~10 ms of sort+tree per 27M bodies vs 1.1 ps per entry)
INTERVAL = 4           # steps between cube / election / migration rounds
MAX_BOXES = 200        # boxes that describe one rank's domain (<= BH_LET_MAX_BOXES)
EMPTY_BOX = np.array([1, 1, 1, -1, -1, -1], f32)   # lo > hi


def global_cube(boxes_lohi: np.ndarray) -> np.ndarray:
    """d_bounds of the union of the ranks' boxes, with the reference's float arithmetic (bench:148-154)."""
    b = np.asarray(boxes_lohi, f32).reshape(-1, 6)
    b = b[b[:, 0] <= b[:, 3]]
    lo = b[:, :3].min(0).astype(f32)
    hi = b[:, 3:].max(0).astype(f32)
    ext = (hi - lo).astype(f32)
    size = f32(max(ext[0], max(ext[1], ext[2])))
    return np.array([lo[0], lo[1], lo[2], f32(lo[0] + size), f32(lo[1] + size), f32(lo[2] + size)], f32)


def expand_cube(cube: np.ndarray, margin: float) -> np.ndarray:
    """The cube grown by `margin` on every side (bodies may move that far before the next election)."""
    c = np.asarray(cube, f32)
    lo = (c[:3] - f32(margin)).astype(f32)
    size = f32(f32(c[3] - c[0]) + f32(2.0) * f32(margin))
    return np.array([lo[0], lo[1], lo[2], f32(lo[0] + size), f32(lo[1] + size), f32(lo[2] + size)], f32)


def split_by_keys(keys: np.ndarray, world: int, sample: int = 1 << 20, seed: int = 0):
    """Start-up partition by sampled key splitters: returns `world` boolean masks."""
    rng = np.random.default_rng(seed)
    samp = np.sort(keys if len(keys) <= sample else rng.choice(keys, sample, replace=False))
    cuts = [samp[(len(samp) * r) // world] for r in range(1, world)]
    edges = [0] + [int(c) for c in cuts] + [1 << 32]
    return [(keys >= edges[r]) & (keys < edges[r + 1]) for r in range(world)]


def elect_splitters(samples: np.ndarray, work: np.ndarray) -> np.ndarray:
    """Key-range edges [world+1] from every rank's key sample (bh_let_elect_splitters, csrc/bh_let_host.cpp).

    samples [world, SAMPLE]: keys of rank r at equal increments of its cumulative work, so that every sample
    stands for work[r]/SAMPLE (rows of ranks with work 0 are ignored).  Edge r is the key at r/world of the
    pooled work.  Identical on all ranks (pure function of all-gathered data)."""
    samples = np.ascontiguousarray(samples, np.int64)
    work = np.ascontiguousarray(work, np.float64)
    world = len(work)
    edges = np.zeros(world + 1, np.int64)
    rc = _lib().bh_let_elect_splitters(samples.ctypes.data_as(C.c_void_p), samples.shape[1], work.ctypes.data_as(C.c_void_p),
                                       world, edges.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"bh_let_elect_splitters failed: {rc}")
    return edges


def domain_cuts(k_lo: int, k_hi: int) -> np.ndarray:
    """MAX_BOXES+1 ascending keys that cut [k_lo, k_hi) at octree-cell boundaries (bh_let_domain_cuts).

    Interior: the 8..64 cells of size S = 8^j (largest with span/S >= 8) that lie inside the range — whole
    cells, convex, owned by this rank alone.  The two ragged ends (parts of one S-cell each) are cut again
    at S/64 so that their boxes reach at most one small cell into the neighbour's range.  Padded with k_hi
    (empty intervals) to a fixed length so the boxes can be all-gathered."""
    cuts = np.zeros(MAX_BOXES + 1, np.uint32)
    rc = _lib().bh_let_domain_cuts(int(k_lo), int(k_hi), MAX_BOXES, cuts.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"bh_let_domain_cuts({k_lo}, {k_hi}) failed: {rc}")
    return cuts


def compact_boxes(boxes: np.ndarray) -> np.ndarray:
    """[world, MAX_BOXES, 6] -> [world, K, 6]: used boxes first, K = the largest used count (>= 1)."""
    used = boxes[:, :, 0] <= boxes[:, :, 3]
    k = max(1, int(used.sum(1).max()))
    out = np.tile(EMPTY_BOX, (boxes.shape[0], k, 1))
    for r in range(boxes.shape[0]):
        b = boxes[r][used[r]]
        out[r, : len(b)] = b
    return out


class LetRank:
    """One rank's bodies and engine context.  The own bodies live in one of two capacity-sized buffer sets
    (ping-pong: migration and ghost removal write the other set), so a step allocates nothing."""

    def __init__(self, bh, torch, device, posm, vel, ids, capacity: int, cap_per_peer: int, npeers: int, **params):
        self.bh, self.torch, self.device = bh, torch, device
        self.capacity, self.cap, self.npeers = capacity, cap_per_peer, npeers
        dev_index = device.index if device.index is not None else 0
        self.eng = BHEngine(capacity, device=dev_index, flags=FLAG_NO_GRAPH, **params)
        # the points imported from the peers live in a context of their own (see forces_and_update)
        self.ghost_cap = max(2, min(capacity, (npeers - 1) * cap_per_peer))
        self.geng = BHEngine(self.ghost_cap, device=dev_index, flags=FLAG_NO_GRAPH, **params)
        self.gposm = torch.zeros((self.ghost_cap, 4), dtype=torch.float32, device=device)
        self.gvel = torch.zeros((self.ghost_cap, 4), dtype=torch.float32, device=device)
        self.gids = torch.full((self.ghost_cap,), -1, dtype=torch.int32, device=device)
        self.out = torch.empty((npeers, cap_per_peer, 4), dtype=torch.float32, device=device)
        self._sets = [(torch.empty((capacity, 4), dtype=torch.float32, device=device),
                       torch.empty((capacity, 4), dtype=torch.float32, device=device),
                       torch.empty((capacity,), dtype=torch.int32, device=device)) for _ in range(2)]
        self._cur, self._n = 0, 0
        self.last = {}
        self._ev = None
        self.adopt(posm, vel, ids)

    @property
    def n(self) -> int:
        return self._n

    # views of the own bodies: device tensors [n,4] f32 {x,y,z,m}, [n,4] f32 {vx,vy,vz,work}, [n] i32
    @property
    def posm(self):
        return self._sets[self._cur][0][: self._n]

    @property
    def vel(self):
        return self._sets[self._cur][1][: self._n]

    @property
    def ids(self):
        return self._sets[self._cur][2][: self._n]

    def spare(self, rows: int):
        """Views [rows] of the buffer set that is NOT holding the own bodies (receive target of a migration)."""
        if rows > self.capacity:
            raise self.bh.BHError(f"{rows} bodies exceed the LET context capacity {self.capacity}")
        p, v, i = self._sets[self._cur ^ 1]
        return p[:rows], v[:rows], i[:rows]

    def adopt(self, posm, vel, ids):
        """The given bodies become the own bodies (copied into the spare set unless they already are its views)."""
        rows = int(posm.shape[0])
        p, v, i = self.spare(rows)
        if rows and posm.data_ptr() != p.data_ptr():
            p.copy_(posm); v.copy_(vel); i.copy_(ids)
        self._cur ^= 1
        self._n = rows

    def _stream(self) -> int:
        return self.torch.cuda.current_stream().cuda_stream

    def local_box(self) -> np.ndarray:
        if self.n == 0:
            return EMPTY_BOX.copy()
        self.eng.import_state(self.posm, self.vel, self.ids, self.n, self._stream())
        return self.eng.local_bounds()

    def fix_cube(self, cube: np.ndarray):
        self.cube = np.asarray(cube, f32).copy()
        self.eng.set_fixed_bounds(self.cube)
        self.geng.set_fixed_bounds(self.cube)

    # ---- ownership: splitter election and migration ------------------------------------------
    def sort_own(self, by_work: bool = True):
        """Key + sort the own bodies (after fix_cube + local_box).  Returns (SAMPLE sorted keys taken at equal increments
        of the cumulative work the bodies carry in vel.w, total work); bodies without a measurement yet (first
        step) or by_work=False count 1 each."""
        if self.n == 0:
            return np.zeros(SAMPLE, np.int64), 0.0
        torch, st = self.torch, self._stream()
        self.eng.sort_coarse(st)                                # state imported by local_box(); 30-bit keys are enough here
        keys, _, vel, _ = self._sorted()
        work = vel[:, 3].double() + WORK_FLOOR if by_work else None
        if work is None or float(vel[:, 3].max().item()) <= 0.0:
            pick = (torch.arange(SAMPLE, device=self.device, dtype=torch.int64) * (self.n - 1)) // (SAMPLE - 1)   # exact: no float index
            return keys[pick].cpu().numpy().astype(np.int64), float(self.n)
        cum = torch.cumsum(work, 0)
        total = float(cum[-1].item())
        targets = (torch.arange(SAMPLE, device=self.device, dtype=torch.float64) + 0.5) * (total / SAMPLE)
        pick = torch.searchsorted(cum, targets).clamp_(max=self.n - 1)
        return keys[pick].cpu().numpy().astype(np.int64), total

    def _sorted(self):
        """Views of this step's Morton-sorted keys / posm / vel / ids (valid until the next import)."""
        torch, p = self.torch, self.eng.sorted_ptrs()
        n = p["n"]
        return (torch.as_tensor(_DevView(p["keys"], (n,), "<i4"), device=self.device),
                torch.as_tensor(_DevView(p["posm"], (n, 4), "<f4"), device=self.device),
                torch.as_tensor(_DevView(p["vel"], (n, 4), "<f4"), device=self.device),
                torch.as_tensor(_DevView(p["ids"], (n,), "<i4"), device=self.device))

    def migration_plan(self, edges: np.ndarray):
        """Rows of the sorted order that go to every rank: (send_counts [world] int64, posm, vel, ids views)."""
        world = len(edges) - 1
        if self.n == 0:
            return np.zeros(world, np.int64), None
        keys, posm, vel, ids = self._sorted()
        inner = self.torch.tensor(edges[1:-1].astype(np.int32), dtype=self.torch.int32, device=self.device)
        pos = self.torch.searchsorted(keys, inner).cpu().numpy().astype(np.int64)
        bounds = np.concatenate([[0], pos, [self.n]])
        return np.diff(bounds), (posm, vel, ids)

    # ---- local tree, domain description, export ----------------------------------------------
    def build_local_tree(self):
        self.last = {}
        if self.n == 0:
            return
        st = self._stream()
        self.eng.import_state(self.posm, self.vel, self.ids, self.n, st)
        for ph in (PHASE.KEYS, PHASE.SORT, PHASE.BUILD, PHASE.COM):
            self.eng.run_phase(ph, st)

    def travel_margin(self, steps: int) -> float:
        """How far a body can get in `steps` steps (speed clamp x dt): the cube is enlarged by this much."""
        return float(self.eng.params.max_speed) * float(self.eng.params.dt) * steps

    def domain_boxes(self, k_lo: int, k_hi: int) -> np.ndarray:
        """[MAX_BOXES, 6] tight body AABBs of the octree-aligned key intervals of [k_lo, k_hi) (see domain_cuts);
        own bodies whose key left the range since the last migration are added to the nearest box."""
        if self.n == 0:
            return np.tile(EMPTY_BOX, (MAX_BOXES, 1))
        boxes, counts = self.eng.let_domain_boxes(domain_cuts(k_lo, k_hi))
        if int(counts.sum()) != self.n:
            self._add_strays(boxes, counts, k_lo)
        return boxes

    def _add_strays(self, boxes: np.ndarray, counts: np.ndarray, k_lo: int):
        """The bodies outside [k_lo, k_hi) are the two end runs of the sorted order; each one extends the box it is
        nearest to (they crossed a face of the domain a few steps ago, so that box hardly grows)."""
        torch, dev = self.torch, self.device
        keys, posm, _, _ = self._sorted()
        a = int(torch.searchsorted(keys, torch.tensor([k_lo], dtype=torch.int32, device=dev)).item())
        b = a + int(counts.sum())
        strays = torch.cat([posm[:a, :3], posm[b:, :3]])
        used = np.nonzero(counts > 0)[0]
        self.last["strays"] = int(strays.shape[0])
        if len(used) == 0:                                      # everything strayed: one box around all of it
            boxes[0, :3] = strays.amin(0).cpu().numpy()
            boxes[0, 3:] = strays.amax(0).cpu().numpy()
            return
        B = torch.from_numpy(boxes[used]).to(dev)
        lo, hi = B[:, :3].clone(), B[:, 3:].clone()
        for c0 in range(0, int(strays.shape[0]), 16384):
            sc = strays[c0:c0 + 16384]
            d = torch.clamp(torch.maximum(B[None, :, :3] - sc[:, None, :], sc[:, None, :] - B[None, :, 3:]), min=0)
            idx = (d * d).sum(2).argmin(1)[:, None].expand(-1, 3)
            lo.scatter_reduce_(0, idx, sc, "amin", include_self=True)
            hi.scatter_reduce_(0, idx, sc, "amax", include_self=True)
        boxes[used, :3] = lo.cpu().numpy()
        boxes[used, 3:] = hi.cpu().numpy()

    def export(self, boxes_lohi: np.ndarray, me: int) -> np.ndarray:
        """boxes_lohi: [world, K, 6].  Returns the number of points emitted for every peer."""
        boxes = np.asarray(boxes_lohi, f32).copy()
        boxes[me] = EMPTY_BOX                                   # nothing is exported to oneself
        if self.n == 0:
            return np.zeros(len(boxes), np.int32)
        return self.eng.let_export(boxes, self.out, self.cap, self._stream())

    def ghost_slot(self, rows: int):
        """View [rows,4] of the ghost buffer: imported points received there need no further copy."""
        if rows > self.ghost_cap:
            raise self.bh.BHError(f"{rows} imported points exceed the ghost capacity {self.ghost_cap}")
        return self.gposm[:rows]

    def forces_and_update(self, received):
        """Forces on the own bodies from the local tree (built by build_local_tree) and from the imported
        points, then kick-drift-clamp.  received: list of [k,4] device tensors (point masses from the peers), or
        the number of points already written to ghost_slot()."""
        torch = self.torch
        if isinstance(received, int):
            n_imp = received
        else:
            n_imp = sum(int(r.shape[0]) for r in received)
            slot, o = self.ghost_slot(n_imp), 0
            for r in received:
                slot[o: o + int(r.shape[0])].copy_(r)
                o += int(r.shape[0])
        self.last["n_import"] = n_imp
        if self.n == 0:
            return
        st = self._stream()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        self.eng.run_phase(PHASE.FORCE, st)
        if n_imp > 0:
            if n_imp == 1:                                      # a tree needs two bodies: split the point in two halves
                self.gposm[1] = self.gposm[0]
                self.gposm[:2, 3] *= 0.5
                n_imp = 2
            self.geng.import_state(self.gposm, self.gvel, self.gids, n_imp, st)
            for ph in (PHASE.KEYS, PHASE.SORT, PHASE.BUILD, PHASE.COM):
                self.geng.run_phase(ph, st)
            self.eng.force_from(self.geng, st)
        self.eng.run_phase(PHASE.UPDATE, st)
        ev[1].record()
        self._ev = (ev[0], ev[1], self.n)
        sp, sv, si = self._sets[self._cur ^ 1]                  # own bodies in Morton order, vel.w = the work they cost
        self._n = self.eng.export_real(sp, sv, si, st)
        self._cur ^= 1

    def last_forces_ms(self) -> float:
        """Device time of the last forces_and_update (both traversals, ghost tree, update); synchronises on its end event."""
        if self._ev is None:
            return 0.0
        self._ev[1].synchronize()
        return float(self._ev[0].elapsed_time(self._ev[1]))

    def last_accelerations(self):
        """(ids, acc[k,3]) of the own bodies for the step just taken (host arrays; test/diagnostic path)."""
        ids_s = self.eng.debug_get(DBG.IDS_SORTED)
        acc = self.eng.debug_get(DBG.ACC)
        keep = ids_s >= 0
        return ids_s[keep], acc[keep, :3]

    def close(self):
        self.eng.close()
        self.geng.close()


def let_step_emulated(ranks, rebalance: bool = True, elect: bool = True, edges=None, margin: float = 0.0):
    """One LET step of several LetRank objects living on one device; the all-to-alls are list shuffles.
    elect=False: no cube / election / migration — the ranks keep their bodies, the cube and `edges` of the last
    election step (which must have used a margin that covers the travel since).
    Returns (export counts per rank, key-range edges)."""
    torch = ranks[0].torch
    world = len(ranks)
    if elect:
        cube = expand_cube(global_cube(np.stack([r.local_box() for r in ranks])), margin)
        for r in ranks:
            r.fix_cube(cube)
        sw = [r.sort_own(rebalance) for r in ranks]
        edges = elect_splitters(np.stack([x[0] for x in sw]), np.array([x[1] for x in sw]))
        plans = [r.migration_plan(edges) for r in ranks]
        for dst, r in enumerate(ranks):
            parts = [[], [], []]
            for src in range(world):
                sc, views = plans[src]
                if views is None or sc[dst] == 0:
                    continue
                a = int(sc[:dst].sum())
                for k in range(3):
                    parts[k].append(views[k][a:a + int(sc[dst])].clone())
            if parts[0]:
                new = [torch.cat(p) for p in parts]
            else:
                new = [torch.empty((0, 4), dtype=torch.float32, device=r.device), torch.empty((0, 4), dtype=torch.float32, device=r.device),
                       torch.empty((0,), dtype=torch.int32, device=r.device)]
            r._new = new
        for r in ranks:                                        # adopt only after every rank's runs were copied out
            r.adopt(*r._new)
            del r._new
    assert edges is not None, "a step without election needs the edges of the last election"
    for r in ranks:
        r.build_local_tree()
    boxes = compact_boxes(np.stack([r.domain_boxes(int(edges[i]), int(edges[i + 1])) for i, r in enumerate(ranks)]))
    counts = [r.export(boxes, i) for i, r in enumerate(ranks)]
    sent = [[r.out[p, : int(counts[i][p])].clone() for p in range(world)] for i, r in enumerate(ranks)]
    for i, r in enumerate(ranks):
        r.forces_and_update([sent[s][i] for s in range(world) if s != i])
    return counts, edges


class LetSimulation:
    """torch.distributed driver: one process per GPU, NCCL all-gather of boxes and all-to-all of point lists."""

    def __init__(self, bh, local_soa, local_ids, rank: int, world: int, local: int, dist, capacity: int,
                 cap_per_peer: int, rebalance: bool = True, interval: int = INTERVAL, **params):
        import torch

        self.torch, self.dist, self.rankno, self.world = torch, dist, rank, world
        self.device = torch.device(f"cuda:{local}")
        self.rebalance, self.interval = rebalance, max(1, int(interval))
        self.steps_done, self.edges = 0, None
        px, py, pz, vx, vy, vz, m = local_soa
        n = len(px)
        posm = torch.from_numpy(np.stack([px, py, pz, m], 1).astype(f32)).to(self.device)
        vel = torch.from_numpy(np.stack([vx, vy, vz, np.zeros(n, f32)], 1).astype(f32)).to(self.device)
        ids = torch.from_numpy(np.asarray(local_ids, np.int32)).to(self.device)
        self.rank = LetRank(bh, torch, self.device, posm, vel, ids, capacity, cap_per_peer, world, **params)
        del posm, vel, ids
        self.stats = {}
        self.trace = {} if os.environ.get("BH_LET_TRACE") else None   # phase -> wall ms of the last step (syncs!)

    def _mark(self, name, t0):
        if self.trace is None:
            return t0
        self.torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.trace[name] = self.trace.get(name, 0.0) + 1e3 * (t1 - t0)
        return t1

    def _all_to_all_rows(self, send, sc, rc, shape_tail, dtype, out=None):
        recv = out if out is not None else self.torch.empty((int(sum(rc)),) + shape_tail, dtype=dtype, device=self.device)
        if send is None:
            send = self.torch.empty((0,) + shape_tail, dtype=dtype, device=self.device)
        self.dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=rc, input_split_sizes=sc)
        return recv

    def _agree(self, ok: bool, what: str):
        """A capacity error seen by ONE rank must be raised by ALL of them, or the others block in the next
        collective: one MIN all-reduce of the local verdict, then everybody raises (or nobody)."""
        flag = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            raise self.rank.bh.BHError(f"{what} (on rank {self.rankno}: {'ok' if ok else 'FAILED'}); every rank stops together")

    def step(self, nsteps: int = 1):
        torch, dist, w, me = self.torch, self.dist, self.world, self.rankno
        for _ in range(nsteps):
            if self.trace is not None:
                self.trace.clear()
                torch.cuda.synchronize()
            t = time.perf_counter()
            migrated = 0
            if self.steps_done % self.interval == 0:
                # global cube, enlarged by what a body can travel before the next election
                box = torch.from_numpy(self.rank.local_box()).to(self.device)
                boxes = torch.empty((w, 6), dtype=torch.float32, device=self.device)
                dist.all_gather_into_tensor(boxes, box)
                margin = self.rank.travel_margin(self.interval) if self.interval > 1 else 0.0
                self.rank.fix_cube(expand_cube(global_cube(boxes.cpu().numpy()), margin))
                t = self._mark("cube", t)
                # splitter election: every rank's work-spaced key sample + its total work
                sample, work = self.rank.sort_own(self.rebalance)
                mine = torch.from_numpy(np.concatenate([sample.astype(np.float64), [work]])).to(self.device)
                pooled = torch.empty((w, SAMPLE + 1), dtype=torch.float64, device=self.device)
                dist.all_gather_into_tensor(pooled, mine)
                pooled = pooled.cpu().numpy()
                self.edges = elect_splitters(pooled[:, :SAMPLE].astype(np.int64), pooled[:, SAMPLE])
                t = self._mark("sort+elect", t)
                # migration: runs of the sorted order
                sc_np, views = self.rank.migration_plan(self.edges)
                send_counts = torch.from_numpy(sc_np).to(self.device)
                recv_counts = torch.empty_like(send_counts)
                dist.all_to_all_single(recv_counts, send_counts)
                sc, rc = sc_np.tolist(), recv_counts.cpu().numpy().tolist()
                migrated = int(sum(sc)) - int(sc[me])
                v = views if views is not None else (None, None, None)
                self._agree(int(sum(rc)) <= self.rank.capacity, f"migration: {int(sum(rc))} bodies exceed the context capacity {self.rank.capacity}")
                tp, tv, ti = self.rank.spare(int(sum(rc)))
                self._all_to_all_rows(v[0], sc, rc, (4,), torch.float32, out=tp)
                self._all_to_all_rows(v[1], sc, rc, (4,), torch.float32, out=tv)
                self._all_to_all_rows(v[2], sc, rc, (), torch.int32, out=ti)
                self.rank.adopt(tp, tv, ti)
                t = self._mark("migrate", t)
            edges = self.edges
            # local tree, domain boxes, export
            self.rank.build_local_tree()
            t = self._mark("local tree", t)
            dom = torch.from_numpy(self.rank.domain_boxes(int(edges[me]), int(edges[me + 1]))).to(self.device)
            doms = torch.empty((w, MAX_BOXES, 6), dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(doms, dom)
            t = self._mark("domain boxes", t)
            try:
                counts, export_err = self.rank.export(compact_boxes(doms.cpu().numpy()), me), None
            except self.rank.bh.BHError as e:   # list or queue overflow (cap_per_peer too small)
                counts, export_err = np.zeros(w, np.int32), e
            self._agree(export_err is None, f"LET export: {export_err}")
            t = self._mark("export walk", t)
            send_counts = torch.from_numpy(counts.astype(np.int64)).to(self.device)
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts)
            sc, rc = counts.astype(np.int64).tolist(), recv_counts.cpu().numpy().tolist()
            send = torch.cat([self.rank.out[p, : sc[p]] for p in range(w)]) if sum(sc) else None
            n_imp = int(sum(rc))
            self._agree(n_imp <= self.rank.ghost_cap, f"LET import: {n_imp} points exceed the ghost capacity {self.rank.ghost_cap}")
            self._all_to_all_rows(send, sc, rc, (4,), torch.float32, out=self.rank.ghost_slot(n_imp))
            t = self._mark("exchange", t)
            self.rank.forces_and_update(n_imp)
            t = self._mark("forces+update", t)
            self.steps_done += 1
            if self.trace is not None:
                self.trace_all = getattr(self, "trace_all", []) + [dict(self.trace)]
                if "migrate" in self.trace:
                    self.trace_migration_step = dict(self.trace)
            self.stats = {"exported": int(sum(sc)), "imported": int(sum(rc)), "n_local": self.rank.n, "migrated_out": migrated,
                          "strays": self.rank.last.get("strays", 0)}

    def step_host(self, host):
        """End-to-end step: the rank's bodies come from pinned HOST buffers and go back there.
        host = dict(posm [cap,4] f32, vel [cap,4] f32, ids [cap] i32, n) — pinned tensors sized for the rank's
        capacity; n bodies are valid.  Uploads them, takes one step (the rank may own different bodies
        afterwards), downloads the new own bodies and updates host["n"].  Returns (h2d_bytes, d2h_bytes)."""
        torch, r = self.torch, self.rank
        n0 = int(host["n"])
        tp, tv, ti = r.spare(n0)
        tp.copy_(host["posm"][:n0], non_blocking=True)
        tv.copy_(host["vel"][:n0], non_blocking=True)
        ti.copy_(host["ids"][:n0], non_blocking=True)
        r.adopt(tp, tv, ti)
        self.step(1)
        n1 = r.n
        host["posm"][:n1].copy_(r.posm, non_blocking=True)
        host["vel"][:n1].copy_(r.vel, non_blocking=True)
        host["ids"][:n1].copy_(r.ids, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host["n"] = n1
        return 36 * n0, 36 * n1

    def host_buffers(self):
        """Pinned host mirror of the rank's bodies for step_host."""
        torch, r = self.torch, self.rank
        cap = r.capacity
        host = {"posm": torch.empty((cap, 4), dtype=torch.float32).pin_memory(),
                "vel": torch.empty((cap, 4), dtype=torch.float32).pin_memory(),
                "ids": torch.empty((cap,), dtype=torch.int32).pin_memory(), "n": r.n}
        host["posm"][: r.n].copy_(r.posm)
        host["vel"][: r.n].copy_(r.vel)
        host["ids"][: r.n].copy_(r.ids)
        torch.cuda.synchronize()
        return host

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


def launches_per_let_step(key_bits: int) -> int:
    """This library's kernel launches in one LET step of one rank that received imports (torch's and NCCL's own
    kernels not counted): bounds 3, coarse election sort 8, local tree (keys, sort, reorder, tree 5, centre of mass 3), domain
    boxes 3, export walk (seed + one per level), force 2, ghost tree, cross-tree force 2, update 1, compaction 3."""
    levels = key_bits // 3
    sort = 6 if key_bits == 30 else 14        # histogram + scan + 4 passes; twice + gather + combine for 60 bits
    tree = 1 + sort + 1 + 5 + 3
    return 3 + 8 + tree + 3 + (1 + levels + 1) + 2 + tree + 2 + 1 + 3


def run_let_bench(args, w, bh, dist, rank, world, local):
    """bench.py leg for body counts that are not replicated (default above 100M bodies)."""
    import time

    import torch

    import bench

    n = w["n"]
    # Start-up ownership is an INDEX range (no key is needed): the first step's migration hands every body to
    # the owner of its key.  The two-disc generator is counter based, so a rank draws only its own share.
    first, last = n * rank // world, n * (rank + 1) // world
    if w["ic"] == "twodisk":
        local_soa = bh.ic_two_disks_range(first, last - first, 42, 4000.0, 20.0, 8.0)
    else:
        soa = bench.make_ic(bh, w)
        local_soa = [a[first:last].copy() for a in soa]
        del soa
    sel = np.arange(first, last, dtype=np.int32)
    key_bits = getattr(args, "key_bits", None) or (60 if n > 100_000_000 else 30)
    cap_peer = max(1 << 20, int(0.15 * n / world))
    # work-balanced key ranges may hold up to ~2x the mean body count; imports come on top
    sim = LetSimulation(bh, local_soa, sel, rank, world, local, dist,
                        capacity=int(2.2 * n / world) + (world - 1) * cap_peer // 2 + 4096, cap_per_peer=cap_peer,
                        rebalance=not getattr(args, "let_no_rebalance", False),
                        interval=getattr(args, "let_interval", None) or INTERVAL, key_bits=key_bits)
    dev = sim.device

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # the warm-up covers the first work-weighted election (start-up ownership -> equal counts -> equal work)
    # and one step on the elected ranges; the timed steps start on an election step
    sim.step(max(args.warmup, sim.interval + 1))
    while sim.steps_done % sim.interval:
        sim.step(1)
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
    st = torch.tensor([float(sim.stats["exported"]), float(sim.stats["imported"]), float(sim.stats["n_local"]),
                       float(sim.stats["migrated_out"]), sim.rank.last_forces_ms()], device=dev, dtype=torch.float64)
    stats_all = torch.empty((world, 5), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(stats_all, st)
    # ---- e2e: every rank's bodies come from pinned host memory and go back there, every step
    e2e = None
    if not getattr(args, "no_e2e", False):
        host = sim.host_buffers()
        esteps = max(3, min(args.steps, 5))
        sim.step_host(host)
        barrier()
        t1 = time.perf_counter()
        nbytes = [0, 0]
        for _ in range(esteps):
            hb, db = sim.step_host(host)
            nbytes[0] += hb; nbytes[1] += db
        barrier()
        e2e_s = torch.tensor([(time.perf_counter() - t1) / esteps], device=dev, dtype=torch.float64)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        tot = torch.tensor([float(nbytes[0]), float(nbytes[1])], device=dev, dtype=torch.float64)
        dist.all_reduce(tot)
        e2e = {"value": n / float(e2e_s.item()), "unit": "body-steps/s", "h2d_bytes_per_step": int(tot[0].item() / esteps),
               "d2h_bytes_per_step": int(tot[1].item() / esteps), "ms_per_step": float(e2e_s.item()) * 1e3,
               "api": "LetSimulation.step_host: each rank uploads its own bodies (posm, vel, ids: 36 B/body) from pinned "
                      "host memory, one LET step, downloads the bodies it owns afterwards"}
        del host
    sim.close()
    total_ms = float(ms.item())
    # the force phases include the small ghost tree and the update: a lower bound of the traversal's rate
    roofline = bench.force_roofline(bh, local, float(inter.item()) / world, max(row[4] for row in stats_all.tolist()))
    line = {
        "metric": "body-steps/s", "value": n * args.steps / (total_ms * 1e-3), "unit": "body-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench.workload_config(args, w),
        "engine": {"group": 32, "key_bits": key_bits, "let_interval": sim.interval,
                   "parallelism": f"locally-essential-tree x{world}: work-weighted sampled key splitters, body migration, "
                                  "per-peer export walk against octree-aligned domain boxes, NCCL all-to-all of point "
                                  "masses, traversal of the own tree + of a small tree of the imported points",
                   "l2": "state far larger than L2; no flush between steps"},
        "interactions_per_body": float(inter.item()) / n, "interactions_per_s": float(inter.item()) * args.steps / (total_ms * 1e-3),
        "force_share": max(row[4] for row in stats_all.tolist()) / (total_ms / args.steps),
        "let_exported_imported_local_migrated_per_rank": [[int(x) for x in row[:4]] for row in stats_all.tolist()],
        "strays_rank0_last_step": int(sim.stats.get("strays", 0)),
        "forces_update_ms_per_rank": [round(row[4], 3) for row in stats_all.tolist()],
        "trace_ms_rank0_last_step": {k: round(v, 3) for k, v in sim.trace.items()} if sim.trace is not None else None,
        "trace_ms_rank0_step_totals": [round(sum(tr.values()), 2) for tr in getattr(sim, "trace_all", [])] or None,
        "trace_ms_rank0_forces": [round(tr.get("forces+update", 0), 2) for tr in getattr(sim, "trace_all", [])] or None,
        "trace_ms_rank0_last_migration_step": {k: round(v, 3) for k, v in getattr(sim, "trace_migration_step", {}).items()} or None,
        "wall_s_timed_loop": wall, "e2e": e2e, "roofline": roofline, "gpu_launches": launches_per_let_step(key_bits) * args.steps,
        "clocks": ck,
    }
    dist.destroy_process_group()
    return line if rank == 0 else None
