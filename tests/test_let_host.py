"""Host logic of the locally-essential-tree mode on CPU: octree-aligned domain cuts, the work-weighted splitter
election, box compaction, the export-walk oracle on a hand-made tree, and a world_size-2 gloo run of the
election + migration exchange (keys only — the engine itself has no CPU path)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_host_decisions_equal_the_python_statements(orc):
    """bh_let_domain_cuts / bh_let_elect_splitters (C++ in libbh.so) against their plain-Python statements."""
    from nbody_barnes_hut_cuda_b200.let import KEY_END, SAMPLE, domain_cuts, elect_splitters

    rng = np.random.default_rng(11)
    for t in range(3000):
        a, b = sorted(rng.integers(0, KEY_END + 1, 2).tolist())
        if t % 3 == 0:
            b = min(KEY_END, a + int(rng.integers(0, 70000)))
        if t % 7 == 0:
            a = a // 4096 * 4096
        assert (domain_cuts(a, b) == orc.let_domain_cuts_ref(a, b)).all(), (a, b)
    for a, b in ((0, KEY_END), (0, 0), (KEY_END, KEY_END), (5, 6), (0, 1), (KEY_END - 1, KEY_END), (1 << 27, 1 << 28)):
        assert (domain_cuts(a, b) == orc.let_domain_cuts_ref(a, b)).all(), (a, b)
    for t in range(60):
        world = int(rng.integers(1, 9))
        ns = SAMPLE if t % 2 == 0 else int(rng.integers(1, 50))
        hi = KEY_END if t % 3 else 1000                         # many duplicate keys in every third case
        samples = np.sort(rng.integers(0, hi, (world, ns)), 1)
        work = rng.uniform(0.5, 50.0, world) * rng.integers(0, 2, world) if t % 4 == 0 else rng.uniform(1.0, 1e9, world)
        assert (elect_splitters(samples, work) == orc.let_elect_splitters_ref(samples, work)).all(), t
    z = np.zeros((3, 8), np.int64)
    assert (elect_splitters(z, np.zeros(3)) == orc.let_elect_splitters_ref(z, np.zeros(3))).all()


def test_domain_cuts_are_octree_aligned():
    from nbody_barnes_hut_cuda_b200.let import KEY_END, MAX_BOXES, domain_cuts

    rng = np.random.default_rng(0)
    for t in range(400):
        a, b = sorted(rng.integers(0, KEY_END + 1, 2).tolist())
        if t % 3 == 0:
            b = min(KEY_END, a + int(rng.integers(0, 5000)))
        c = domain_cuts(a, b).astype(np.int64)
        assert len(c) == MAX_BOXES + 1 and c[0] == a and c[-1] == b and (np.diff(c) >= 0).all()
        u = np.unique(c)
        for lo_, hi_ in zip(u[:-1], u[1:]):                  # every interval lies inside ONE cell of its size class
            size = 1
            while size < hi_ - lo_:
                size *= 8
            assert lo_ // size == (hi_ - 1) // size
    full = np.unique(domain_cuts(0, KEY_END))
    assert len(full) == 9 and (np.diff(full.astype(np.int64)) == 1 << 27).all()   # the 8 octants
    assert len(np.unique(domain_cuts(77, 77))) == 1                               # empty range


def _work_spaced_sample(keys, work, m):
    cum = np.cumsum(work)
    return keys[np.minimum(np.searchsorted(cum, (np.arange(m) + 0.5) * cum[-1] / m), len(keys) - 1)]


def test_splitter_election_balances_weighted_work():
    from nbody_barnes_hut_cuda_b200.let import KEY_END, SAMPLE, elect_splitters

    rng = np.random.default_rng(3)
    world = 4
    allkeys = np.sort(rng.integers(0, KEY_END, 400_000))
    owned = np.array_split(allkeys, world)                    # current ownership: equal-count key ranges
    ones = [np.ones(len(o)) for o in owned]
    samples = np.stack([_work_spaced_sample(o, w, SAMPLE) for o, w in zip(owned, ones)])
    e = elect_splitters(samples, np.array([w.sum() for w in ones]))
    assert e[0] == 0 and e[-1] == KEY_END
    got = np.diff(np.searchsorted(allkeys, e))
    assert abs(got - len(allkeys) / world).max() < 0.01 * len(allkeys)            # unit work: equal counts
    # work that varies inside and between the ranks (rank 1 is 3x as expensive, with a hot spot)
    work = [np.ones(len(o)) for o in owned]
    work[1] *= 3.0
    work[1][1000:3000] *= 10.0
    samples = np.stack([_work_spaced_sample(o, w, SAMPLE) for o, w in zip(owned, work)])
    e2 = elect_splitters(samples, np.array([w.sum() for w in work]))
    allwork = np.concatenate(work)
    cum = np.concatenate([[0], np.cumsum(allwork)])
    per_rank = np.diff(cum[np.searchsorted(allkeys, e2)])
    assert abs(per_rank - allwork.sum() / world).max() < 0.01 * allwork.sum()      # equal work afterwards
    # ranks without bodies are ignored, edges stay monotone
    e3 = elect_splitters(samples, np.array([100.0, 0, 100, 100]))
    assert (np.diff(e3) >= 0).all() and e3[-1] == KEY_END
    assert (elect_splitters(samples, np.zeros(world)) == [0] + [KEY_END] * world).all()


def test_compact_boxes_keeps_used_boxes_in_front():
    from nbody_barnes_hut_cuda_b200.let import EMPTY_BOX, compact_boxes

    b = np.tile(EMPTY_BOX, (3, 10, 1))
    b[0, 4] = [0, 0, 0, 1, 1, 1]
    b[0, 7] = [2, 2, 2, 3, 3, 3]
    b[2, 9] = [5, 5, 5, 6, 6, 6]
    c = compact_boxes(b)
    assert c.shape == (3, 2, 6)
    assert (c[0, 0] == b[0, 4]).all() and (c[0, 1] == b[0, 7]).all() and (c[2, 0] == b[2, 9]).all()
    assert c[1, 0, 0] > c[1, 0, 3] and c[2, 1, 0] > c[2, 1, 3]


def test_export_oracle_on_a_two_level_tree(orc):
    """Root with one loose body and one child cell of two bodies.  A far box receives one point (the root), a
    box at the child's distance scale receives the loose body + the child's monopole, a box on top of the
    child receives the three bodies."""
    posm = np.array([[0, 0, 0, 1], [100, 100, 100, 2], [101, 100, 100, 4]], np.float32)
    meta = np.array([[0, 3, 0, -1], [1, 2, 6, 0]], np.int32)                      # root level 0, child level 6
    E = 0x7F7F7F7F
    child = np.array([[-2147483648 | 0, 1, E, E, E, E, E, E], [-2147483648 | 1, -2147483648 | 2, E, E, E, E, E, E]]).astype(np.int64).astype(np.int32)
    com = np.array([[(200 + 404) / 7.0, 600 / 7.0, 600 / 7.0, 7], [(200 + 404) / 6.0, 100, 100, 6]], np.float32)
    root_w = 1024.0                                                                 # child width 16

    def pts(box):
        p = orc.let_export_points(meta, child, com, posm, 0, np.array([box], np.float32), root_w)
        return sorted(map(tuple, p.tolist()))

    far = pts([1e5, 0, 0, 1e5 + 1, 1, 1])
    assert far == [tuple(com[0].tolist())]
    mid = pts([300, 100, 100, 301, 101, 101])                                       # 200 away: root opens, child (16 < 0.5*200) accepted
    assert mid == sorted([tuple(posm[0].tolist()), tuple(com[1].tolist())])
    near = pts([100, 100, 100, 101, 101, 101])
    assert near == sorted(map(tuple, posm.tolist()))
    assert len(orc.let_export_points(meta, child, com, posm, 0, np.array([[1, 1, 1, -1, -1, -1]], np.float32), root_w)) == 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from nbody_barnes_hut_cuda_b200.let import KEY_END, SAMPLE, elect_splitters

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = 5000 + 3000 * rank                                     # unequal ranks, overlapping key ranges
    keys = np.sort(rng.integers(0, KEY_END // (1 + rank), n))
    sample = keys[np.linspace(0, n - 1, SAMPLE).astype(int)]
    mine = torch.from_numpy(np.concatenate([sample.astype(np.float64), [float(n)]]))
    pooled = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(pooled, mine)
    pooled = torch.stack(pooled).numpy()
    edges = elect_splitters(pooled[:, :SAMPLE].astype(np.int64), pooled[:, SAMPLE])
    pos = np.searchsorted(keys, edges[1:-1])
    sc = np.diff(np.concatenate([[0], pos, [n]])).astype(np.int64)
    rc_t = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(rc_t, torch.from_numpy(sc))
    rc = rc_t.numpy()
    recv = torch.empty(int(rc.sum()), dtype=torch.int64)
    dist.all_to_all_single(recv, torch.from_numpy(keys.astype(np.int64)), output_split_sizes=rc.tolist(), input_split_sizes=sc.tolist())
    got = recv.numpy()
    q.put((rank, edges.tolist(), len(got), bool(((got >= edges[rank]) & (got < edges[rank + 1])).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_election_and_migration_over_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                              # both ranks elected the same edges
    assert res[0][2] + res[1][2] == 5000 + 8000                # no body lost or duplicated
    assert res[0][3] and res[1][3]                             # every received key lies in the receiver's range
    assert abs(res[0][2] - res[1][2]) < 0.03 * 13000           # equal counts at unit cost


@pytest.mark.parametrize("kind", ["uniform", "disk"])
def test_export_oracle_properties_on_oracle_trees(orc, kind):
    """The CPU statement of the export rule on trees the oracle itself builds: the emitted points always carry
    the whole mass, every emitted cell passes the acceptance test at the peer's nearest box, a peer far away
    receives the root alone, and a peer box on top of the bodies receives (almost) every body individually."""
    import nbody_barnes_hut_cuda_b200 as bh

    n = 1500
    soa = bh.ic_uniform_cube(n, 3, 1000.0) if kind == "uniform" else bh.ic_refdisk(n, 3)
    posm, _, _ = orc.soa_to_internal(soa)
    b = orc.bounds(*soa[:3])
    keys, idx = orc.morton_keys(*soa[:3], b)
    ks, perm = orc.stable_sort(keys, idx)
    ps = np.ascontiguousarray(posm[perm])
    meta, child, root = orc.tree_build(ks)
    _, com = orc.tree_com(ps, meta, child, root)
    meta_l = meta.copy()
    meta_l[:, 2] &= 0x1FF
    root_w = float(b[3] - b[0])
    total = ps[:, 3].astype(np.float64).sum()
    lo, hi = ps[:, :3].min(0), ps[:, :3].max(0)
    ext = hi - lo
    cells = {tuple(c.tolist()): int(m[2]) & 0xFF for c, m in zip(com, meta_l)}
    bodies = set(map(tuple, ps.tolist()))
    for name, boxes in (("far", [np.concatenate([hi + 60 * ext, hi + 61 * ext])]),
                        ("near", [np.concatenate([hi + 0.05 * ext, hi + 0.4 * ext]), np.concatenate([lo - 0.4 * ext, lo - 0.05 * ext])]),
                        ("inside", [np.concatenate([lo, hi])])):
        pts = orc.let_export_points(meta_l, child, com, ps, root, np.array(boxes, np.float32), root_w)
        assert abs(pts[:, 3].astype(np.float64).sum() - total) < 1e-5 * total, name
        if name == "far":
            assert len(pts) == 1 and tuple(pts[0].tolist()) == tuple(com[root].tolist())
        n_cells = 0
        for p in pts:
            t = tuple(p.tolist())
            if t in bodies:
                continue
            assert t in cells, name                        # everything else is a cell's {com, mass}
            n_cells += 1
            c = np.array(boxes, np.float64)
            d = np.maximum(0.0, np.maximum(c[:, :3] - p[:3], p[:3] - c[:, 3:]))
            d2 = (d * d).sum(1).min()
            w = root_w / 2.0 ** cells[t]
            assert w * w < 0.25 * (d2 + 50.0) * (1 + 1e-5), name
        if name == "inside":
            # a box over everything: only cells narrower than theta*sqrt(SOFTENING) = 3.5 survive as monopoles
            assert all(root_w / 2.0 ** cells[tuple(p.tolist())] < 3.54 for p in pts if tuple(p.tolist()) not in bodies)
            assert len(pts) > 0.9 * n


# ---- the whole LET protocol between two ranks over gloo, the oracle standing in for the GPU engine -------------------
def _let_worker(rank, world, port, n, q):
    """Each rank: start-up share by index, cube all-gather, election + migration (keys and bodies), oracle tree of the
    own bodies, octree-aligned boxes, export walk per peer (oracle_lib.let_export_points), all-to-all of the points,
    forces on the own bodies = direct sum over own bodies + imported points."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import nbody_barnes_hut_cuda_b200 as bh
    import oracle_lib as O
    from nbody_barnes_hut_cuda_b200.let import SAMPLE, domain_cuts, elect_splitters, global_cube

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    soa = bh.ic_refdisk(n, 21)
    posm_all, _, _ = O.soa_to_internal(soa)
    first, last = n * rank // world, n * (rank + 1) // world
    own, ids = posm_all[first:last].copy(), np.arange(first, last)

    def gather(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return torch.stack(out).numpy()

    def exchange(rows, counts):                                # rows [sum(counts), k] float64, split by destination
        sc = torch.tensor(counts, dtype=torch.int64)
        rc = torch.empty(world, dtype=torch.int64)
        dist.all_to_all_single(rc, sc)
        recv = torch.empty((int(rc.sum()), rows.shape[1]), dtype=torch.float64)
        dist.all_to_all_single(recv, torch.from_numpy(np.ascontiguousarray(rows, np.float64)), output_split_sizes=rc.tolist(),
                               input_split_sizes=sc.tolist())
        return recv.numpy()

    # cube
    box = np.concatenate([own[:, :3].min(0), own[:, :3].max(0)]).astype(np.float32)
    cube = global_cube(gather(torch.from_numpy(box)))
    # election (unit work) + migration
    keys, _ = O.morton_keys(own[:, 0].copy(), own[:, 1].copy(), own[:, 2].copy(), cube)
    order = np.argsort(keys, kind="stable")
    keys, own, ids = keys[order].astype(np.int64), own[order], ids[order]
    sample = keys[(np.arange(SAMPLE) * (len(keys) - 1)) // (SAMPLE - 1)]
    pooled = gather(torch.from_numpy(np.concatenate([sample.astype(np.float64), [float(len(keys))]])))
    edges = elect_splitters(pooled[:, :SAMPLE].astype(np.int64), pooled[:, SAMPLE])
    cut = np.searchsorted(keys, edges[1:-1])
    counts = np.diff(np.concatenate([[0], cut, [len(keys)]])).tolist()
    got = exchange(np.concatenate([own.astype(np.float64), ids[:, None].astype(np.float64)], 1), counts)
    own, ids = got[:, :4].astype(np.float32), got[:, 4].astype(np.int64)
    # own tree
    keys, _ = O.morton_keys(own[:, 0].copy(), own[:, 1].copy(), own[:, 2].copy(), cube)
    assert ((keys >= edges[rank]) & (keys < edges[rank + 1])).all()
    order = np.argsort(keys, kind="stable")
    ks, ps, ids = keys[order], np.ascontiguousarray(own[order]), ids[order]
    meta, child, root = O.tree_build(ks)
    _, com = O.tree_com(ps, meta, child, root)
    meta[:, 2] &= 0x1FF
    # domain boxes of the octree-aligned intervals
    cuts = domain_cuts(int(edges[rank]), int(edges[rank + 1])).astype(np.int64)
    idx = np.searchsorted(ks.astype(np.int64), cuts)
    mine = np.tile(np.array([1, 1, 1, -1, -1, -1], np.float32), (len(cuts) - 1, 1))
    for k in range(len(cuts) - 1):
        if idx[k + 1] > idx[k]:
            run = ps[idx[k]:idx[k + 1], :3]
            mine[k] = np.concatenate([run.min(0), run.max(0)])
    boxes = gather(torch.from_numpy(mine))
    # export walk per peer, exchange, forces
    root_w = float(cube[3] - cube[0])
    lists = [O.let_export_points(meta, child, com, ps, root, boxes[p], root_w) if p != rank else np.zeros((0, 4), np.float32)
             for p in range(world)]
    for p in range(world):
        if p != rank:
            assert abs(lists[p][:, 3].astype(np.float64).sum() - ps[:, 3].astype(np.float64).sum()) < 1e-5 * ps[:, 3].sum()
    imports = exchange(np.concatenate(lists).astype(np.float64), [len(x) for x in lists]).astype(np.float32)
    union = np.ascontiguousarray(np.concatenate([ps, imports]))
    acc = O.direct_sum(union, np.arange(len(ps), dtype=np.int32))
    exact = O.direct_sum(posm_all, ids.astype(np.int32))
    q.put((rank, len(ps), len(imports), O.rel_rms(acc, exact), sorted(ids.tolist())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_let_protocol_over_gloo():
    import torch.multiprocessing as mp

    n = 3000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_let_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == n and sorted(res[0][4] + res[1][4]) == list(range(n))   # every body owned once
    for _, n_own, n_imp, err, _ in res:
        assert 0 < n_imp < n - n_own          # the peer was summarised, not copied
        assert err < 3e-3                     # forces from own bodies + imported points: theta = 0.5 multipole error class
