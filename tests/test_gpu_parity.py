"""GPU parity: every phase of the CUDA path against the CPU oracle on the same seeded inputs,
called through the C ABI (include/bh.h).  Bit-exact for bounds / keys / permutation / tree
topology / integrator; accelerations within 1e-4 relative RMS (north_star) — in practice ~1e-6."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
f = np.float32


def make_case(bh, kind, n, seed=1):
    rng = np.random.default_rng(seed)
    z = np.zeros(n, f)
    if kind == "uniform":
        return bh.ic_uniform_cube(n, seed, 1000.0)
    if kind == "disk":
        return bh.ic_refdisk(n, 42)
    if kind == "plummer":
        return bh.ic_plummer(n, seed, 200.0, 10.0, 4.5, 0.5)
    if kind == "clustered":
        pos = rng.normal(0, 5, (n, 3)) + rng.integers(0, 4, (n, 1)) * 700.0
        return [pos[:, 0].astype(f), pos[:, 1].astype(f), pos[:, 2].astype(f), z.copy(), z.copy(), z.copy(),
                rng.uniform(2, 7, n).astype(f)]
    if kind == "coincident":   # every body at one point: a single level-10 bucket is the root
        return [np.full(n, 3.5, f), np.full(n, -2.0, f), np.full(n, 9.0, f), z.copy(), z.copy(), z.copy(),
                rng.uniform(2, 7, n).astype(f)]
    if kind == "lattice":      # few distinct positions -> many identical keys (buckets) + loose bodies
        pos = rng.integers(0, 6, (n, 3)).astype(f) * 100.0
        pos[: n // 4] += rng.uniform(-40, 40, (n // 4, 3)).astype(f)
        return [pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), z.copy(), z.copy(), z.copy(),
                rng.uniform(2, 7, n).astype(f)]
    if kind == "line":         # bodies on a line: a chain-like tree, the cell count approaches n-1 (worst-case pools)
        x = np.sort(rng.uniform(-900, 900, n)).astype(f)
        return [x, z.copy(), z.copy(), z.copy(), z.copy(), z.copy(), rng.uniform(2, 7, n).astype(f)]
    if kind == "bigbucket":    # one very large identical-key bucket next to an ordinary cloud
        pos = rng.uniform(-500, 500, (n, 3)).astype(f)
        pos[: n // 2] = np.array([123.0, -45.0, 67.0], f)
        return [pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), z.copy(), z.copy(), z.copy(),
                rng.uniform(2, 7, n).astype(f)]
    if kind == "tracers":      # massless bodies feel forces and exert none
        soa = bh.ic_uniform_cube(n, seed, 800.0)
        soa[6][::3] = 0.0
        return soa
    raise ValueError(kind)


CASES = [("uniform", 2), ("uniform", 3), ("uniform", 31), ("uniform", 33), ("uniform", 1000), ("uniform", 16384),
         ("disk", 50001), ("clustered", 20000), ("coincident", 500), ("lattice", 4096), ("plummer", 30000),
         ("line", 20000), ("bigbucket", 12000), ("tracers", 9000),
         # the tree build answers range questions per tile of 2,048 key pairs: sizes around the tile edge, buckets and
         # chains that cross it
         ("uniform", 2048), ("uniform", 2049), ("uniform", 2050), ("coincident", 5000), ("lattice", 4097), ("line", 4099)]


@pytest.mark.parametrize("kind,n", CASES)
def test_every_phase_against_oracle(bh, kind, n):
    check_every_phase(bh, make_case(bh, kind, n))


def test_headline_workload_at_full_size_against_oracle(bh):
    """BASELINE.json configs[1] at its own size: the 1,000,000-body reference disk (bench:294-308), every phase
    against the oracle — ids, keys, tree and centre of mass bit for bit, acceptance decisions equal, accelerations
    within 1e-4 relative RMS."""
    check_every_phase(bh, bh.ic_refdisk(1_000_000, 42))


@pytest.mark.parametrize("theta", [0.3, 0.8])
def test_plummer_theta_sweep_decisions_equal_the_oracle(bh, theta):
    """BASELINE.json configs[2]: Plummer sphere at theta != 0.5 (bench:207-208 with another THETA)."""
    check_every_phase(bh, make_case(bh, "plummer", 100_000), theta=theta)


def check_every_phase(bh, soa, theta=O.THETA):
    n = len(soa[0])
    posm, vel, ids = O.soa_to_internal(soa)
    P, D, S = bh.PHASE, bh.DBG, bh.STAT
    with bh.BHEngine(n, flags=1, theta=theta) as eng:
        eng.load_soa(*soa)
        # --- bounds + keys: bit exact (bench:134-156, 42-63)
        eng.run_phase(P.KEYS)
        b = O.bounds(*soa[:3])
        assert eng.debug_get(D.BOUNDS).tobytes() == b.tobytes()
        keys, idx = O.morton_keys(*soa[:3], b)
        assert (eng.debug_get(D.KEYS) == keys).all()
        # --- sort: stable permutation, bit exact (bench:262-264)
        eng.run_phase(P.SORT)
        ks, perm = O.stable_sort(keys, idx)
        assert (eng.debug_get(D.KEYS) == ks).all()
        assert (eng.debug_get(D.PERM) == perm).all()
        ps = np.ascontiguousarray(posm[perm])
        assert eng.debug_get(D.POSM_SORTED).tobytes() == ps.tobytes()
        assert (eng.debug_get(D.IDS_SORTED) == perm).all()
        # --- tree topology: identical arrays (numbering is canonical) and identical tuple sets
        eng.run_phase(P.BUILD)
        meta, child, root = O.tree_build(ks)
        assert eng.stat(S.DEVICE_ERROR) == 0
        assert eng.stat(S.CELLS) == len(meta) and eng.stat(S.ROOT) == root
        gmeta, gchild = eng.debug_get(D.CELL_META), eng.debug_get(D.CELL_CHILD)
        assert O.cell_tuples(gmeta, ks) == O.cell_tuples(meta, ks)
        assert (gmeta == meta).all()
        internal = ((meta[:, 2] >> 8) & 1) == 0
        assert (gchild[internal] == child[internal]).all()
        # --- centre of mass: same slot-ordered float sums -> bit exact
        eng.run_phase(P.COM)
        mom, com = O.tree_com(ps, meta, child, root)
        gcom = eng.debug_get(D.CELL_COM)
        assert np.allclose(gcom, com, rtol=1e-6, atol=1e-6)
        assert gcom.tobytes() == com.tobytes()
        # --- force: same decisions (interaction counts equal), accelerations to rounding
        eng.run_phase(P.FORCE)
        assert eng.stat(S.DEVICE_ERROR) == 0
        groups = O.make_groups(ps, ks, O.GROUP, O.SPLIT)        # same cut rule as the warp applies
        acc, counts = O.force_groups(ps, b, meta, child, com, root, groups, theta=theta)
        gacc = eng.debug_get(D.ACC)
        assert eng.stat(S.INTERACTIONS_CELL) == counts[0]
        assert eng.stat(S.INTERACTIONS_BODY) == counts[1]
        scale = np.abs(acc[:, :3]).max() if n > 1 else 1.0
        if scale > 0:
            assert O.rel_rms(gacc[:, :3], acc[:, :3]) < 1e-4   # north_star tolerance
            assert np.abs(gacc[:, :3] - acc[:, :3]).max() < 1e-4 * scale
        # --- kick-drift-clamp on the GPU's own accelerations: bit exact (bench:227-249)
        eng.run_phase(P.UPDATE)
        vs = vel[perm]
        want = O.integrate(ps[:, 0], ps[:, 1], ps[:, 2], vs[:, 0], vs[:, 1], vs[:, 2], gacc[:, 0], gacc[:, 1], gacc[:, 2])
        gp, gv = eng.debug_get(D.POSM), eng.debug_get(D.VEL)
        for a in range(3):
            assert gp[:, a].tobytes() == want[a].tobytes()
            assert gv[:, a].tobytes() == want[3 + a].tobytes()
        assert gp[:, 3].tobytes() == ps[:, 3].tobytes() and (eng.debug_get(D.IDS) == perm).all()


def test_speed_clamp_on_device(bh):
    n = 64
    soa = make_case(bh, "uniform", n)
    soa[3][:] = 400.0
    soa[4][:] = 400.0   # |v| = 565 > MAX_SPEED
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(1)
        out = eng.read_soa()
        sp = np.sqrt(out[3].astype(np.float64) ** 2 + out[4] ** 2 + out[5] ** 2)
        assert np.abs(sp - 500.0).max() < 1e-3


@pytest.mark.parametrize("kind,n,steps", [("uniform", 16384, 10), ("disk", 20000, 10), ("lattice", 3000, 5)])
def test_multi_step_against_oracle_and_graph_equals_direct(bh, kind, n, steps):
    """BASELINE.json configs[0]: 16,384-body uniform cube, 10 leapfrog steps."""
    soa = make_case(bh, kind, n)
    posm, vel, ids = O.soa_to_internal(soa)
    want = O.engine_step(posm, vel, ids, steps)
    outs = []
    for flags in (0, 1):   # CUDA-graph replay and direct launches must agree bit for bit
        with bh.BHEngine(n, flags=flags) as eng:
            eng.load_soa(*soa)
            eng.simulation_step(steps)
            eng.check_device_error()
            outs.append((eng.debug_get(bh.DBG.POSM), eng.debug_get(bh.DBG.VEL), eng.debug_get(bh.DBG.IDS),
                         eng.read_soa()))
    assert outs[0][0].tobytes() == outs[1][0].tobytes() and outs[0][1].tobytes() == outs[1][1].tobytes()
    gp, gv, gid, soa_out = outs[0]
    # same bodies in the same Morton slots
    assert (gid == want["ids"]).all()
    assert O.rel_rms(gp[:, :3], want["posm"][:, :3]) < 1e-6
    assert O.rel_rms(gv[:, :3], want["vel"][:, :3]) < 1e-4 if np.abs(want["vel"]).max() > 0 else True
    # export is in ORIGINAL body order (bench:31-40: slot i is body i)
    back = np.zeros((n, 3), f)
    back[gid] = gp[:, :3]
    assert back[:, 0].tobytes() == soa_out[0].tobytes() and back[:, 2].tobytes() == soa_out[2].tobytes()


def test_step_host_equals_load_step_read(bh):
    n = 10000
    soa = make_case(bh, "disk", n)
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(3)
        a = eng.read_soa(want_acc=False)
    with bh.BHEngine(n) as eng:
        arrs = [x.copy() for x in soa]
        eng.step_host(*arrs, nsteps=3)
        for k in range(6):
            assert arrs[k].tobytes() == a[k].tobytes()


def test_repeated_host_steps_equal_one_run(bh):
    """Call after call on the same system (bh_step_host keeps the chunk-cost record between calls: a scheduling hint,
    it must never change a result) == the same number of steps in one go; a call with another n in between drops it."""
    n = 20000
    soa = make_case(bh, "uniform", n)   # no two bodies share a key: the order within the sort does not depend on the import order
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(4)
        want = eng.read_soa(want_acc=False)
    with bh.BHEngine(n) as eng:
        arrs = [x.copy() for x in soa]
        eng.step_host(*arrs, nsteps=1)
        eng.step_host(*arrs, nsteps=1)
        small = [x[:777].copy() for x in make_case(bh, "uniform", 1000)]
        eng.step_host(*small, nsteps=1)              # a different system through the same context
        eng.step_host(*arrs, nsteps=2)
        eng.check_device_error()
        for k in range(6):
            assert arrs[k].tobytes() == want[k].tobytes()


def test_update_fused_into_the_traversal_equals_the_separate_kernel(bh, monkeypatch):
    """bh_step's whole-step path integrates a chunk inside force_kernel; BH_NO_FUSED_UPDATE=1 (read at creation) keeps
    integrate_kernel as a launch of its own.  Same arithmetic: states and the next step's cube must agree bit for bit."""
    n = 30000
    soa = make_case(bh, "plummer", n)
    outs = []
    for env in ("0", "1"):
        monkeypatch.setenv("BH_NO_FUSED_UPDATE", env)
        with bh.BHEngine(n) as eng:
            eng.load_soa(*soa)
            eng.simulation_step(5)
            eng.check_device_error()
            outs.append((eng.debug_get(bh.DBG.POSM).tobytes(), eng.debug_get(bh.DBG.VEL).tobytes(),
                         eng.debug_get(bh.DBG.IDS).tobytes(), eng.debug_get(bh.DBG.ACC).tobytes(),
                         eng.debug_get(bh.DBG.BOUNDS).tobytes()))
    assert outs[0] == outs[1]


def test_phase_timer_reports_every_phase(bh):
    n = 100000
    soa = make_case(bh, "disk", n)
    with bh.BHEngine(n, flags=2) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(3)
        ms = eng.phase_ms()
        assert all(ms[k] > 0 for k in ("keys", "sort", "build", "com", "force", "update"))
        assert abs(sum(ms[k] for k in ("keys", "sort", "build", "com", "force", "update")) - ms["total"]) < 1e-3


def test_accuracy_against_direct_sum_on_device(bh):
    """north_star: within the theta=0.5 multipole error of an O(N^2) direct sum, monotone in theta."""
    n = 16384
    soa = make_case(bh, "uniform", n, seed=42)
    sample = np.arange(0, n, 8, dtype=np.int32)
    errs = []
    for theta in (0.3, 0.5, 0.8):
        with bh.BHEngine(n, theta=theta) as eng:
            eng.load_soa(*soa)
            eng.simulation_step(1)
            out = eng.read_soa()
            acc = np.stack(out[6:9], 1)[sample]
            ref = eng.direct_sample(sample)
            errs.append(O.rel_rms(acc, ref))
            if theta == 0.5:
                posm, _, _ = O.soa_to_internal(soa)
                cpu = O.direct_sum(posm, sample)
                assert O.rel_rms(ref, cpu) < 1e-10      # the GPU direct sum itself is right
    assert errs[0] < errs[1] < errs[2]
    assert errs[1] < 3e-3                                # 2.8e-3 with the per-body test


def test_energy_on_device_matches_oracle(bh):
    n = 3000
    soa = make_case(bh, "disk", n)
    posm, vel, ids = O.soa_to_internal(soa)
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        ke, pe = eng.energy()
    wke, wpe = O.energy(posm, vel)
    assert abs(ke - wke) / wke < 1e-10 and abs(pe - wpe) / abs(wpe) < 1e-10


# measured on B200 with the group acceptance test: 1.14e-2 on this sample (profiles/r02_quadrupole_1m.json, monopole,
# theta 0.5; the per-body test of SURVEY §6 gives 1.3e-2 on this thin disk).  measured x 1.2 would be 1.37e-2: the
# round-1 bound already is tighter than that and stays
MILLION_BODY_DIRECT_SUM_BOUND = 1.3e-2


def test_million_body_invariants(bh):
    """BASELINE.json configs[1] size: properties that do not need the oracle at full size."""
    n = 1_000_000
    soa = bh.ic_refdisk(n, 42)
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(2)
        eng.check_device_error()
        keys = eng.debug_get(bh.DBG.KEYS)
        assert (keys[1:] >= keys[:-1]).all()                       # sortedness
        ids = eng.debug_get(bh.DBG.IDS)
        assert (np.sort(ids) == np.arange(n)).all()                # a permutation of the bodies
        meta = eng.debug_get(bh.DBG.CELL_META)
        root = eng.stat(bh.STAT.ROOT)
        assert meta[root, 0] == 0 and meta[root, 1] == n
        com = eng.debug_get(bh.DBG.CELL_COM)
        assert abs(com[root, 3] - soa[6].astype(np.float64).sum()) / soa[6].sum() < 1e-5   # mass conservation
        per_body = (eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY)) / n
        assert 600 < per_body < 2500                               # oracle at this input: ~1180
        # accuracy on a sample against the on-device double direct sum
        sample = np.arange(0, n, 997, dtype=np.int32)
        out = eng.read_soa()
        err = O.rel_rms(np.stack(out[6:9], 1)[sample], eng.direct_sample(sample))
        print(f"refdisk 1M: rel-RMS vs direct sum on {len(sample)} bodies = {err:.3e}, interactions/body = {per_body:.1f}")
        assert err < MILLION_BODY_DIRECT_SUM_BOUND


def test_two_morton_slices_emulated_on_one_gpu_equal_the_full_run(bh):
    """Multi-GPU path with fewer GPUs than ranks: two contexts on one device play ranks 0 and 1 of 2;
    the all-gather is emulated by device copies.  Must equal the single-context run bit for bit."""
    import torch

    from nbody_barnes_hut_cuda_b200.sliced import device_views, slice_bounds

    n, steps = 50001, 3
    soa = make_case(bh, "disk", n)
    with bh.BHEngine(n) as ref:
        ref.load_soa(*soa)
        ref.simulation_step(steps)
        want = (ref.debug_get(bh.DBG.POSM), ref.debug_get(bh.DBG.VEL), ref.debug_get(bh.DBG.IDS))
        want_inter = ref.stat(bh.STAT.INTERACTIONS_CELL) + ref.stat(bh.STAT.INTERACTIONS_BODY)
    dev = torch.device("cuda:0")
    engs, views, bounds = [], [], []
    for r in range(2):
        e = bh.BHEngine(n)
        e.set_slice(r, 2)
        e.load_soa(*soa)
        st = e.state_ptrs()
        first, count, per = slice_bounds(n, r, 2)
        assert (st["first"], st["count"]) == (first, count)
        engs.append(e)
        bounds.append((first, count, per))
        views.append(device_views(torch, st, per, 2, dev))
    inter = 0
    for s in range(steps):
        for e in engs:
            e.simulation_step(1)
        torch.cuda.synchronize()
        for r in range(2):
            first, count, per = bounds[r]
            for mine, other in zip(views[r], views[1 - r]):
                other[first:first + count] = mine[first:first + count]
        torch.cuda.synchronize()
    inter = sum(e.stat(bh.STAT.INTERACTIONS_CELL) + e.stat(bh.STAT.INTERACTIONS_BODY) for e in engs)
    for e in engs:
        e.check_device_error()
        got = (e.debug_get(bh.DBG.POSM), e.debug_get(bh.DBG.VEL), e.debug_get(bh.DBG.IDS))
        assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes()
        assert (got[2] == want[2]).all()
        e.close()
    assert inter == want_inter
