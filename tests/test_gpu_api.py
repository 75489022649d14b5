"""Behaviour of the C ABI itself: call order, error codes, re-use of a context, degenerate sizes."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
f = np.float32


def test_call_order_and_argument_errors(bh):
    L = bh.lib()
    eng = bh.BHEngine(1000)
    assert L.bh_step(eng._ctx, 1, None) == -3              # BH_E_STATE: nothing imported yet
    assert L.bh_export_soa(eng._ctx, *([None] * 9), None) == -3
    soa = bh.ic_uniform_cube(2000, 1, 100.0)
    with pytest.raises(bh.BHError):
        eng.load_soa(*soa)                                 # n > n_max
    empty = [np.zeros(0, f) for _ in range(7)]
    assert L.bh_import_soa_host(eng._ctx, *[x.ctypes.data_as(C.c_void_p) for x in empty], 0) == -1   # empty input
    assert L.bh_run_phase(eng._ctx, 99, None) == -1
    assert L.bh_debug_get(eng._ctx, 0, None, 24) == -1
    eng.close()
    with pytest.raises(bh.BHError):
        bh.BHEngine(100, softening=0.0)                    # softening must be positive (self term)
    with pytest.raises(bh.BHError):
        bh.BHEngine(100, leaf_cap=8)                       # declared but not implemented: refused loudly


@pytest.mark.parametrize("n", [1, 2])
def test_tiny_systems(bh, n):
    soa = [np.array([1.0, 4.0][:n], f), np.array([2.0, 6.0][:n], f), np.array([3.0, 3.0][:n], f),
           np.array([0.5, 0.0][:n], f), np.zeros(n, f), np.zeros(n, f), np.array([2.0, 3.0][:n], f)]
    with bh.BHEngine(8) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(1)
        eng.check_device_error()
        out = eng.read_soa()
    if n == 1:   # a lone body feels nothing and drifts: x += v*dt (bench:246)
        assert out[6][0] == 0 and out[7][0] == 0 and out[0][0] == f(np.float64(f(0.5)) * np.float64(f(0.02)) + 1.0)
    else:        # two bodies: the hand formula of bench:205-213
        d = np.array([3.0, 4.0, 0.0])
        want0 = 0.5 * 3.0 * d / (25.0 + 50.0) ** 1.5
        assert np.allclose([out[6][0], out[7][0], out[8][0]], want0, rtol=1e-5)
        assert np.allclose([out[6][1], out[7][1], out[8][1]], -0.5 * 2.0 * d / 75.0 ** 1.5, rtol=1e-5)


def test_context_reuse_with_different_sizes_and_zero_steps(bh):
    with bh.BHEngine(40000) as eng:
        for n in (40000, 1234, 40000):
            soa = bh.ic_refdisk(n, 42)
            posm, vel, ids = O.soa_to_internal(soa)
            want = O.engine_step(posm, vel, ids, 2)
            eng.load_soa(*soa)
            eng.simulation_step(0)
            eng.simulation_step(2)                         # graph is re-captured when n changes
            eng.check_device_error()
            assert eng.n == n and eng.stat(bh.STAT.STEPS) == 2
            assert (eng.debug_get(bh.DBG.IDS) == want["ids"]).all()
            assert O.rel_rms(eng.debug_get(bh.DBG.POSM)[:, :3], want["posm"][:, :3]) < 1e-6


def test_two_contexts_are_independent(bh):
    a_soa, b_soa = bh.ic_uniform_cube(5000, 1, 1000.0), bh.ic_refdisk(7000, 42)
    with bh.BHEngine(5000) as a, bh.BHEngine(7000) as b:
        a.load_soa(*a_soa)
        b.load_soa(*b_soa)
        a.simulation_step(2)
        b.simulation_step(2)
        ra, rb = a.read_soa(), b.read_soa()
    with bh.BHEngine(5000) as a2:
        a2.load_soa(*a_soa)
        a2.simulation_step(2)
        ra2 = a2.read_soa()
    for x, y in zip(ra, ra2):
        assert x.tobytes() == y.tobytes()
    assert len(rb[0]) == 7000


def test_device_pointer_import_export(bh):
    """bh_import_soa / bh_export_soa with caller-owned DEVICE arrays (the reference's own layout, bench:32-35)."""
    import torch

    n = 8000
    soa = bh.ic_refdisk(n, 42)
    dev = [torch.from_numpy(x).cuda() for x in soa]
    out = [torch.zeros(n, dtype=torch.float32, device="cuda") for _ in range(9)]
    stream = torch.cuda.current_stream().cuda_stream
    with bh.BHEngine(n) as eng:
        eng.load_soa_device(dev, n, stream)
        eng.simulation_step(3, stream)
        bh.lib().bh_export_soa(eng._ctx, *[C.c_void_p(t.data_ptr()) for t in out], C.c_void_p(stream))
        torch.cuda.synchronize()
        host = eng.read_soa()
    for t, h in zip(out, host):
        assert t.cpu().numpy().tobytes() == h.tobytes()


def test_theta_zero_is_the_direct_sum(bh):
    """theta = 0 rejects every cell: the traversal degenerates to all pairs (via buckets and loose bodies)."""
    n = 700
    soa = bh.ic_uniform_cube(n, 4, 300.0)
    posm, _, _ = O.soa_to_internal(soa)
    want = O.direct_sum(posm, np.arange(n))
    with bh.BHEngine(n, theta=0.0) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(1)
        out = eng.read_soa()
        assert eng.stat(bh.STAT.INTERACTIONS_CELL) == 0 and eng.stat(bh.STAT.INTERACTIONS_BODY) == n * n
    assert O.rel_rms(np.stack(out[6:9], 1), want) < 2e-6


def test_step_halves_equal_a_full_step(bh):
    """bh_step_half(0) + bh_step_half(1) == bh_step(1), with and without graphs."""
    n = 30000
    soa = bh.ic_refdisk(n, 42)
    with bh.BHEngine(n) as ref:
        ref.load_soa(*soa)
        ref.simulation_step(3)
        want = (ref.debug_get(bh.DBG.POSM), ref.debug_get(bh.DBG.VEL), ref.debug_get(bh.DBG.IDS))
    for flags in (0, 1):
        with bh.BHEngine(n, flags=flags) as eng:
            eng.load_soa(*soa)
            for _ in range(3):
                eng.step_half(0)
                eng.step_half(1)
            eng.check_device_error()
            assert eng.stat(bh.STAT.STEPS) == 3
            got = (eng.debug_get(bh.DBG.POSM), eng.debug_get(bh.DBG.VEL), eng.debug_get(bh.DBG.IDS))
            assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes()
            assert (got[2] == want[2]).all()


def test_step_parts_and_host_step_equal_a_full_step(bh):
    """bh_step_part(0) + (1) + (2) == bh_step(1), with and without graphs, also for a state with ghosts (where part 1
    moves the ids too); bh_step_host — which uploads in order of need and runs the parts as the data lands — gives the
    same bits as load + step + read."""
    import torch

    n = 30000
    soa = bh.ic_refdisk(n, 42)
    with bh.BHEngine(n) as ref:
        ref.load_soa(*soa)
        ref.simulation_step(3)
        want = (ref.debug_get(bh.DBG.POSM), ref.debug_get(bh.DBG.VEL), ref.debug_get(bh.DBG.IDS))
        want_soa = ref.read_soa(want_acc=False)
    for flags in (0, 1):
        with bh.BHEngine(n, flags=flags) as eng:
            eng.load_soa(*soa)
            for _ in range(3):
                for part in range(3):
                    eng.step_part(part)
            eng.check_device_error()
            assert eng.stat(bh.STAT.STEPS) == 3
            got = (eng.debug_get(bh.DBG.POSM), eng.debug_get(bh.DBG.VEL), eng.debug_get(bh.DBG.IDS))
            assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes() and (got[2] == want[2]).all()
    # a state that came through bh_import_state (ghosts possible): same physics when every id is >= 0
    posm = torch.from_numpy(np.stack([soa[0], soa[1], soa[2], soa[6]], 1)).cuda()
    vel = torch.from_numpy(np.stack([soa[3], soa[4], soa[5], np.zeros(n, np.float32)], 1)).cuda()
    ids = torch.arange(n, dtype=torch.int32, device="cuda")
    with bh.BHEngine(n) as eng:
        eng.import_state(posm, vel, ids, n)
        for _ in range(3):
            for part in range(3):
                eng.step_part(part)
        got = (eng.debug_get(bh.DBG.POSM), eng.debug_get(bh.DBG.VEL), eng.debug_get(bh.DBG.IDS))
        assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes() and (got[2] == want[2]).all()
    # host step: 3 steps inside one call
    host = [a.copy() for a in soa]
    with bh.BHEngine(n) as eng:
        eng.step_host(*host, nsteps=3)
        for k in range(6):
            assert host[k].tobytes() == want_soa[k].tobytes()
        eng.step_host(*[a.copy() for a in soa], nsteps=1)     # a second call on the same context (graphs reused)
        eng.check_device_error()


def test_soa_and_visual_exports_skip_slots_without_a_local_id(bh):
    """contexts filled through bh_import_state may hold ghosts (id -1) and global ids >= n; the SoA and
    vertex-buffer exports index the caller's n-element arrays by id and must leave those slots alone."""
    import torch

    n = 4096
    soa = bh.ic_uniform_cube(n, 5, 500.0)
    posm = torch.from_numpy(np.stack([soa[0], soa[1], soa[2], soa[6]], 1)).cuda()
    vel = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    ids_np = np.arange(n, dtype=np.int32)
    ids_np[::7] = -1                       # ghosts
    ids_np[3::7] = 1_000_000_000           # ids of another rank's numbering
    ids = torch.from_numpy(ids_np).cuda()
    with bh.BHEngine(n) as eng:
        eng.import_state(posm, vel, ids, n)
        pad = 1024
        buf = torch.full((9, n + 2 * pad), -777.0, dtype=torch.float32, device="cuda")
        ptrs = [buf[k, pad:].data_ptr() for k in range(9)]
        L = bh.lib()
        import ctypes as C

        assert L.bh_export_soa(eng._ctx, *[C.c_void_p(p) for p in ptrs], None) == 0
        vbo = torch.full((2, 3 * n + 2 * pad), -777.0, dtype=torch.float32, device="cuda")
        eng.export_visuals(vbo[0, pad:].data_ptr(), vbo[1, pad:].data_ptr())
        torch.cuda.synchronize()
        h = buf.cpu().numpy()
        assert (h[:, :pad] == -777.0).all() and (h[:, pad + n:] == -777.0).all()      # nothing written outside
        local = (ids_np >= 0) & (ids_np < n)
        assert (h[0, pad:pad + n][local] == soa[0][local]).all()
        assert (h[0, pad:pad + n][~local] == -777.0).all()                            # skipped, not scattered somewhere
        v = vbo.cpu().numpy()
        assert (v[:, :pad] == -777.0).all() and (v[:, pad + 3 * n:] == -777.0).all()


def test_cpp_multi_gpu_driver_with_one_rank_equals_bh_step(bh):
    """bh_mg_* (csrc/bh_mg.cu) on a world of ONE: NCCL is dlopen'ed, a communicator is made from a unique id, the step
    loop runs its three-part schedule with the gather skipped — and must reproduce bh_step bit for bit.  The 2/4/8-GPU
    equivalence is tools/check_sliced.py under torchrun; this keeps the driver inside the single-GPU suite."""
    n = 50_001
    soa = bh.ic_refdisk(n, 42)
    with bh.BHEngine(n) as ref:
        ref.load_soa(*soa)
        ref.simulation_step(4)
        want = (ref.debug_get(bh.DBG.POSM), ref.debug_get(bh.DBG.VEL), ref.debug_get(bh.DBG.IDS))
    uid = bh.mg_unique_id()
    assert len(uid) == 128 and any(uid)
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        mg = bh.MultiGpu(eng, uid, 0, 1, 0)
        try:
            info = mg.info()
            assert (info["rank"], info["world"], info["first"], info["count"]) == (0, 1, 0, n) and info["per"] % 32 == 0
            mg.step(3)
            mg.step(1)
            mg.finish()
            eng.check_device_error()
            got = (eng.debug_get(bh.DBG.POSM), eng.debug_get(bh.DBG.VEL), eng.debug_get(bh.DBG.IDS))
        finally:
            mg.close()
    assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes() and (got[2] == want[2]).all()
