"""Locally-essential-tree mode (north_star "beyond ~100M bodies"): several ranks emulated on one device.
Every rank holds only its own bodies plus the point masses its peers export for its box; the forces must
agree with the single-context run and with the O(N^2) sum to the multipole error, and every export must
carry the exporter's whole mass."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
f = np.float32


def make_ranks(bh, soa, world, cap_factor=3.0):
    import torch

    from nbody_barnes_hut_cuda_b200.let import LetRank, split_by_keys

    n = len(soa[0])
    b = O.bounds(*soa[:3])
    keys, _ = O.morton_keys(*soa[:3], b)
    masks = split_by_keys(keys, world)
    dev = torch.device("cuda:0")
    ranks = []
    for r in range(world):
        sel = np.nonzero(masks[r])[0]
        posm = torch.from_numpy(np.stack([soa[0][sel], soa[1][sel], soa[2][sel], soa[6][sel]], 1).astype(f)).to(dev)
        vel = torch.from_numpy(np.stack([soa[3][sel], soa[4][sel], soa[5][sel], np.zeros(len(sel), f)], 1).astype(f)).to(dev)
        ids = torch.from_numpy(sel.astype(np.int32)).to(dev)
        ranks.append(LetRank(bh, torch, dev, posm, vel, ids, capacity=n + 4096,   # tiny ranks import more than they own
                             cap_per_peer=n, npeers=world))
    return ranks


@pytest.mark.parametrize("kind,n,world", [("uniform", 40000, 4), ("disk", 60000, 3), ("plummer", 50000, 8)])
def test_let_forces_match_the_replicated_engine_and_the_direct_sum(bh, kind, n, world):
    from nbody_barnes_hut_cuda_b200.let import let_step_emulated

    soa = {"uniform": lambda: bh.ic_uniform_cube(n, 5, 1000.0), "disk": lambda: bh.ic_refdisk(n, 42),
           "plummer": lambda: bh.ic_plummer(n, 5, 200.0, 10.0, 4.5, 0.5)}[kind]()
    posm_all, _, _ = O.soa_to_internal(soa)
    total_mass = soa[6].astype(np.float64).sum()
    ranks = make_ranks(bh, soa, world)
    local_mass = [float(r.posm[:, 3].double().sum().item()) for r in ranks]
    counts = let_step_emulated(ranks)
    acc = np.zeros((n, 3), np.float64)
    seen = np.zeros(n, int)
    for i, r in enumerate(ranks):
        ids, a = r.last_accelerations()
        acc[ids] = a
        seen[ids] += 1
        # every list this rank exported stands for ALL of its mass
        for p in range(world):
            if p == i:
                assert counts[i][p] == 0
                continue
            m = float(r.out[p, : int(counts[i][p]), 3].double().sum().item())
            assert abs(m - local_mass[i]) < 1e-5 * local_mass[i]
        # far peers need far fewer points than the rank has bodies
        assert counts[i].sum() < (world - 1) * max(r.n, 1)
    assert (seen == 1).all()                      # each body is owned by exactly one rank
    with bh.BHEngine(n) as ref:
        ref.load_soa(*soa)
        ref.simulation_step(1)
        full = np.stack(ref.read_soa()[6:9], 1).astype(np.float64)
    sample = np.arange(0, n, 16, dtype=np.int32)
    direct = O.direct_sum(posm_all, sample)
    e_let, e_full = O.rel_rms(acc[sample], direct), O.rel_rms(full[sample], direct)
    assert e_let < max(2.0 * e_full, 3e-3)        # same multipole-error class as the replicated engine
    assert O.rel_rms(acc, full) < 3.0 * max(e_full, 1e-3)
    assert abs(sum(local_mass) - total_mass) < 1e-6 * total_mass
    for r in ranks:
        r.close()


def test_let_multi_step_tracks_the_replicated_run(bh):
    from nbody_barnes_hut_cuda_b200.let import let_step_emulated

    n, world, steps = 30000, 4, 5
    soa = bh.ic_refdisk(n, 42)
    ranks = make_ranks(bh, soa, world)
    for _ in range(steps):
        let_step_emulated(ranks)
    pos = np.zeros((n, 3), f)
    for r in ranks:
        pos[r.ids.cpu().numpy()] = r.posm[:, :3].cpu().numpy()
        r.eng.check_device_error()
        r.close()
    with bh.BHEngine(n) as ref:
        ref.load_soa(*soa)
        ref.simulation_step(steps)
        want = np.stack(ref.read_soa()[:3], 1)
    assert O.rel_rms(pos, want) < 1e-5            # same physics, forces differ at the multipole-error level


def test_global_cube_matches_the_reference_bounds(bh):
    from nbody_barnes_hut_cuda_b200.let import global_cube

    soa = bh.ic_refdisk(5000, 42)
    b = O.bounds(*soa[:3])
    halves = [np.arange(5000) < 2000, np.arange(5000) >= 2000]
    boxes = np.array([[soa[a][h].min() for a in range(3)] + [soa[a][h].max() for a in range(3)] for h in halves], f)
    assert global_cube(boxes).tobytes() == b.tobytes()
