"""ctypes access to oracle/libbh_oracle.so — the CPU restatement (checker only, never the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(ROOT, "oracle", "libbh_oracle.so")
_lib = None

CHILD_EMPTY = 0x7F7F7F7F
G, THETA, DT, SOFT, VMAX = 0.5, 0.5, 0.02, 50.0, 500.0   # nbody_v5_bench.cu:13-18


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), _PATH], check=True, capture_output=True)
        _lib = C.CDLL(_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(x):
    return C.c_float(float(x))


def num_threads():
    return int(lib().orc_num_threads())


def bounds(px, py, pz):
    b = np.zeros(6, np.float32)
    lib().orc_bounds(_p(px), _p(py), _p(pz), C.c_int64(len(px)), _p(b))
    return b


def morton_keys(px, py, pz, b):
    n = len(px)
    keys, idx = np.zeros(n, np.uint32), np.zeros(n, np.int32)
    lib().orc_morton_keys(_p(px), _p(py), _p(pz), C.c_int64(n), _p(b), _p(keys), _p(idx))
    return keys, idx


def morton_keys60(px, py, pz, b):
    """hi = the reference 30-bit key, lo = 10 more bits per axis (bh_params.key_bits = 60)."""
    n = len(px)
    hi, lo = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    lib().orc_morton_keys60(_p(px), _p(py), _p(pz), C.c_int64(n), _p(b), _p(hi), _p(lo))
    return hi, lo


def stable_sort64(keys64, idx):
    k, i = np.ascontiguousarray(keys64, np.uint64).copy(), idx.copy()
    lib().orc_stable_sort64(_p(k), _p(i), C.c_int64(len(k)))
    return k, i


def tree_build64(sorted_keys64, levels):
    n = len(sorted_keys64)
    cap = max(n, 1)
    meta, child = np.zeros((cap, 4), np.int32), np.zeros((cap, 8), np.int32)
    root = C.c_int32(-1)
    k = np.ascontiguousarray(sorted_keys64, np.uint64)
    M = lib().orc_tree_build64(_p(k), C.c_int64(n), levels, _p(meta), _p(child), C.c_int64(cap), C.byref(root))
    assert M >= 0
    return meta[:M].copy(), child[:M].copy(), int(root.value)


def make_groups64(posm, sorted_keys64, levels, chunk=None, alpha=None):
    n = len(posm)
    gs = np.zeros(n + 1, np.int32)
    k = np.ascontiguousarray(sorted_keys64, np.uint64)
    ng = lib().orc_make_groups64(_p(posm), _p(k), C.c_int64(n), levels, chunk or GROUP, f32(SPLIT if alpha is None else alpha), _p(gs))
    return gs[: ng + 1].copy()


def stable_sort(keys, idx):
    k, i = keys.copy(), idx.copy()
    lib().orc_stable_sort(_p(k), _p(i), C.c_int64(len(k)))
    return k, i


def integrate(px, py, pz, vx, vy, vz, ax, ay, az, dt=DT, vmax=VMAX):
    out = [np.ascontiguousarray(a, np.float32).copy() for a in (px, py, pz, vx, vy, vz)]
    acc = [np.ascontiguousarray(a, np.float32) for a in (ax, ay, az)]
    lib().orc_integrate(*[_p(a) for a in out], *[_p(a) for a in acc], C.c_int64(len(px)), f32(dt), f32(vmax))
    return out


def reference_step(soa, nsteps=1, fixed=0, G_=G, theta=THETA, dt=DT, soft=SOFT, vmax=VMAX):
    """Oracle-L (fixed=0) / id-fixed reference tree (fixed=1): nsteps of simulationStep() on copies."""
    px, py, pz, vx, vy, vz, m = [np.ascontiguousarray(a, np.float32).copy() for a in soa]
    n = len(px)
    ax, ay, az = [np.zeros(n, np.float32) for _ in range(3)]
    ph, info = np.zeros(6), np.zeros(3, np.int64)
    keys, idx, b = np.zeros(n, np.uint32), np.zeros(n, np.int32), np.zeros(6, np.float32)
    rc = lib().orc_reference_step(_p(px), _p(py), _p(pz), _p(vx), _p(vy), _p(vz), _p(ax), _p(ay), _p(az), _p(m),
                                  C.c_int64(n), nsteps, fixed, f32(G_), f32(theta), f32(dt), f32(soft), f32(vmax),
                                  _p(ph), _p(info), _p(keys), _p(idx), _p(b))
    assert rc == 0
    return dict(px=px, py=py, pz=pz, vx=vx, vy=vy, vz=vz, ax=ax, ay=ay, az=az, phase_ms=ph, nodes=int(info[0]),
                interactions=int(info[1]), max_stack=int(info[2]), keys=keys, idx=idx, bounds=b)


def tree_build(sorted_keys):
    n = len(sorted_keys)
    cap = max(n, 1)
    meta, child = np.zeros((cap, 4), np.int32), np.zeros((cap, 8), np.int32)
    root = C.c_int32(-1)
    M = lib().orc_tree_build(_p(sorted_keys), C.c_int64(n), _p(meta), _p(child), C.c_int64(cap), C.byref(root))
    assert M >= 0
    return meta[:M].copy(), child[:M].copy(), int(root.value)


def tree_com(posm, meta, child, root):
    M = len(meta)
    mom, com = np.zeros((max(M, 1), 4), np.float32), np.zeros((max(M, 1), 4), np.float32)
    lib().orc_tree_com(_p(posm), C.c_int64(len(posm)), _p(meta), _p(child), C.c_int64(M), C.c_int32(root), _p(mom), _p(com))
    return mom[:M], com[:M]


def tree_quad(posm, meta):
    """Traceless quadrupoles {xx,xy,xz,yy,yz,zz} of every cell about its centre of mass (direct double sums)."""
    M = len(meta)
    quad = np.zeros((max(M, 1), 6), np.float32)
    lib().orc_tree_quad(_p(posm), _p(np.ascontiguousarray(meta)), C.c_int64(M), _p(quad))
    return quad[:M]


def force_groups_quad(posm, b, meta, child, com, quad, root, gstart, theta=THETA, soft=SOFT, G_=G):
    n = len(posm)
    acc, counts = np.zeros((n, 4), np.float32), np.zeros(2, np.int64)
    lib().orc_force_groups_quad(_p(posm), C.c_int64(n), _p(b), _p(meta), _p(child), _p(com), _p(np.ascontiguousarray(quad)),
                                C.c_int64(len(meta)), C.c_int32(root), _p(gstart), len(gstart) - 1, f32(theta), f32(soft), f32(G_),
                                _p(acc), _p(counts), None)
    return acc, counts


def force_group(posm, b, meta, child, com, root, group=32, theta=THETA, soft=SOFT, G_=G, entries=None):
    n = len(posm)
    acc, counts = np.zeros((n, 4), np.float32), np.zeros(2, np.int64)
    lib().orc_force_group(_p(posm), C.c_int64(n), _p(b), _p(meta), _p(child), _p(com), C.c_int64(len(meta)),
                          C.c_int32(root), group, f32(theta), f32(soft), f32(G_), _p(acc), _p(counts), _p(entries))
    return acc, counts


SPLIT = 0.5   # bh_params.group_split default
GROUP = 32    # bodies per traversal chunk = bh_group_size() (tests/test_abi.py checks the two agree)


def make_groups(posm, sorted_keys, chunk=GROUP, alpha=SPLIT):
    """Traversal groups of the engine: chunks of the Morton order, cut at coarse key boundaries."""
    n = len(posm)
    gs = np.zeros(n + 1, np.int32)
    ng = lib().orc_make_groups(_p(posm), _p(sorted_keys), C.c_int64(n), chunk, f32(alpha), _p(gs))
    return gs[: ng + 1].copy()


def force_groups(posm, b, meta, child, com, root, gstart, theta=THETA, soft=SOFT, G_=G, entries=None):
    n = len(posm)
    acc, counts = np.zeros((n, 4), np.float32), np.zeros(2, np.int64)
    lib().orc_force_groups(_p(posm), C.c_int64(n), _p(b), _p(meta), _p(child), _p(com), C.c_int64(len(meta)),
                           C.c_int32(root), _p(gstart), len(gstart) - 1, f32(theta), f32(soft), f32(G_), _p(acc),
                           _p(counts), _p(entries))
    return acc, counts


def force_body(posm, b, meta, child, com, root, theta=THETA, soft=SOFT, G_=G):
    n = len(posm)
    acc, counts = np.zeros((n, 4), np.float32), np.zeros(2, np.int64)
    lib().orc_force_body(_p(posm), C.c_int64(n), _p(b), _p(meta), _p(child), _p(com), C.c_int64(len(meta)),
                         C.c_int32(root), f32(theta), f32(soft), f32(G_), _p(acc), _p(counts))
    return acc, counts


def direct_sum(posm, sample, soft=SOFT, G_=G):
    s = np.ascontiguousarray(sample, np.int32)
    out = np.zeros((len(s), 3), np.float64)
    lib().orc_direct_sum(_p(posm), C.c_int64(len(posm)), _p(s), len(s), f32(soft), f32(G_), _p(out))
    return out


def energy(posm, vel, soft=SOFT, G_=G):
    ke, pe = C.c_double(), C.c_double()
    lib().orc_energy(_p(posm), _p(vel), C.c_int64(len(posm)), f32(soft), f32(G_), C.byref(ke), C.byref(pe))
    return ke.value, pe.value


def engine_step(posm, vel, ids, nsteps=1, group=GROUP, G_=G, theta=THETA, dt=DT, soft=SOFT, vmax=VMAX, alpha=SPLIT,
                slice_first=0, slice_count=-1, key_bits=30):
    """Oracle-I: nsteps of the shipped algorithm on copies of the internal-layout state."""
    posm, vel, ids = posm.copy(), vel.copy(), ids.copy()
    n = len(posm)
    acc, keys, perm = np.zeros((n, 4), np.float32), np.zeros(n, np.uint32), np.zeros(n, np.int32)
    b, counts, ph = np.zeros(6, np.float32), np.zeros(3, np.int64), np.zeros(6)
    keys_lo = np.zeros(n, np.uint32)
    rc = lib().orc_engine_step(_p(posm), _p(vel), _p(ids), C.c_int64(n), nsteps, f32(G_), f32(theta), f32(dt), f32(soft),
                               f32(vmax), group, f32(alpha), _p(acc), _p(keys), _p(perm), _p(b), _p(counts), _p(ph),
                               C.c_int64(slice_first), C.c_int64(slice_count), key_bits, _p(keys_lo))
    assert rc == 0
    return dict(posm=posm, vel=vel, ids=ids, acc=acc, keys=keys, keys_lo=keys_lo, perm=perm, bounds=b, inter_cell=int(counts[0]),
                inter_body=int(counts[1]), cells=int(counts[2]), phase_ms=ph)


def soa_to_internal(soa):
    px, py, pz, vx, vy, vz, m = soa
    n = len(px)
    posm = np.stack([px, py, pz, m], 1).astype(np.float32).copy()
    vel = np.stack([vx, vy, vz, np.zeros(n, np.float32)], 1).astype(np.float32).copy()
    return posm, vel, np.arange(n, dtype=np.int32)


def rel_rms(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.sqrt(((a - ref) ** 2).sum() / (ref ** 2).sum()))


def cell_tuples(meta, keys_sorted):
    """numbering-independent view of a tree: set of (level, prefix, first, count)."""
    out = set()
    for first, count, lvflag, _parent in meta:
        L = int(lvflag) & 0xFF
        prefix = int(keys_sorted[first]) >> (30 - 3 * L) if L > 0 else 0
        out.add((L, prefix, int(first), int(count)))
    return out


# ---- locally-essential-tree export (csrc/bh_let.cu) -----------------------------------------------------
# The reference has no multi-GPU path, so there is nothing of its own to restate here; this is the CPU
# statement of OUR export rule, used only to check the kernel: walk the tree from the root; a cell whose
# width passes the reference's acceptance test (bench:207-208, squared) at the nearest point of the peer's
# boxes is emitted as {com, mass}; otherwise it is opened; loose bodies and the bodies of rejected buckets are
# emitted as they are.  Same float32 operation order as the kernel (the hull shortcut included).
def _box_centre_half(lohi):
    lohi = np.asarray(lohi, np.float32)
    return ((lohi[..., :3] + lohi[..., 3:]) * np.float32(0.5)).astype(np.float32), \
           ((lohi[..., 3:] - lohi[..., :3]) * np.float32(0.5)).astype(np.float32)


def _fma32(a, b, c):
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def _box_dist2(c, h, p):
    d = np.maximum(np.float32(0), (np.abs((p[:3] - c).astype(np.float32)) - h).astype(np.float32)).astype(np.float32)
    return _fma32(d[2], d[2], _fma32(d[1], d[1], np.float32(d[0] * d[0])))


def let_export_points(meta, child, com, posm_sorted, root, boxes_lohi, root_w, theta=THETA, soft=SOFT):
    """Points (k,4 float32, order unspecified) the export walk emits for ONE peer described by boxes_lohi [K,6]."""
    boxes = np.asarray(boxes_lohi, np.float32).reshape(-1, 6)
    boxes = boxes[boxes[:, 0] <= boxes[:, 3]]
    if len(boxes) == 0:
        return np.zeros((0, 4), np.float32)
    bc, bh_ = _box_centre_half(boxes)
    hull = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)]).astype(np.float32)
    hc, hh = _box_centre_half(hull)
    theta2, soft = np.float32(np.float32(theta) * np.float32(theta)), np.float32(soft)
    w2_root = np.float32(np.float32(root_w) * np.float32(root_w))

    def accepts(p, level):
        w2 = np.float32(w2_root * np.float32(4.0 ** -level))
        if w2 < np.float32(theta2 * np.float32(_box_dist2(hc, hh, p) + soft)):
            return True
        d2 = min(_box_dist2(bc[k], bh_[k], p) for k in range(len(boxes)))
        return bool(w2 < np.float32(theta2 * np.float32(d2 + soft)))

    out = []

    def visit(cell):
        first, count, lvflag, _ = (int(v) for v in meta[cell])
        if accepts(com[cell], lvflag & 0xFF):
            out.append(com[cell])
        elif (lvflag >> 8) & 1:
            out.extend(posm_sorted[first:first + count])
        else:
            for c in child[cell]:
                c = int(c)
                if c == 0x7F7F7F7F:
                    continue
                if c < 0:
                    out.append(posm_sorted[c & 0x7FFFFFFF])
                else:
                    stack.append(c)

    stack = [int(root)]
    while stack:
        visit(stack.pop())
    return np.array(out, np.float32).reshape(-1, 4)


# ---- host-side decisions of the LET mode: plain-Python statements the native versions (csrc/bh_let_host.cpp:
# bh_let_elect_splitters, bh_let_domain_cuts) are checked against ------------------------------------------
LET_KEY_END = 1 << 30
LET_MAX_BOXES = 200


def let_elect_splitters_ref(samples: np.ndarray, work: np.ndarray) -> np.ndarray:
    """Key-range edges [world+1] from every rank's key sample.

    samples [world, SAMPLE]: keys of rank r at equal increments of its cumulative work, so that every sample
    stands for work[r]/SAMPLE (rows of ranks with work 0 are ignored).  Edge r is the key at r/world of the
    pooled work.  Identical on all ranks (pure function of all-gathered data)."""
    world = len(work)
    work = np.asarray(work, np.float64)
    if not (work > 0).any():
        return np.array([0] + [LET_KEY_END] * world, np.int64)
    wgt = np.repeat(np.where(work > 0, work / samples.shape[1], 0.0), samples.shape[1])
    keys = np.asarray(samples, np.int64).reshape(-1)
    order = np.argsort(keys, kind="stable")
    keys, cum = keys[order], np.cumsum(wgt[order])
    edges = [0]
    for r in range(1, world):
        k = int(keys[min(int(np.searchsorted(cum, cum[-1] * r / world)), len(keys) - 1)])
        edges.append(max(k, edges[-1]))
    edges.append(LET_KEY_END)
    return np.array(edges, np.int64)


def let_domain_cuts_ref(k_lo: int, k_hi: int) -> np.ndarray:
    """LET_MAX_BOXES+1 ascending keys that cut [k_lo, k_hi) at octree-cell boundaries.

    Interior: the 8..64 cells of size S = 8^j (largest with span/S >= 8) that lie inside the range — whole
    cells, convex, owned by this rank alone.  The two ragged ends (parts of one S-cell each) are cut again
    at S/64 so that their boxes reach at most one small cell into the neighbour's range.  Padded with k_hi
    (empty intervals) to a fixed length so the boxes can be all-gathered."""
    cuts = {int(k_lo), int(k_hi)}
    span = int(k_hi) - int(k_lo)
    if span > 0:
        S = 1
        while S * 64 <= span:
            S *= 8
        first = -(-int(k_lo) // S) * S
        last = int(k_hi) // S * S
        if first <= last:
            cuts.update(range(first, last + 1, S))
            fine = max(S // 64, 1)
            for a, b in ((int(k_lo), first), (last, int(k_hi))):
                if b - a > fine:
                    cuts.update(range(-(-a // fine) * fine, b, fine))
        else:                                  # the whole range lies inside one S-cell
            fine = max(S // 64, 1)
            cuts.update(range(-(-int(k_lo) // fine) * fine, int(k_hi), fine))
    out = sorted(c for c in cuts if k_lo <= c <= k_hi)
    assert len(out) <= LET_MAX_BOXES + 1, len(out)
    out += [int(k_hi)] * (LET_MAX_BOXES + 1 - len(out))
    return np.array(out, np.uint32)
