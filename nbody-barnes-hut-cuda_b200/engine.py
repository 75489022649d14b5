"""ctypes binding of include/bh.h plus a thin object wrapper.

Names follow the reference: ``BHEngine.simulation_step`` is ``simulationStep()``
(nbody_v5_bench.cu:255), ``load_soa`` is the H2D block of ``main()`` (bench:329-335).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("BH_LIB") or os.path.join(_HERE, "libbh.so")  # BH_LIB: tuning variants only
_IC_LIB_PATH = os.path.join(_HERE, "libbh_ic.so")   # host-only build of csrc/bh_ic.cpp (same bh_ic_* symbols as libbh.so)
_lib: Optional[C.CDLL] = None
_ic_lib: Optional[C.CDLL] = None


class BHError(RuntimeError):
    pass


class BHParams(C.Structure):
    """bh_params — the reference's #defines (nbody_v5_bench.cu:13-18)."""

    _fields_ = [
        ("theta", C.c_float),
        ("G", C.c_float),
        ("dt", C.c_float),
        ("softening", C.c_float),
        ("max_speed", C.c_float),
        ("key_bits", C.c_int),
        ("leaf_cap", C.c_int),
        ("flags", C.c_int),
        ("group_split", C.c_float),
    ]


FLAG_NO_GRAPH = 1
FLAG_PHASE_TIMER = 2
FLAG_QUADRUPOLE = 4   # accepted cells act with their quadrupole too (set at creation)


class PHASE:
    KEYS, SORT, BUILD, COM, FORCE, UPDATE, TOTAL, COUNT = range(8)
    NAMES = ["keys", "sort", "build", "com", "force", "update", "total"]


class DBG:
    (BOUNDS, KEYS, PERM, IDS, POSM, VEL, ACC, CELL_META, CELL_COM, CELL_CHILD,
     POSM_SORTED, VEL_SORTED, IDS_SORTED, KEYS64, CELL_QUAD) = range(15)


class STAT:
    N, CELLS, ROOT, INTERACTIONS_CELL, INTERACTIONS_BODY, DEVICE_ERROR, STEPS, MAX_STACK = range(8)


CHILD_EMPTY = 0x7F7F7F7F


def library_path() -> str:
    return _LIB_PATH


def build_library(verbose: bool = False) -> str:
    """Compile libbh.so in tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise BHError("building libbh.so failed")
    return _LIB_PATH


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):  # torch tensor (host or device)
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot pass {type(a)} as a pointer")


def lib() -> C.CDLL:
    """Load libbh.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise BHError(f"{_LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
    L.bh_default_params.argtypes = [C.POINTER(BHParams)]
    L.bh_default_params.restype = None
    L.bh_abi_version.restype = i32
    L.bh_group_size.restype = i32
    L.bh_error_string.argtypes = [i32]
    L.bh_error_string.restype = C.c_char_p
    L.bh_create.argtypes = [C.POINTER(vp), i64, C.POINTER(BHParams), i32]
    L.bh_destroy.argtypes = [vp]
    L.bh_destroy.restype = None
    L.bh_import_soa.argtypes = [vp] + [vp] * 7 + [i64, vp]
    L.bh_import_soa_host.argtypes = [vp] + [vp] * 7 + [i64]
    L.bh_step.argtypes = [vp, i32, vp]
    L.bh_step_half.argtypes = [vp, i32, vp]
    L.bh_step_part.argtypes = [vp, i32, vp]
    L.bh_export_soa.argtypes = [vp] + [vp] * 9 + [vp]
    L.bh_export_soa_host.argtypes = [vp] + [vp] * 9
    L.bh_step_host.argtypes = [vp] + [vp] * 7 + [i64, i32]
    L.bh_phase_ms.argtypes = [vp, C.POINTER(f32)]
    L.bh_run_phase.argtypes = [vp, i32, vp]
    L.bh_debug_get.argtypes = [vp, i32, vp, C.c_size_t]
    L.bh_debug_set.argtypes = [vp, i32, vp, C.c_size_t]
    L.bh_stat.argtypes = [vp, i32]
    L.bh_stat.restype = i64
    L.bh_set_slice.argtypes = [vp, i32, i32]
    L.bh_set_fixed_bounds.argtypes = [vp, vp]
    L.bh_local_bounds.argtypes = [vp, vp]
    L.bh_import_state.argtypes = [vp, vp, vp, vp, i64, vp]
    L.bh_let_export.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp]
    L.bh_let_domain_cuts.argtypes = [C.c_uint32, C.c_uint32, i32, vp]
    L.bh_let_elect_splitters.argtypes = [vp, i32, vp, i32, vp]
    L.bh_force_from.argtypes = [vp, vp, vp]
    L.bh_sort_coarse.argtypes = [vp, vp]
    L.bh_export_real.argtypes = [vp, vp, vp, vp, C.POINTER(i64), vp]
    L.bh_let_domain_boxes.argtypes = [vp, vp, i32, vp, vp]
    L.bh_sorted_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    L.bh_state_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.bh_sort_pairs_u32.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp, C.POINTER(C.c_size_t), vp]
    L.bh_direct_sample.argtypes = [vp, vp, i32, vp]
    L.bh_energy.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.bh_export_visuals.argtypes = [vp, vp, vp, vp]
    L.bh_momentum.argtypes = [vp, vp]
    L.bh_dump_text.argtypes = [vp, C.c_char_p]
    L.bh_save_checkpoint.argtypes = [vp, C.c_char_p]
    L.bh_load_checkpoint.argtypes = [vp, C.c_char_p]
    L.bh_ic_refdisk.argtypes = [i64, C.c_uint] + [vp] * 7
    L.bh_ic_uniform_cube.argtypes = [i64, C.c_uint64, f32] + [vp] * 7
    L.bh_ic_two_disks.argtypes = [i64, C.c_uint64, f32, f32, f32] + [vp] * 7
    L.bh_ic_two_disks_range.argtypes = [i64, i64, C.c_uint64, f32, f32, f32] + [vp] * 7
    L.bh_set_flags.argtypes = [vp, i32]
    L.bh_ic_plummer.argtypes = [i64, C.c_uint64, f32, f32, f32, f32] + [vp] * 7
    L.bh_mg_unique_id.argtypes = [vp]
    L.bh_mg_create.argtypes = [C.POINTER(vp), vp, vp, i32, i32, i32]
    L.bh_mg_step.argtypes = [vp, i32, vp]
    L.bh_mg_finish.argtypes = [vp, vp]
    L.bh_mg_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.bh_mg_destroy.argtypes = [vp]
    L.bh_mg_destroy.restype = None
    L.bh_probe_fp32_tflops.argtypes = [i32, C.POINTER(f32)]
    L.bh_probe_hbm_gbs.argtypes = [i32, C.POINTER(f32)]
    L.bh_probe_fp32x2_tflops.argtypes = [i32, C.POINTER(f32)]
    _lib = L
    return L


def ic_lib() -> C.CDLL:
    """The initial-condition generators (host C++).  A library of their own, so that callers which must not touch
    the engine — bench.py's reference arm, the CPU baseline — can make inputs without mapping libbh.so."""
    global _ic_lib
    if _ic_lib is not None:
        return _ic_lib
    if not os.path.exists(_IC_LIB_PATH):
        raise BHError(f"{_IC_LIB_PATH} is missing: run __graft_entry__.build()")
    L = C.CDLL(_IC_LIB_PATH)
    vp, i64, f32 = C.c_void_p, C.c_int64, C.c_float
    L.bh_ic_refdisk.argtypes = [i64, C.c_uint] + [vp] * 7
    L.bh_ic_uniform_cube.argtypes = [i64, C.c_uint64, f32] + [vp] * 7
    L.bh_ic_two_disks.argtypes = [i64, C.c_uint64, f32, f32, f32] + [vp] * 7
    L.bh_ic_two_disks_range.argtypes = [i64, i64, C.c_uint64, f32, f32, f32] + [vp] * 7
    L.bh_ic_plummer.argtypes = [i64, C.c_uint64, f32, f32, f32, f32] + [vp] * 7
    _ic_lib = L
    return L


def _check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().bh_error_string(code).decode()
        raise BHError(f"{what} failed: {code} ({msg})")


def _check_ic(code: int, what: str) -> None:
    if code != 0:
        raise BHError(f"{what} failed: {code}")   # no bh_error_string here: the IC path never loads libbh.so


# ---- initial conditions (host) ------------------------------------------------------------
def _soa(n: int):
    return [np.zeros(n, np.float32) for _ in range(7)]


def ic_refdisk(n: int, seed: int = 42):
    """main()'s disk, nbody_v5_bench.cu:294-308 (glibc rand)."""
    a = _soa(n)
    _check_ic(ic_lib().bh_ic_refdisk(n, seed, *[_vp(x) for x in a]), "bh_ic_refdisk")
    return a


def ic_uniform_cube(n: int, seed: int = 42, half_edge: float = 1000.0):
    a = _soa(n)
    _check_ic(ic_lib().bh_ic_uniform_cube(n, seed, half_edge, *[_vp(x) for x in a]), "bh_ic_uniform_cube")
    return a


def ic_plummer(n: int, seed: int = 42, scale_a: float = 200.0, rcut_in_a: float = 10.0,
               body_mass: float = 4.5, G: float = 0.5):
    a = _soa(n)
    _check_ic(ic_lib().bh_ic_plummer(n, seed, scale_a, rcut_in_a, body_mass, G, *[_vp(x) for x in a]), "bh_ic_plummer")
    return a


def ic_two_disks(n: int, seed: int = 42, sep: float = 4000.0, vx: float = 20.0, vy: float = 8.0):
    """BASELINE.json configs[4]: two reference-style discs on a collision course."""
    a = _soa(n)
    _check_ic(ic_lib().bh_ic_two_disks(n, seed, sep, vx, vy, *[_vp(x) for x in a]), "bh_ic_two_disks")
    return a


def ic_two_disks_range(first: int, n: int, seed: int = 42, sep: float = 4000.0, vx: float = 20.0, vy: float = 8.0):
    """Bodies [first, first+n) of ic_two_disks: each rank of a multi-GPU run generates only its share."""
    a = _soa(n)
    _check_ic(ic_lib().bh_ic_two_disks_range(first, n, seed, sep, vx, vy, *[_vp(x) for x in a]), "bh_ic_two_disks_range")
    return a


MG_ID_BYTES = 128


def mg_unique_id() -> bytes:
    """Rank 0: the 128-byte NCCL id the other ranks need for MultiGpu (send it with whatever the host has)."""
    buf = C.create_string_buffer(MG_ID_BYTES)
    _check(lib().bh_mg_unique_id(buf), "bh_mg_unique_id")
    return buf.raw


class MultiGpu:
    """bh_mg_* — the C++/NCCL driver of the Morton-slice mode (csrc/bh_mg.cu).  One per process/GPU; the engine must
    hold the FULL state.  step() is asynchronous; finish() makes `stream` wait for the gathers in flight."""

    def __init__(self, engine: "BHEngine", unique_id: bytes, rank: int, world: int, device: int):
        assert len(unique_id) == MG_ID_BYTES
        self._mg = C.c_void_p()
        self._id = C.create_string_buffer(unique_id, MG_ID_BYTES)
        _check(lib().bh_mg_create(C.byref(self._mg), engine._ctx, self._id, rank, world, device), "bh_mg_create")

    def step(self, nsteps: int = 1, stream: int = 0):
        _check(lib().bh_mg_step(self._mg, nsteps, C.c_void_p(stream)), "bh_mg_step")

    def finish(self, stream: int = 0):
        _check(lib().bh_mg_finish(self._mg, C.c_void_p(stream)), "bh_mg_finish")

    def info(self) -> dict:
        r, w = C.c_int(), C.c_int()
        first, count, per = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().bh_mg_info(self._mg, C.byref(r), C.byref(w), C.byref(first), C.byref(count), C.byref(per)), "bh_mg_info")
        return {"rank": r.value, "world": w.value, "first": first.value, "count": count.value, "per": per.value}

    def close(self):
        if self._mg:
            lib().bh_mg_destroy(self._mg)
            self._mg = C.c_void_p()


def probe_fp32_tflops(device: int = 0) -> float:
    v = C.c_float()
    _check(lib().bh_probe_fp32_tflops(device, C.byref(v)), "bh_probe_fp32_tflops")
    return float(v.value)


def probe_fp32x2_tflops(device: int = 0) -> float:
    v = C.c_float()
    _check(lib().bh_probe_fp32x2_tflops(device, C.byref(v)), "bh_probe_fp32x2_tflops")
    return float(v.value)


def probe_hbm_gbs(device: int = 0) -> float:
    v = C.c_float()
    _check(lib().bh_probe_hbm_gbs(device, C.byref(v)), "bh_probe_hbm_gbs")
    return float(v.value)


def sort_pairs_u32(keys_in, vals_in, keys_out, vals_out, n: int, begin_bit: int = 0, end_bit: int = 32,
                   tmp=None, stream: int = 0):
    """Standalone onesweep sort on DEVICE pointers (ints or torch tensors). Returns scratch bytes if tmp is None."""
    need = C.c_size_t(0)
    if tmp is None:
        _check(lib().bh_sort_pairs_u32(None, None, None, None, n, begin_bit, end_bit, None, C.byref(need), None),
               "bh_sort_pairs_u32(size)")
        return int(need.value)
    need = C.c_size_t(tmp.numel() * tmp.element_size())
    _check(lib().bh_sort_pairs_u32(_vp(keys_in), _vp(vals_in), _vp(keys_out), _vp(vals_out), n, begin_bit, end_bit,
                                   _vp(tmp), C.byref(need), C.c_void_p(stream)), "bh_sort_pairs_u32")
    return 0


class BHEngine:
    """One Barnes-Hut context on one GPU (bh_ctx)."""

    def __init__(self, n_max: int, device: int = 0, flags: int = 0, **overrides):
        L = lib()
        self.params = BHParams()
        L.bh_default_params(C.byref(self.params))
        self.params.flags = flags
        for k, v in overrides.items():
            if not hasattr(self.params, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(self.params, k, v)
        self._ctx = C.c_void_p()
        _check(L.bh_create(C.byref(self._ctx), n_max, C.byref(self.params), device), "bh_create")
        self.n_max = n_max
        self.device = device

    # -- lifetime
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            lib().bh_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state in / out
    def load_soa(self, px, py, pz, vx, vy, vz, mass):
        """Host numpy (or pinned torch) arrays -> device state (bench:329-335)."""
        arrs = [px, py, pz, vx, vy, vz, mass]
        n = len(px)
        self._keep = arrs
        _check(lib().bh_import_soa_host(self._ctx, *[_vp(a) for a in arrs], n), "bh_import_soa_host")

    def load_soa_device(self, ptrs: Sequence, n: int, stream: int = 0):
        _check(lib().bh_import_soa(self._ctx, *[_vp(p) for p in ptrs], n, C.c_void_p(stream)), "bh_import_soa")

    def simulation_step(self, nsteps: int = 1, stream: int = 0):
        """≙ simulationStep() x nsteps (bench:255-283); asynchronous."""
        _check(lib().bh_step(self._ctx, nsteps, C.c_void_p(stream)), "bh_step")

    def step_half(self, half: int, stream: int = 0):
        """Head (0: bounds, keys, sort — positions only) or tail (1: the rest) of one step."""
        _check(lib().bh_step_half(self._ctx, half, C.c_void_p(stream)), "bh_step_half")

    def step_part(self, part: int, stream: int = 0):
        """One of the three parts of a step (0: cube, keys, sort; 1: tree, centre of mass, traversal; 2: update)."""
        _check(lib().bh_step_part(self._ctx, part, C.c_void_p(stream)), "bh_step_part")

    def read_soa(self, want_acc: bool = True):
        """Device state -> host SoA in ORIGINAL body order: (px,py,pz,vx,vy,vz[,ax,ay,az])."""
        n = self.n
        out = [np.zeros(n, np.float32) for _ in range(9 if want_acc else 6)]
        ptrs = [_vp(a) for a in out] + [None] * (9 - len(out))
        _check(lib().bh_export_soa_host(self._ctx, *ptrs), "bh_export_soa_host")
        return out

    def step_host(self, px, py, pz, vx, vy, vz, mass, nsteps: int = 1):
        """Host in -> nsteps -> host out, every copy inside the call (bench.py e2e leg)."""
        n = len(px)
        _check(lib().bh_step_host(self._ctx, *[_vp(a) for a in (px, py, pz, vx, vy, vz, mass)], n, nsteps), "bh_step_host")

    # -- introspection
    @property
    def n(self) -> int:
        return int(lib().bh_stat(self._ctx, STAT.N))

    def stat(self, which: int) -> int:
        return int(lib().bh_stat(self._ctx, which))

    def check_device_error(self):
        e = self.stat(STAT.DEVICE_ERROR)
        if e:
            raise BHError(f"device error flag = {e}")

    def run_phase(self, phase: int, stream: int = 0):
        _check(lib().bh_run_phase(self._ctx, phase, C.c_void_p(stream)), f"bh_run_phase({phase})")

    def set_flags(self, flags: int):
        _check(lib().bh_set_flags(self._ctx, flags), "bh_set_flags")

    def phase_ms(self):
        out = (C.c_float * PHASE.COUNT)()
        _check(lib().bh_phase_ms(self._ctx, out), "bh_phase_ms")
        return {PHASE.NAMES[i]: float(out[i]) for i in range(PHASE.TOTAL + 1)}

    def debug_get(self, what: int) -> np.ndarray:
        n, M = self.n, None
        if what in (DBG.CELL_META, DBG.CELL_COM, DBG.CELL_CHILD, DBG.CELL_QUAD):
            M = self.stat(STAT.CELLS)
        shapes = {
            DBG.BOUNDS: ((6,), np.float32), DBG.KEYS: ((n,), np.uint32), DBG.PERM: ((n,), np.int32),
            DBG.IDS: ((n,), np.int32), DBG.POSM: ((n, 4), np.float32), DBG.VEL: ((n, 4), np.float32),
            DBG.ACC: ((n, 4), np.float32), DBG.CELL_META: ((M, 4), np.int32), DBG.CELL_COM: ((M, 4), np.float32),
            DBG.CELL_CHILD: ((M, 8), np.int32), DBG.POSM_SORTED: ((n, 4), np.float32),
            DBG.VEL_SORTED: ((n, 4), np.float32), DBG.IDS_SORTED: ((n,), np.int32), DBG.KEYS64: ((n,), np.uint64),
            DBG.CELL_QUAD: ((M, 8), np.float32),
        }
        shape, dt = shapes[what]
        out = np.zeros(shape, dt)
        _check(lib().bh_debug_get(self._ctx, what, _vp(out), out.nbytes), f"bh_debug_get({what})")
        return out

    def debug_set(self, what: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        _check(lib().bh_debug_set(self._ctx, what, _vp(arr), arr.nbytes), f"bh_debug_set({what})")

    def direct_sample(self, sample_ids: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(sample_ids, np.int32)
        out = np.zeros((len(s), 3), np.float64)
        _check(lib().bh_direct_sample(self._ctx, _vp(s), len(s), _vp(out)), "bh_direct_sample")
        return out

    def energy(self):
        ke, pe = C.c_double(), C.c_double()
        _check(lib().bh_energy(self._ctx, C.byref(ke), C.byref(pe)), "bh_energy")
        return float(ke.value), float(pe.value)

    def momentum(self):
        out = np.zeros(7, np.float64)
        _check(lib().bh_momentum(self._ctx, _vp(out)), "bh_momentum")
        return out

    def export_visuals(self, vbo_p, vbo_c, stream: int = 0):
        """≙ updateVisualsKernel (nbody_v5.cu:278-292); DEVICE pointers / torch tensors."""
        _check(lib().bh_export_visuals(self._ctx, _vp(vbo_p), _vp(vbo_c), C.c_void_p(stream)), "bh_export_visuals")

    def dump_text(self, path: str):
        _check(lib().bh_dump_text(self._ctx, path.encode()), "bh_dump_text")

    def save_checkpoint(self, path: str):
        _check(lib().bh_save_checkpoint(self._ctx, path.encode()), "bh_save_checkpoint")

    def load_checkpoint(self, path: str):
        _check(lib().bh_load_checkpoint(self._ctx, path.encode()), "bh_load_checkpoint")

    # -- locally-essential-tree mode (include/bh.h "multi-GPU: locally-essential-tree exchange")
    def set_fixed_bounds(self, b):
        arr = None if b is None else np.ascontiguousarray(b, np.float32)
        self._fixed_bounds = arr
        _check(lib().bh_set_fixed_bounds(self._ctx, _vp(arr)), "bh_set_fixed_bounds")

    def local_bounds(self) -> np.ndarray:
        out = np.zeros(6, np.float32)
        _check(lib().bh_local_bounds(self._ctx, _vp(out)), "bh_local_bounds")
        return out

    def import_state(self, posm, vel, ids, n: int, stream: int = 0):
        _check(lib().bh_import_state(self._ctx, _vp(posm), _vp(vel), _vp(ids), n, C.c_void_p(stream)), "bh_import_state")

    def sort_coarse(self, stream: int = 0):
        """keys + sort + reorder on the 30-bit reference key only (splitter election / migration)."""
        _check(lib().bh_sort_coarse(self._ctx, C.c_void_p(stream)), "bh_sort_coarse")

    def force_from(self, sources: "BHEngine", stream: int = 0):
        """acc += attraction of every body of `sources`, traversing ITS tree with this context's groups."""
        _check(lib().bh_force_from(self._ctx, sources._ctx, C.c_void_p(stream)), "bh_force_from")

    def export_real(self, posm_out, vel_out, ids_out, stream: int = 0) -> int:
        """Own bodies (id >= 0) of the current state -> device arrays, vel.w = chunk work; returns their number."""
        m = C.c_int64()
        _check(lib().bh_export_real(self._ctx, _vp(posm_out), _vp(vel_out), _vp(ids_out), C.byref(m), C.c_void_p(stream)),
               "bh_export_real")
        return int(m.value)

    def let_domain_boxes(self, cuts):
        """cuts: K+1 ascending u32 keys -> (boxes [K,6] lo/hi, body counts [K])."""
        cuts = np.ascontiguousarray(cuts, np.uint32)
        k = len(cuts) - 1
        out, counts = np.zeros((k, 6), np.float32), np.zeros(k, np.int32)
        _check(lib().bh_let_domain_boxes(self._ctx, _vp(cuts), k, _vp(out), _vp(counts)), "bh_let_domain_boxes")
        return out, counts

    def sorted_ptrs(self):
        vp, i64 = C.c_void_p, C.c_int64
        keys, posm, vel, ids, acc, n = vp(), vp(), vp(), vp(), vp(), i64()
        _check(lib().bh_sorted_ptrs(self._ctx, C.byref(keys), C.byref(posm), C.byref(vel), C.byref(ids), C.byref(acc),
                                    C.byref(n)), "bh_sorted_ptrs")
        return dict(keys=keys.value, posm=posm.value, vel=vel.value, ids=ids.value, acc=acc.value, n=n.value)

    def let_export(self, boxes_lohi, out, cap_per_peer: int, stream: int = 0) -> np.ndarray:
        """boxes_lohi: [npeers, K, 6] host array (lo xyz, hi xyz per box)."""
        boxes = np.ascontiguousarray(boxes_lohi, np.float32)
        npeers, k = boxes.shape[0], boxes.shape[1]
        counts = np.zeros(npeers, np.int32)
        _check(lib().bh_let_export(self._ctx, _vp(boxes), npeers, k, _vp(out), cap_per_peer, _vp(counts),
                                   C.c_void_p(stream)), "bh_let_export")
        return counts

    # -- multi-GPU slices
    def set_slice(self, rank: int, world: int):
        _check(lib().bh_set_slice(self._ctx, rank, world), "bh_set_slice")

    def state_ptrs(self):
        vp, i64 = C.c_void_p, C.c_int64
        posm, vel, ids, n, first, count = vp(), vp(), vp(), i64(), i64(), i64()
        _check(lib().bh_state_ptrs(self._ctx, C.byref(posm), C.byref(vel), C.byref(ids), C.byref(n), C.byref(first),
                                   C.byref(count)), "bh_state_ptrs")
        return dict(posm=posm.value, vel=vel.value, ids=ids.value, n=n.value, first=first.value, count=count.value)
