"""The C-ABI library loads without a GPU and exports every symbol include/bh.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bh.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bh_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("bh_create", "bh_destroy", "bh_import_soa", "bh_step", "bh_export_soa", "bh_step_host",
                 "bh_phase_ms", "bh_debug_get", "bh_sort_pairs_u32", "bh_set_slice"):
        assert must in names


def test_library_exports_every_declared_symbol(bh):
    L = bh.lib()
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing
    assert L.bh_abi_version() == 2
    import oracle_lib as O

    assert L.bh_group_size() == O.GROUP


def test_default_params_are_the_reference_constants(bh):
    p = bh.BHParams()
    bh.lib().bh_default_params(C.byref(p))
    # nbody_v5_bench.cu:13-18
    assert (p.theta, p.G, p.softening, p.max_speed) == (0.5, 0.5, 50.0, 500.0)
    assert abs(p.dt - 0.02) < 1e-9 and p.key_bits == 30 and p.leaf_cap == 1


def test_error_strings(bh):
    L = bh.lib()
    assert L.bh_error_string(0) == b"ok"
    assert b"invalid" in L.bh_error_string(-1)


def test_no_silent_cpu_fallback(bh):
    """Without a device the engine must refuse loudly, not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bh.BHError):
        bh.BHEngine(1024)


def test_bad_arguments_rejected_before_touching_cuda(bh):
    L = bh.lib()
    ctx = C.c_void_p()
    assert L.bh_create(C.byref(ctx), 0, None, 0) == -1
    assert L.bh_create(C.byref(ctx), 1 << 30, None, 0) == -1      # body slots are 31-bit tagged ints: n < 2^30
    assert L.bh_create(None, 10, None, 0) == -1
    assert L.bh_step(None, 1, None) == -1
    need = C.c_size_t(0)
    assert L.bh_sort_pairs_u32(None, None, None, None, 1000, 0, 32, None, C.byref(need), None) == 0
    assert need.value > 8000


def test_ic_generators_are_deterministic_and_shaped_like_the_reference(bh):
    a, b = bh.ic_refdisk(4096, 42), bh.ic_refdisk(4096, 42)
    for x, y in zip(a, b):
        assert (x == y).all()
    px, py, pz, vx, vy, vz, m = a
    r = np.sqrt(px.astype(np.float64) ** 2 + py.astype(np.float64) ** 2)
    # bench:297-307: r in [200,1700], |z| <= 0.025 r, m in [2,7], tangential speed sqrt(G(50000+100r)/r)
    assert r.min() >= 199.9 and r.max() <= 1700.1
    assert (np.abs(pz) <= 0.025 * r + 1e-3).all() and m.min() >= 2 and m.max() <= 7
    v = np.sqrt(vx.astype(np.float64) ** 2 + vy.astype(np.float64) ** 2)
    assert np.allclose(v, np.sqrt(0.5 * (50000 + 100 * r) / r), rtol=1e-4)
    assert np.abs(px * vx + py * vy).max() < 1e-2 * (r * v).max()
    u = bh.ic_uniform_cube(1000, 1, 1000.0)
    assert np.abs(u[0]).max() <= 1000 and (u[3] == 0).all()
    p = bh.ic_plummer(20000, 5, 200.0, 10.0, 4.5, 0.5)
    rr = np.sqrt(p[0].astype(np.float64) ** 2 + p[1] ** 2 + p[2] ** 2)
    assert rr.max() <= 2000.01 and (p[6] == 4.5).all()
    # Plummer half-mass radius ~ 1.305 a (slightly less with the cut)
    assert 200 * 1.15 < np.median(rr) < 200 * 1.35
