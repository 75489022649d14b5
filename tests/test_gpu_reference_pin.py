"""Pins Oracle-L against the UNMODIFIED reference: oracle/_ref/libref_step.so is
/root/reference/nbody_v5_bench.cu's kernels + simulationStep() compiled for sm_100
(oracle/ref_wrap.cu).  Bounds, keys and the sort permutation must match bit for bit; the
accelerations must be the whole-system monopole of SURVEY F2 to float-atomic noise."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
REF = os.path.join(O.ROOT, "oracle", "_ref", "libref_step.so")


def run_reference(soa, nsteps=1):
    L = C.CDLL(REF)
    n = len(soa[0])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.ref_init(n, *[p(np.ascontiguousarray(a, np.float32)) for a in soa]) == 0
    ms = C.c_float()
    assert L.ref_step(nsteps, C.byref(ms)) == 0
    out = {}
    names = ["px", "py", "pz", "vx", "vy", "vz", "ax", "ay", "az"]
    for i, nm in enumerate(names):
        out[nm] = np.zeros(n, np.float32)
        assert L.ref_get(i, p(out[nm])) == 0
    out["keys"], out["idx"], out["bounds"] = np.zeros(n, np.uint32), np.zeros(n, np.int32), np.zeros(6, np.float32)
    assert L.ref_get(10, p(out["keys"])) == 0 and L.ref_get(11, p(out["idx"])) == 0 and L.ref_get(12, p(out["bounds"])) == 0
    cnt = np.zeros(1, np.int32)
    L.ref_get(13, p(cnt))
    out["nodes"], out["ms"] = int(cnt[0]), float(ms.value)
    L.ref_free()
    return out


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("kind,n", [("uniform", 16384), ("disk", 100_000)])
def test_oracle_l_matches_the_real_reference(bh, kind, n):
    soa = bh.ic_uniform_cube(n, 42, 1000.0) if kind == "uniform" else bh.ic_refdisk(n, 42)
    ref = run_reference(soa, 1)
    lit = O.reference_step(soa, 1, fixed=0)
    assert ref["bounds"].tobytes() == lit["bounds"].tobytes()
    assert (ref["keys"] == lit["keys"]).all()            # sorted keys of the step
    assert (ref["idx"] == lit["idx"]).all()              # stable permutation
    racc = np.stack([ref["ax"], ref["ay"], ref["az"]], 1)
    lacc = np.stack([lit["ax"], lit["ay"], lit["az"]], 1)
    assert O.rel_rms(racc, lacc) < 1e-3                  # float atomics in another order (bench:167-170)
    for k in ("px", "py", "pz", "vx", "vy", "vz"):
        assert np.allclose(ref[k], lit[k], rtol=1e-4, atol=1e-3)
    # F2: one interaction per body == the monopole of everything
    m = soa[6].astype(np.float64)
    com = [(soa[a].astype(np.float64) * m).sum() / m.sum() for a in range(3)]
    d = np.stack([com[a] - soa[a] for a in range(3)], 1)
    mono = 0.5 * m.sum() * d / ((d ** 2).sum(1) + 50.0)[:, None] ** 1.5
    assert O.rel_rms(racc, mono) < 1e-3
    # and the engine's integrator reproduces the reference's, given the reference's accelerations
    want = O.integrate(*soa[:6], ref["ax"], ref["ay"], ref["az"])
    for k, w in zip(("px", "py", "pz", "vx", "vy", "vz"), want):
        assert ref[k].tobytes() == w.tobytes()
