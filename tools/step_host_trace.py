"""Device timeline of bh_step_host (needs a libbh built with -DBH_TRACE_STEP_HOST, selected with BH_LIB)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import nbody_barnes_hut_cuda_b200 as bh
n = 1_000_000
px, py, pz, vx, vy, vz, m = bh.ic_refdisk(n)
harr = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in (px, py, pz, vx, vy, vz, m)]
eng = bh.BHEngine(n, flags=int(os.environ.get('BH_TRACE_FLAGS', '0')))
for i in range(8):
    t0 = time.perf_counter()
    eng.step_host(*harr, nsteps=1)
    print(f"call {i}: wall {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
