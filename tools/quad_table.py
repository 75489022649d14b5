"""SURVEY §8f N4 / VERDICT r1 item 7: rel-RMS vs interactions/body vs ms, monopole against BH_FLAG_QUADRUPOLE, theta in
{0.5, 0.65, 0.8}, 1M-body Plummer sphere and the 1M reference disk.  -> profiles/r02_quadrupole_1m.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402
import oracle_lib as O  # noqa: E402

out = {}
for wl in ("plummer_1m", "refdisk_1m"):
    w = bench.WORKLOADS[wl]
    n = w["n"]
    soa = bench.make_ic(bh, w)
    sample = np.arange(0, n, 997, dtype=np.int32)
    rows = []
    for quad in (0, 4):
        for theta in (0.5, 0.65, 0.8):
            with bh.BHEngine(n, flags=2 | quad, theta=theta) as eng:
                eng.load_soa(*soa)
                eng.simulation_step(3)
                eng.simulation_step(10)
                ms = {k: v / 10 for k, v in eng.phase_ms().items()}
                per_body = (eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY)) / n
                acc = np.stack(eng.read_soa()[6:9], 1)[sample]
                err = O.rel_rms(acc, eng.direct_sample(sample))
                eng.check_device_error()
            rows.append({"moments": "quadrupole" if quad else "monopole", "theta": theta, "rel_rms_vs_direct_sum": err,
                         "interactions_per_body": per_body, "force_ms": ms["force"], "com_ms": ms["com"], "step_ms": ms["total"]})
            print(wl, rows[-1], flush=True)
    out[wl] = rows
json.dump(out, open(os.path.join("gpurun_out", "quadrupole_1m.json"), "w"), indent=1)
