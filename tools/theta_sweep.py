"""BASELINE.json configs[2]: 1,000,000-body Plummer sphere, theta 0.3 / 0.5 / 0.8 — accuracy against the
on-device double-precision direct sum (4,096-body sample) and throughput.  Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "plummer_1m"
w = bench.WORKLOADS[wl]
n = w["n"]
soa = bench.make_ic(bh, w)
sample = np.arange(0, n, n // 4096, dtype=np.int32)[:4096]
out = {"workload": wl, "n": n, "sample": len(sample), "rows": []}
for theta in (0.3, 0.5, 0.8):
    eng = bh.BHEngine(n, theta=theta)
    eng.load_soa(*soa)
    eng.simulation_step(1)
    res = eng.read_soa()
    acc = np.stack(res[6:9], 1)[sample].astype(np.float64)
    ref = eng.direct_sample(sample)
    err = float(np.sqrt(((acc - ref) ** 2).sum() / (ref ** 2).sum()))
    inter = (eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY)) / n
    eng.simulation_step(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.simulation_step(20, torch.cuda.current_stream().cuda_stream)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    eng.check_device_error()
    out["rows"].append({"theta": theta, "rel_rms_vs_direct": err, "interactions_per_body": inter, "ms_per_step": ms,
                        "body_steps_per_s": n / (ms * 1e-3), "interactions_per_s": inter * n / (ms * 1e-3)})
    eng.close()
print(json.dumps(out, indent=1))
