"""30-bit (the reference key) against 60-bit Morton keys on ONE GPU: interactions per body, step time, cells and
accuracy against the on-device double-precision direct sum.  usage: keybits_compare.py IC N [STEPS]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402

ic, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
soa = bench.make_ic(bh, dict(n=n, ic=ic))
sample = np.arange(0, n, max(1, n // 1024), dtype=np.int32)[:1024]
out = {"ic": ic, "n": n, "sample": len(sample), "rows": []}
ref = None
for bits in (30, 60):
    eng = bh.BHEngine(n, key_bits=bits, flags=2)
    eng.load_soa(*soa)
    eng.simulation_step(1)
    res = eng.read_soa()
    acc = np.stack(res[6:9], 1)[sample].astype(np.float64)
    if ref is None:
        ref = eng.direct_sample(sample)      # positions of step 0 (the state was sorted, not yet needed again)
    err = float(np.sqrt(((acc - ref) ** 2).sum() / (ref ** 2).sum()))
    inter = (eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY)) / n
    cells = eng.stat(bh.STAT.CELLS)
    eng.simulation_step(steps)
    torch.cuda.synchronize()
    ph = {k: v / steps for k, v in eng.phase_ms().items()}
    eng.check_device_error()
    out["rows"].append({"key_bits": bits, "rel_rms_vs_direct": err, "interactions_per_body": inter, "cells": cells,
                        "phase_ms": {k: round(v, 3) for k, v in ph.items()}, "body_steps_per_s": n / (ph["total"] * 1e-3)})
    eng.close()
    del eng
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
