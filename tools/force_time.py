"""Phase timing of one workload for the library selected by BH_LIB (tuning helper)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "refdisk_1m"
w = bench.WORKLOADS[wl]
soa = bench.make_ic(bh, w)
kw = {"group_split": float(os.environ["BH_SPLIT"])} if "BH_SPLIT" in os.environ else {}
eng = bh.BHEngine(w["n"], flags=2, **kw)
eng.load_soa(*soa)
eng.simulation_step(5)
eng.simulation_step(20)
ms = eng.phase_ms()
print(os.environ.get("BH_LIB", "default").split("/")[-1], "split", os.environ.get("BH_SPLIT", "-"), wl, {k: round(v / 20, 4) for k, v in ms.items()},
      "int/body", (eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY)) / w["n"],
      "cell/body", eng.stat(bh.STAT.INTERACTIONS_CELL) / w["n"], "direct/body", eng.stat(bh.STAT.INTERACTIONS_BODY) / w["n"],
      "cells/n", eng.stat(bh.STAT.CELLS) / w["n"])
import numpy as np
meta = eng.debug_get(bh.DBG.CELL_META)
bucket = ((meta[:, 2] >> 8) & 1) == 1
print("buckets", int(bucket.sum()), "bodies in buckets", int(meta[bucket, 1].sum()), "max bucket", int(meta[bucket, 1].max()) if bucket.any() else 0,
      "mean bucket", float(meta[bucket, 1].mean()) if bucket.any() else 0)
eng.check_device_error()
