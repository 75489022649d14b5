"""Small driver for ncu: a few direct-launch steps of one workload (no graph, so every kernel is a plain launch)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="refdisk_1m")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
soa = bench.make_ic(bh, w)
eng = bh.BHEngine(w["n"], flags=1)
eng.load_soa(*soa)
eng.simulation_step(a.steps)
eng.check_device_error()
n = w["n"]
print("ok", a.workload, "interactions/body", (eng.stat(bh.STAT.INTERACTIONS_CELL) + eng.stat(bh.STAT.INTERACTIONS_BODY)) / n,
      "max_stack", eng.stat(bh.STAT.MAX_STACK))
eng.close()
