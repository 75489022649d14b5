"""north_star: "energy drift is reported over 1000 steps".  Total energy (kinetic + softened pair potential,
double precision O(N^2) on the device, bh_energy) every 100 steps of the engine's kick-drift-clamp."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402

runs = [("uniform_16k", bench.WORKLOADS["uniform_16k"]), ("plummer_64k", dict(n=65536, ic="plummer", desc="Plummer a=200 cut 10a")),
        ("refdisk_100k", dict(n=100_000, ic="refdisk", desc="reference disk"))]
key_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 30
out = {"key_bits": key_bits}
for name, w in runs:
    soa = bench.make_ic(bh, w)
    eng = bh.BHEngine(w["n"], key_bits=key_bits)
    eng.load_soa(*soa)
    ke0, pe0 = eng.energy()
    e0 = ke0 + pe0
    series = [{"step": 0, "ke": ke0, "pe": pe0, "rel_drift": 0.0}]
    for s in range(10):
        eng.simulation_step(100)
        ke, pe = eng.energy()
        series.append({"step": (s + 1) * 100, "ke": ke, "pe": pe, "rel_drift": (ke + pe - e0) / abs(e0)})
    eng.check_device_error()
    eng.close()
    out[name] = {"n": w["n"], "dt": 0.02, "theta": 0.5, "e0": e0, "max_abs_rel_drift": max(abs(r["rel_drift"]) for r in series),
                 "final_rel_drift": series[-1]["rel_drift"], "series": series}
print(json.dumps(out, indent=1))
