"""gpurun_out/launches_bench.csv (tools/gpu_bench_launchlist.sh) -> profiles/<tag>_launches_bench.md"""
import csv, os, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches_bench.csv"))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        d[r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")].append(float(r[vi].replace(",", "")))
ours = {k: v for k, v in d.items() if not k.startswith(("at::", "std::", "cub::", "thrust::")) and "elementwise" not in k}
step = {k: v for k, v in ours.items() if "probe" not in k}
tot, tot_step = sum(map(sum, ours.values())), sum(map(sum, step.values()))
with open(os.path.join(ROOT, "profiles", f"{tag}_launches_bench.md"), "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scale-ref` itself (tools/gpu_bench_launchlist.sh)\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600`; graph replays appear kernel by kernel; cold-cache, serialised:\n"
            "compare SHARES with the line's `phase_ms`, not absolutes.  torch's own kernels (the L2 flush fill, arange) are listed but left out of the share;\n"
            "the last column leaves out the FP32 peak probe too (it is not part of a step).\n\n"
            "| kernel | launches | avg us | share of this library's kernels | share without the peak probe |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        a = f"{100 * sum(v) / tot:.1f}%" if k in ours else "(torch)"
        b = f"{100 * sum(v) / tot_step:.1f}%" if k in step else ""
        f.write(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {a} | {b} |\n")
print("wrote", f.name)
