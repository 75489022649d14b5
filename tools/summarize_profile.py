"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (per round)."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
workload = sys.argv[2] if len(sys.argv) > 2 else "refdisk_1m"
kernel = sys.argv[3] if len(sys.argv) > 3 else "force"

os.makedirs(PROF, exist_ok=True)
# ---- launch list ------------------------------------------------------------------------
src = os.path.join(OUT, f"launches_{workload}.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            d[r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(PROF, f"{tag}_launches_{workload}.md"), "w") as f:
        f.write(f"# ncu launch list — {workload}, 3 direct-launch steps (tools/gpu_profile.sh)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and\n"
                "serialised, so compare SHARES with bench.py's phase_ms, not absolutes.\n\n")
        f.write("| kernel | launches | avg us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot * 100:.1f}% |\n")
    print("wrote launch list")

# ---- full capture of one kernel ---------------------------------------------------------------
rep = os.path.join(OUT, f"prof_{kernel}_{workload}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    keep = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "smsp__warps_eligible.avg.per_cycle_active"]
    keep += [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
    vals = {n: (v[i], u[i]) for i, n in enumerate(h) if n in keep}
    with open(os.path.join(PROF, f"{tag}_{kernel}_{workload}.md"), "w") as f:
        f.write(f"# ncu --set full — {kernel} kernel, {workload} (tools/gpu_profile.sh, second step)\n\n")
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for n in keep:
            if n in vals:
                f.write(f"| {n} | {vals[n][0]} | {vals[n][1]} |\n")
    def num(name):
        return float(vals[name][0].replace(",", "")) if name in vals else None
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd = num("dram__bytes_read.sum") * unit.get(vals["dram__bytes_read.sum"][1], 1.0)
    wr = num("dram__bytes_write.sum") * unit.get(vals["dram__bytes_write.sum"][1], 1.0)
    tj = os.path.join(PROF, "force_traffic.json")
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    if kernel == "force":
        cur[workload] = rd + wr
        json.dump(cur, open(tj, "w"), indent=1)
    print("wrote kernel summary; dram bytes/launch =", rd + wr)
