"""gpurun_out/ ncu artefacts of tools/gpu_r2_c.sh -> tracked summaries under profiles/ (round 2)."""
import csv, json, os, re, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
traffic = {}
for W in ("refdisk_1m", "plummer_16m"):
    src = os.path.join(OUT, f"launches_{W}.csv")
    if os.path.exists(src):
        rows = list(csv.reader(open(src)))
        hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        h = rows[hdr]; ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
        d = defaultdict(lambda: defaultdict(list))
        for r in rows[hdr + 1:]:
            if len(r) > vi:
                k = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
                d[k][r[mi]].append(float(r[vi].replace(",", "")))
        tot = sum(sum(v["gpu__time_duration.sum"]) for v in d.values())
        with open(os.path.join(PROF, f"{tag}_launches_{W}.md"), "w") as f:
            f.write(f"# ncu launch list — {W}, import + 2 direct-launch steps (tools/gpu_launchlist2.sh)\n\n"
                    "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`;\n"
                    "per-launch times are cold-cache and serialised (the first step after an import also sorts unsorted bodies),\n"
                    "so compare SHARES with bench.py's phase_ms, not absolutes.\n\n"
                    "| kernel | launches | avg us | share | DRAM read MB | DRAM write MB | DRAM GB/s |\n|---|---:|---:|---:|---:|---:|---:|\n")
            for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
                t = v["gpu__time_duration.sum"]; n = len(t)
                rd, wr = sum(v["dram__bytes_read.sum"]) / n, sum(v["dram__bytes_write.sum"]) / n
                f.write(f"| {k} | {n} | {sum(t)/n/1e3:.1f} | {100*sum(t)/tot:.1f}% | {rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/(sum(t)/n):.0f} |\n")
    rep = os.path.join(OUT, f"prof_force_{W}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    keep = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__thread_inst_executed_per_inst_executed.ratio"]
    keep += [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
    vals = {n: (v[i], u[i]) for i, n in enumerate(h) if n in keep}
    def num(n):
        x, unit = vals[n]; x = float(x.replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(unit, 1)
    traffic[W] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(sass.splitlines()))
    hdr = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
    sh = srows[hdr]; si, ii = sh.index("Source"), sh.index("Instructions Executed")
    ops = defaultdict(int); total = 0; body = srows[hdr + 1:]
    for r in body:
        if len(r) <= ii: continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si].strip())
        op = m.group(2) if m else r[si].strip()
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDG", "LDS", "STS", "MUFU", "SHFL", "REDUX", "VOTE")) else op.split(".")[0]
        n = int(r[ii] or 0); ops[op] += n; total += n
    # the tile loop = the run of instructions sharing the largest execution count
    top = max(int(r[ii] or 0) for r in body if len(r) > ii)
    loop = [r[si].strip() for r in body if len(r) > ii and int(r[ii] or 0) == top]
    with open(os.path.join(PROF, f"{tag}_force_{W}.md"), "w") as f:
        f.write(f"# ncu --set full — force_kernel, {W} (tools/gpu_r2_c.sh; second step after the import)\n\n| metric | value | unit |\n|---|---:|---|\n")
        for n in keep:
            if n in vals: f.write(f"| {n} | {vals[n][0]} | {vals[n][1]} |\n")
        f.write(f"\n## opcode histogram (warp instructions executed, {total:,} in the SASS view)\n\n| opcode | count | share |\n|---|---:|---:|\n")
        for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:24]:
            f.write(f"| {op} | {n:,} | {100*n/total:.1f}% |\n")
        f.write(f"\n## the interaction loop ({len(loop)} instructions per 8 sources x 32 bodies, executed {top:,} times)\n\n```\n" + "\n".join(loop) + "\n```\n")
    print("wrote", W)
json.dump({**{k: v for k, v in traffic.items()}, "source": f"dram__bytes_read.sum + dram__bytes_write.sum of one force_kernel launch, ncu --set full, profiles/{tag}_force_<workload>.md"},
          open(os.path.join(PROF, "force_traffic.json"), "w"), indent=1)
