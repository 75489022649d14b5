"""Brief per-launch table from an .ncu-rep: time, throughputs, occupancy, top stalls.  usage: ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
col = {c: i for i, c in enumerate(h)}
def g(r, k):
    try: return float(r[col[k]])
    except Exception: return float('nan')
stalls = [c for c in h if c.startswith('smsp__average_warps_issue_stalled_') and c.endswith('_per_issue_active.ratio')]
print(f"{'kernel':28s} {'us':>8s} {'dram%':>6s} {'sm%':>5s} {'issue%':>6s} {'warps%':>6s} {'regs':>4s} {'inst(M)':>8s} {'rdMB':>7s} {'wrMB':>7s}  top stalls")
for r in rows[2:]:
    name = r[col['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:28]
    st = sorted(((g(r, s), s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for s in stalls), reverse=True)[:4]
    print(f"{name:28s} {g(r,'gpu__time_duration.sum'):8.1f} {g(r,'dram__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(r,'sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {g(r,'launch__registers_per_thread'):4.0f} "
          f"{g(r,'smsp__inst_executed.sum')/1e6:8.2f} {g(r,'dram__bytes_read.sum'):7.1f} {g(r,'dram__bytes_write.sum'):7.1f}  "
          + ", ".join(f"{n} {v:.1f}" for v, n in st))
