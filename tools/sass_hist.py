"""Opcode histogram + hottest instructions of an `ncu --page source --csv --print-source sass` export (tuning helper)."""
import csv, sys, re
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hdr]
si, ii, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, samples = defaultdict(int), defaultdict(int)
tot = tots = 0
body = rows[hdr + 1:]
for r in body:
    if len(r) <= ii: continue
    src = r[si].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDG", "LDS", "STS", "STG", "MUFU", "SHFL", "REDUX", "VOTE")) else op.split(".")[0]
    n = int(r[ii] or 0); s = int(r[sm] or 0)
    ops[op] += n; samples[op] += s; tot += n; tots += s
print(f"total warp instructions {tot:,}  samples {tots:,}")
for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{op:14s} {n:14,d} {100*n/tot:5.1f}%   samples {100*samples[op]/max(tots,1):5.1f}%")
if len(sys.argv) > 2:
    print("\nhottest instructions by samples")
    for r in sorted(body, key=lambda r: -int(r[sm] or 0))[:int(sys.argv[2])]:
        print(f"{int(r[sm]):7d} {int(r[ii] or 0):12,d}  {r[si].strip()[:110]}")
