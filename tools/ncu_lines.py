#!/usr/bin/env python
"""Per-source-line stall samples of one ncu report:  python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep [top]"""
import csv, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
lines = [r for r in rows[h + 1:] if len(r) > si and r[0].isdigit() and r[2] == "-"]      # source-line rows (SASS rows carry an address)
tot = sum(int(r[si]) for r in lines)
print("total samples", tot, " total warp-instructions", sum(int(r[ie]) for r in lines))
agg = {}
for r in lines:
    agg.setdefault(int(r[0]), [0, 0, r[1], [0] * len(stall)])
    a = agg[int(r[0])]; a[0] += int(r[si]); a[1] += int(r[ie])
    for k, i in enumerate(stall): a[3][k] += int(r[i]) if r[i].isdigit() else 0
for ln, (s, n, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    why = sorted(zip(st, [hdr[i][6:] for i in stall]), reverse=True)[:3]
    print(f"{ln:5d} {s:7d} {100 * s / tot:5.1f}%  inst {n:9d}  {src.strip()[:90]:90s} {[(w, c) for c, w in why if c]}")
