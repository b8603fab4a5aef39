#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/launches_<tag>.csv > profiles/launches_<tag>.md
"""
import collections
import csv
import sys


def main(path: str) -> None:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui].replace("second", "s").replace("n", "n"), 1e-3)
        key = r[ki].split("(")[0]
        a = agg.setdefault(key, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {path}\n")
    print("Per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total us | avg us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, (n, v, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {v:.1f} | {v / n:.1f} | {100 * v / tot:.1f}% | {g} | {b} |")


if __name__ == "__main__":
    main(sys.argv[1])
