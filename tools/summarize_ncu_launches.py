#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        rows.append((name, val * scale))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    print(f"# {len(rows)} launches, total {tot:.2f} ms (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{n[:70]:70s} {c:8d} {v:10.3f} {100 * v / tot:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
