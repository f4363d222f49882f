#!/usr/bin/env python
"""Instruction mix (by executed warp-instructions) and top stall lines of one kernel from an .ncu-rep source page."""
import collections
import csv
import subprocess
import sys


def main(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the first line is the kernel name; the header follows
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ie, src, st = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= max(ie, src, st) or r[0] == "Address":
            continue
        try:
            data.append((float(r[ie]), float(r[st] or 0), r[src].strip()))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    tots = sum(d[1] for d in data) or 1
    print(rows[0][1][:100] if len(rows[0]) > 1 else "")
    print(f"total warp-instructions {tot:.0f}; stall samples {tots:.0f}")
    agg, aggs = collections.Counter(), collections.Counter()
    for c, s, text in data:
        toks = text.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        agg[op.split(".")[0]] += c
        aggs[op.split(".")[0]] += s
    print("opcode       inst%   stall%")
    for k, v in agg.most_common(top):
        print(f"{k:12s} {v / tot * 100:6.2f}  {aggs[k] / tots * 100:6.2f}")
    print("--- top stall instructions")
    for c, s, text in sorted(data, key=lambda d: -d[1])[:14]:
        print(f"{s / tots * 100:5.2f}%  exec {c / tot * 100:5.2f}%  {text[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
