#!/usr/bin/env python3
"""Hot SASS lines of an .ncu-rep source page: samples, executed count and dominant stall reasons.
usage: ncu_hot.py rep [top] [lo hi]   (lo..hi = SASS line range to dump in order)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[ix["# Samples"]]) for r in data)


def fmt(i, r):
    s = float(r[ix["# Samples"]])
    st = sorted(((float(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    return (f"{i:5d} {100 * s / tot:5.2f}% ex={float(r[ix['Instructions Executed']]):.3g} "
            f"{r[ix['Source']].strip()[:70]:70s} " + " ".join(f"{n}:{v:.0f}" for v, n in st if v > 0))


if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for i in range(lo, hi):
        print(fmt(i, data[i]))
else:
    order = sorted(range(len(data)), key=lambda i: -float(data[i][ix["# Samples"]]))[:top]
    for i in order:
        print(fmt(i, data[i]))
