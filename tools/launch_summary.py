#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: launch_summary.py <launches.csv> <plain_bench.json> <label> > profiles/rNN_launches_<label>_summary.csv"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
data = rows[rows.index(hdr) + 1:]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in data:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v if u.startswith("m") else v * 1e3
    a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
plain = json.load(open(sys.argv[2]))
print(f"# ncu launch list, {sys.argv[3]}")
print("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv python bench.py --steps 1 --warmup 0 "
      "--no-cpu-baseline --no-e2e")
print(f"# (bench.py of the same build ran to completion without ncu first -- default arguments, 3 warm-up + 2 timed steps: {plain['value'] / 1e6:.2f} M likelihoods/s, "
      f"{plain['ms_per_step'] / 1e3:.2f} s/step, kernel share {plain['roofline']['kernel_share_of_step']}).")
print("# Per-launch times under ncu are serialised/cold: compare SHARES.  " + plain["config"]["workload"])
print("kernel,launches,total_ms,share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'"{k}",{a[0]},{a[1]:.3f},{a[1] / tot:.4f}')
