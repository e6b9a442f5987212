#!/usr/bin/env python3
"""Summarise an .ncu-rep of the fused kernel: key raw metrics, opcode mix, cost per barrier segment."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
nlik = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: v for h, v in zip(rows[0], rows[2])}
u = {h: v for h, v in zip(rows[0], rows[1])}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for k in keys:
    if k in d:
        print(f"{k} = {d[k]} {u[k]}")
for k in sorted(d):
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = float(d[k])
        if v >= 0.05:
            print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:22s} {v:.2f}")
if nlik:
    wi = float(d["smsp__inst_executed.sum"])
    print(f"warp instr per likelihood {wi / nlik:.0f}  (thread instr {32 * wi / nlik / 1e6:.3f} M)")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(float(r[ix["Instructions Executed"]]) for r in data)
samples = sum(float(r[ix["# Samples"]]) for r in data)
op = collections.Counter()
ops = collections.Counter()
for r in data:
    s = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    o = m.group(2).split(".")[0] if m else s
    op[o] += float(r[ix["Instructions Executed"]])
    ops[o] += float(r[ix["# Samples"]])
print("opcode mix (share of warp instructions / of stall samples):")
for o, c in op.most_common(22):
    print(f"  {o:8s} {100 * c / tot:5.1f}%  {100 * ops[o] / samples:5.1f}%")
cum = last = 0.0
ss = 0.0
print("segments between BAR.SYNC (SASS line, share of instructions, share of samples):")
for i, r in enumerate(data):
    cum += float(r[ix["Instructions Executed"]])
    ss += float(r[ix["# Samples"]])
    if "BAR.SYNC" in r[ix["Source"]] or i == len(data) - 1:
        print(f"  {i:5d} {100 * (cum - last) / tot:5.1f}%  {100 * ss / samples:5.1f}%")
        last = cum
        ss = 0.0
