#!/usr/bin/env python3
"""One profiling pass of the headline workload, in the order the measurement rules ask for.

  on the GPU box (under gpurun):   python tools/profile_pass.py gpu <label>
      1. bench.py without a profiler (must exit 0)                -> gpurun_out/bench_<label>.json
      2. ncu launch list of one bench step                          -> gpurun_out/launches_<label>.csv
      3. tools/prof_lik.py cfg2 150 without a profiler              -> gpurun_out/prof_<label>_plain.log
      4. ncu --set full of one launch of the fused kernel           -> gpurun_out/prof_<label>.ncu-rep
  here, afterwards:                python tools/profile_pass.py post <label> <round, e.g. r01>
      writes profiles/<round>_bench_<label>.json, <round>_launches_<label>_summary.csv,
      <round>_likelihood_kernel_<label>_ncu.txt and refreshes profiles/ncu_traffic.json
  any other capture:               python tools/profile_pass.py summarise <rep> <likelihoods per launch> <plain.log> <out.txt> <title> <command>
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
OUT = os.path.join(ROOT, "gpurun_out")
NLIK = 150 * 32 * 1000  # tools/prof_lik.py cfg2 150: one launch


def sh(cmd, **kw):
    print("+", cmd, flush=True)
    return subprocess.run(cmd, shell=True, cwd=ROOT, **kw)


def gpu(label):
    os.makedirs(OUT, exist_ok=True)
    r = sh(f"python bench.py > gpurun_out/bench_{label}.json 2> gpurun_out/bench_{label}.err")
    if r.returncode:
        sys.exit("bench.py failed without a profiler: not profiling")
    sh("ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv "
       f"--log-file gpurun_out/launches_{label}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e "
       f"> gpurun_out/launches_{label}.log 2>&1")
    r = sh(f"python tools/prof_lik.py cfg2 150 > gpurun_out/prof_{label}_plain.log 2>&1")
    if r.returncode:
        sys.exit("prof_lik.py failed without a profiler: not profiling")
    sh("ncu --set full --clock-control none --import-source on -k regex:likelihood_kernel -s 1 -c 1 "
       f"-o gpurun_out/prof_{label} -f python tools/prof_lik.py cfg2 150 > gpurun_out/prof_{label}_ncu.log 2>&1")


UNITS = ["SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
         "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg",
         "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg",
         "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum"]


def summarise(rep, nlik, plain, out, title, command):
    """key counters, stall table, opcode mix, unit utilisation and hottest SASS lines of one capture -> a text file"""
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(nlik)],
                          capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    with open(out, "w") as f:
        f.write(f"# {title} -- ncu --set full --clock-control none --import-source on\n")
        f.write(f"# command: {command}   (plain run of the same command first: next line)\n")
        f.write(open(plain).read())
        f.write(f"# one launch = {int(nlik):,} likelihoods\n")
        f.write(summ + "\n# unit utilisation:\n")
        for k in UNITS:
            f.write(f"{k} = {d.get(k)}\n")
        sectors = float(d.get("lts__t_sectors_srcunit_tex_op_read.sum", "nan"))
        f.write(f"L2 -> SM bytes per likelihood = {32.0 * sectors / float(nlik):.0f}\n")
        f.write("\n# hottest SASS lines (share of warp-stall samples, dominant stall reasons):\n")
        f.write("\n".join(hot.split("\n")[:14]) + "\n")
    print("written", out)


def post(label, rnd):
    prof = os.path.join(ROOT, "profiles")
    bench = os.path.join(OUT, f"bench_{label}.json")
    with open(os.path.join(prof, f"{rnd}_bench_{label}.json"), "w") as f:
        f.write(open(bench).read().strip().split("\n")[-1] + "\n")
    s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"),
                        os.path.join(OUT, f"launches_{label}.csv"), bench, f"{rnd}, {label}"], capture_output=True, text=True)
    open(os.path.join(prof, f"{rnd}_launches_{label}_summary.csv"), "w").write(s.stdout)
    rep = os.path.join(OUT, f"prof_{label}.ncu-rep")
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(NLIK)],
                          capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    units = ["SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
             "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg",
             "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
             "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg",
             "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
             "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum"]
    with open(os.path.join(prof, f"{rnd}_likelihood_kernel_{label}_ncu.txt"), "w") as f:
        f.write(f"# likelihood_kernel<224,3>, {rnd} {label} -- ncu --set full --clock-control none --import-source on\n")
        f.write("# command: ncu ... -k regex:likelihood_kernel -s 1 -c 1 python tools/prof_lik.py cfg2 150   "
                "(plain run of the same command first: next line)\n")
        f.write(open(os.path.join(OUT, f"prof_{label}_plain.log")).read())
        f.write("# one launch = 150 orientations x 32 CTFs x 1000 particles = 4,800,000 likelihoods (N = 224, 81 x 81 displacements)\n")
        f.write(summ + "\n# unit utilisation:\n")
        for k in units:
            f.write(f"{k} = {d.get(k)}\n")
        f.write("\n# hottest SASS lines (share of warp-stall samples, dominant stall reasons):\n")
        f.write("\n".join(hot.split("\n")[:14]) + "\n")
    rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
    cols = rows[0]
    ur, uw = rows[1][cols.index("dram__bytes_read.sum")], rows[1][cols.index("dram__bytes_write.sum")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tj = os.path.join(prof, "ncu_traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    def pct(k):
        try:
            return round(float(d[k]), 1)
        except Exception:
            return None
    sectors = float(d.get("lts__t_sectors_srcunit_tex_op_read.sum", "nan"))
    t["likelihood_kernel<224>"] = {
        "dram_bytes_per_likelihood": round((rd * scale[ur] + wr * scale[uw]) / NLIK, 1),
        # what actually binds (the roofline "frac" of bench.py is the algorithmic ratio of SURVEY 8d, not DRAM utilisation)
        "dram_pct_of_peak": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "lsu_data_pipe_pct": pct("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "fma_pipe_pct": pct("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_slot_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l2_to_sm_bytes_per_likelihood": round(32.0 * sectors / NLIK, 1) if sectors == sectors else None,
        "capture": f"profiles/{rnd}_likelihood_kernel_{label}_ncu.txt",
        "source": f"ncu --set full, profiles/{rnd}_likelihood_kernel_{label}_ncu.txt: dram__bytes_read.sum + dram__bytes_write.sum of one "
                  f"launch of {NLIK} likelihoods"}
    json.dump(t, open(tj, "w"), indent=1)
    print("profiles written for", rnd, label)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "gpu":
        gpu(sys.argv[2])
    elif len(sys.argv) >= 4 and sys.argv[1] == "post":
        post(sys.argv[2], sys.argv[3])
    elif len(sys.argv) >= 8 and sys.argv[1] == "summarise":
        summarise(*sys.argv[2:8])
    else:
        sys.exit(__doc__)
