"""Time the reference's own CUDA path (bioem_cuda.cu + cuFFT, rebuilt for sm_100a as
oracle/_ref/bioEM_ref_cuda) on a slice of a workload on this box's GPU, and compare its
Output_Probabilities with the ones of bioEM_b200 on the same files.

    python tools/ref_cuda_time.py [workload] [n_orient] [n_particles]

TEST / MEASUREMENT INFRASTRUCTURE: not imported by the product path.
"""
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bioem_b200.cases import build_case, reference_cli  # noqa: E402


def parse_probs(path):
    rows = {}
    for ln in open(path):
        m = re.match(r"RefMap:\s+(\d+)\s+LogProb:\s+(\S+)\s+Constant:", ln)
        if m:
            rows.setdefault(int(m.group(1)), {})["logp"] = float(m.group(2))
        m = re.match(r"RefMap:\s+(\d+)\s+Maximizing Param:\s+(.*)", ln)
        if m:
            rows.setdefault(int(m.group(1)), {})["max"] = m.group(2).split()
    return rows


def run(binary, cd, workdir, env, out):
    t = time.time()
    r = subprocess.run([binary] + reference_cli(cd, out), cwd=workdir, env=env, capture_output=True, text=True,
                       timeout=900)
    wall = time.time() - t
    if r.returncode != 0:
        raise RuntimeError(f"{binary} failed: {r.stdout[-800:]}{r.stderr[-800:]}")
    sec = None
    for ln in r.stdout.splitlines():
        if "The code ran for" in ln and sec is None:
            sec = float(ln.split("for")[1].split("seconds")[0])
        if "Likelihood path (upload .. merge):" in ln:  # bioEM_b200: the span of the reference's run()
            sec = float(ln.split(":")[1].split("seconds")[0])
    return sec, wall, r.stdout


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n_or = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    n_part = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    refbin = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref_cuda")
    ours = os.path.join(ROOT, "bioem_b200", "bin", "bioEM_b200")
    with tempfile.TemporaryDirectory() as d:
        cd = build_case(workload, d, n_particles=n_part, n_orient=n_or)
        n = cd.case.likelihoods
        env = {**os.environ, "GPU": "1", "GPUWORKLOAD": "100", "GPUDEVICE": "0", "BIOEM_DEBUG_OUTPUT": "1",
               "OMP_NUM_THREADS": str(os.cpu_count() or 1)}
        sec, wall, out = run(refbin, cd, d, env, "ref_cuda_out")
        for ln in out.splitlines():
            if any(k in ln for k in ("SUMMARY", "GPU", "ran for")):
                print("   ref|", ln)
        print(f"reference CUDA path (bioem_cuda.cu + cuFFT, sm_100a): {n} likelihoods in {sec:.3f} s "
              f"(its own timer; process wall {wall:.2f} s) = {n / sec / 1e6:.3f} M likelihoods/s")
        env2 = {k: v for k, v in os.environ.items() if k not in ("GPU", "GPUWORKLOAD")}
        env2["BIOEM_B200_GPUS"] = "1"
        sec2, wall2, _ = run(ours, cd, d, env2, "ours_out")
        print(f"bioEM_b200 on the same files: {sec2:.3f} s (upload .. merge, the span of the reference's run(); process wall {wall2:.2f} s) = "
              f"{n / sec2 / 1e6:.3f} M likelihoods/s")
        a, b = parse_probs(os.path.join(d, "ref_cuda_out")), parse_probs(os.path.join(d, "ours_out"))
        worst, differ = 0.0, 0
        for m in a:
            worst = max(worst, abs(a[m]["logp"] - b[m]["logp"]) / max(1.0, abs(a[m]["logp"])))
            # first token = the maximum log posterior itself (printed with 6 significant digits); the rest
            # are the maximizing orientation / CTF / displacement
            if a[m]["max"][1:] != b[m]["max"][1:]:
                differ += 1
                if differ <= 3:
                    print("   differ: RefMap", m, "\n      ref ", a[m]["max"], "\n      ours", b[m]["max"])
        print(f"outputs: {len(a)} particles, max relative log P difference {worst:.2e}, "
              f"{differ} maximizing-parameter lines differ")


if __name__ == "__main__":
    main()
