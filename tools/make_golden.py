#!/usr/bin/env python3
"""Generate tests/golden/<case>/ by running the UNMODIFIED reference (oracle/_ref/bioEM_ref,
built by oracle/Makefile from /root/reference + the FFTW-API shim) on the synthetic cases of
bioem_b200/cases.py.  Run in the build container; the outputs are committed because
/root/reference does not exist on the GPU box.

  Output_Probabilities        reference result (Algo 1, OMP_NUM_THREADS=8)
  ANG_PROB                    when the case sets WRITE_PROB_ANGLES
  debug_prob.txt (toy32 only) -DDEBUG_PROB per-evaluation stream of bioEM_ref_dbg with
                              BIOEM_DEBUG_BREAK=2 BIOEM_DEBUG_NMAPS=1 OMP_NUM_THREADS=1
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from bioem_b200.cases import build_case, reference_cli  # noqa: E402

REFBIN = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref")
DBGBIN = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref_dbg")
CASES = sys.argv[1:] or ["toy32", "toy36g2", "toy64", "cfg1", "cfg2_slice", "cfg5_slice"]

for name in CASES:
    gold = os.path.join(ROOT, "tests", "golden", name)
    os.makedirs(gold, exist_ok=True)
    with tempfile.TemporaryDirectory() as d:
        cd = build_case(name, d)
        env = {**os.environ, "OMP_NUM_THREADS": "8"}
        t = time.time()
        r = subprocess.run([REFBIN] + reference_cli(cd), cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        dt = time.time() - t
        shutil.copy(os.path.join(d, "Output_Probabilities"), gold)
        if cd.case.write_angles:
            shutil.copy(os.path.join(d, "ANG_PROB"), gold)
        ran = [l for l in r.stdout.splitlines() if "The code ran for" in l]
        print(f"{name}: {cd.case.likelihoods} likelihoods, wall {dt:.2f}s; {ran[-1] if ran else ''}")
        if name == "toy32":
            env = {**os.environ, "OMP_NUM_THREADS": "1", "BIOEM_DEBUG_BREAK": "2", "BIOEM_DEBUG_NMAPS": "1"}
            r = subprocess.run([DBGBIN] + reference_cli(cd, "dbg_out"), cwd=d, env=env,
                               capture_output=True, text=True)
            assert r.returncode == 0
            lines = [l for l in r.stdout.splitlines() if "Prob: iRefMap" in l]
            with open(os.path.join(gold, "debug_prob.txt"), "w") as f:
                f.write("\n".join(lines) + "\n")
            print(f"   debug stream: {len(lines)} evaluations")
