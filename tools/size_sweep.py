#!/usr/bin/env python3
"""Fused-kernel throughput for every instantiated image edge: ns per likelihood (CUDA events around
the likelihood kernels) against the HBM roofline 8F / measured copy bandwidth."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from bioem_b200 import api, synth  # noqa: E402
from bioem_b200.cases import Case, build_case  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
bw = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
sizes = [int(a) for a in sys.argv[1:]] or [32, 36, 48, 64, 96, 128, 160, 192, 224, 256, 288, 320, 360, 384, 400]
print("| N | window | likelihoods | ns / likelihood | M likelihoods/s | HBM roofline ns | frac |")
print("|---|---|---|---|---|---|---|")
for n in sizes:
    maxd = min(40, n // 4)
    case = Case(f"sweep{n}", n, 1.0 if n >= 200 else 1.5, 200, 64, 576, 16, synth.PRODUCTION_GRID, maxd, 1,
                model_sigma=n / 12.0, model_rmax=n / 4.0, particle_format="mrc")
    cd = build_case(case)
    hi, parts = api.inputs_for_case(cd)
    parts = np.concatenate([parts] * 10)[:592]  # two images per SM
    try:
        eng = api.Engine(hi.cfg)
    except api.BioemError as e:
        print(f"| {n} | {2 * maxd + 1}² | - | {str(e)[:60]} | | | |")
        continue
    eng.upload_all(hi, parts)
    eng.set_kernel_timing(True)
    eng.run()
    eng.synchronize()
    eng.kernel_time()  # drain the warm-up launches
    eng.reset()
    eng.run()
    ms, _ = eng.kernel_time()
    lik = hi.O * hi.C * parts.shape[0]
    ns = 1e6 * ms / lik
    roof = 8.0 * n * (n // 2 + 1) / bw
    eng.close()
    print(f"| {n} | {2 * maxd + 1}² | {lik} | {ns:.1f} | {1e3 / ns:.2f} | {roof:.1f} | {roof / ns:.2f} |", flush=True)
