#!/usr/bin/env python3
"""Time the direct-DFT path (image edges without a fused FFT kernel) next to the fused kernel of a neighbouring edge:
ns per likelihood, CUDA events around the likelihood kernels.  usage: generic_time.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from bioem_b200 import api, synth  # noqa: E402
from bioem_b200.cases import Case, build_case  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [63, 64, 225, 224, 490]
print("| N | path | window | likelihoods | ns / likelihood |")
print("|---|---|---|---|---|")
for n in sizes:
    maxd = min(40, n // 4)
    case = Case(f"gen{n}", n, 1.0 if n >= 200 else 1.5, 200, 16, 576, 4, synth.PRODUCTION_GRID, maxd, 1,
                model_sigma=n / 12.0, model_rmax=n / 4.0, particle_format="mrc")
    cd = build_case(case)
    hi, parts = api.inputs_for_case(cd)
    parts = np.concatenate([parts] * 10)[:148]
    eng = api.Engine(hi.cfg)
    eng.upload_all(hi, parts)
    eng.set_kernel_timing(True)
    eng.run()
    eng.synchronize()
    eng.kernel_time()
    eng.reset()
    eng.run()
    ms, _ = eng.kernel_time()
    lik = hi.O * hi.C * parts.shape[0]
    path = "fused FFT" if api.lib().bioem_b200_supported_size(n) else "direct DFT"
    print(f"| {n} | {path} | {2 * maxd + 1}² | {lik} | {1e6 * ms / lik:.1f} |", flush=True)
    eng.close()
