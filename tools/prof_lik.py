#!/usr/bin/env python3
"""Short profiling driver: cfg2 inputs, a few orientation batches (default 45 orientations =
3 launches of the fused kernel).  Used under ncu; prints kernel time per likelihood."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bioem_b200 import api  # noqa: E402
from bioem_b200.cases import build_case  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n_or = int(sys.argv[2]) if len(sys.argv) > 2 else 45
n_part = int(sys.argv[3]) if len(sys.argv) > 3 else None
cd = build_case(name, n_particles=n_part, n_orient=n_or if n_part else None)
hi, parts = api.inputs_for_case(cd)
eng = api.Engine(hi.cfg, 0)
eng.upload_all(hi, parts)
eng.set_kernel_timing(True)
eng.reset()
eng.run(0, min(n_or, hi.O))
eng.synchronize()
eng.kernel_time()  # drain the warm-up launch
eng.reset()
t = time.time()
eng.run(0, min(n_or, hi.O))
ms, n = eng.kernel_time()
lik = min(n_or, hi.O) * hi.C * parts.shape[0]
print(f"{name}: {lik} likelihoods, {n} launches, kernel {ms:.2f} ms, {1e6 * ms / lik:.1f} ns/likelihood, "
      f"{lik / ms / 1e3:.2f} M/s")
eng.close()
