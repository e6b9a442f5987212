#!/usr/bin/env python3
"""Front-end cost at the cfg4 shape: 110,592 model points (a 48^3 voxel model), N = 360, 256 CTFs.
Prints the wall time of a run next to the time inside the fused kernel."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bioem_b200 import api  # noqa: E402
from bioem_b200.cases import CASES, Case, build_case  # noqa: E402

n_or = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_pa = int(sys.argv[2]) if len(sys.argv) > 2 else 148
cd = build_case("cfg4", n_particles=n_pa, n_orient=n_or)  # 48^3 MRC volume = 110,592 points of radius 2 px
npts = cd.model.shape[0]
hi, parts = api.inputs_for_case(cd)
eng = api.Engine(hi.cfg, 0)
eng.upload_all(hi, parts)
eng.set_kernel_timing(True)
eng.reset(); eng.run(0, n_or); eng.synchronize()
eng.kernel_time()  # drain the warm-up launches
eng.reset()
t = time.time(); eng.run(0, n_or); eng.synchronize(); dt = time.time() - t
ms, n = eng.kernel_time()
lik = n_or * hi.C * parts.shape[0]
print(f"cfg4 shape, {npts} model points, {n_or} orientations x {hi.C} CTFs x {parts.shape[0]} particles: wall {dt * 1e3:.1f} ms, "
      f"fused kernel {ms:.1f} ms ({1e6 * ms / lik:.1f} ns/likelihood), front end + rest {dt * 1e3 - ms:.1f} ms "
      f"= {(dt * 1e3 - ms) / n_or:.2f} ms per orientation; at 2000 particles the fused kernel takes "
      f"{1e-6 * 1e6 * ms / lik * hi.C * 2000:.1f} ms per orientation")
eng.close()
