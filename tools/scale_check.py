#!/usr/bin/env python3
"""Large-shape sanity run (cfg3 / cfg5 scale): many particles, angle table on; checks that the
results are finite, that evaluating the orientation range in one call or in two gives the same
records, and prints the throughput."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from bioem_b200 import api  # noqa: E402
from bioem_b200.cases import CASES, Case, build_case  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
O = int(sys.argv[2]) if len(sys.argv) > 2 else 100
base = CASES["cfg2"]
small = build_case(Case(**{**base.__dict__, "n_particles": 64, "n_orient": O, "write_angles": 5}))
hi, parts = api.inputs_for_case(small)
# tile the 64 synthetic particles (with a per-copy scale so that they are not identical)
rng = np.random.default_rng(0)
reps = (M + 63) // 64
big = np.concatenate([parts * np.float32(rng.uniform(0.5, 2.0)) for _ in range(reps)])[:M]
eng = api.Engine(hi.cfg, 0)
eng.upload_model(hi.points, hi.NormDen)
eng.upload_orientations(hi.angles)
eng.upload_ctf(hi.refCTF, hi.CtfParam)
t = time.time()
eng.upload_particles(big)
print(f"upload {M} particles: {time.time() - t:.2f} s")
eng.reset()
t = time.time()
eng.run()
eng.synchronize()
dt = time.time() - t
a, aa = eng.download()
print(f"run: {O} x {hi.C} x {M} = {O * hi.C * M / 1e6:.1f} M likelihoods in {dt:.2f} s = {O * hi.C * M / dt / 1e6:.2f} M/s")
eng.reset()
eng.run(0, O // 3)
eng.run(O // 3, O)
b, bb = eng.download()
assert np.isfinite(a["Total"]).all() and np.isfinite(a["Constoadd"]).all() and (a["Total"] > 0).all()
for k in ("Constoadd", "cent_x", "cent_y", "orient", "conv", "norm", "mu"):
    assert np.array_equal(a[k], b[k]), k
np.testing.assert_allclose(a["Total"], b["Total"], rtol=1e-12)
assert np.array_equal(aa["ConstAngle"], bb["ConstAngle"]) and np.allclose(aa["forAngles"], bb["forAngles"], rtol=1e-12)
# every image of a tiled copy is a scaled copy of one of the 64: same arg-max orientation
assert np.array_equal(a["orient"][:64], a["orient"][64:128])
print("ok: finite, split == single, angle table consistent, tiled copies agree")
eng.close()
