#!/usr/bin/env python3
"""Parity report: GPU path vs CPU oracle on the synthetic cases — max |d logP|, relative error,
images whose arg-max record (orient, conv, cent_x, cent_y) is identical, and for the others the
oracle's logpro gap at the GPU's choice (near-tie classification)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from bioem_b200 import api  # noqa: E402
from bioem_b200.cases import build_case  # noqa: E402
from oracle import pyoracle  # noqa: E402

names = sys.argv[1:] or ["toy32", "toy32psf", "toy32g2odd", "toy32g3", "toy36g2", "toy64", "cfg1", "cfg2_slice", "cfg5_slice",
                         "cfg4_slice", "cfg4_voxel_slice"]
print("| case | N | images | likelihoods | max abs d logP | max rel d logP | identical arg-max | near-tie gaps (oracle logpro) |")
print("|---|---|---|---|---|---|---|---|")
for name in names:
    cd = build_case(name)
    hi, parts = api.inputs_for_case(cd)
    eng = api.Engine(hi.cfg)
    eng.upload_all(hi, parts)
    eng.reset()
    eng.run()
    pm, _ = eng.download()
    eng.close()
    P = pyoracle.Prepared(cd.case, cd.model, cd.quats, cd.particles)
    ref = P.run()["prob"]
    dabs = drel = 0.0
    same = 0
    gaps = []
    for m in range(P.M):
        lg = hi.final_logprob(pm[m]["Total"], pm[m]["Constoadd"])
        lo = P.final_logprob(ref[m]["Total"], ref[m]["Constoadd"])
        dabs = max(dabs, abs(lg - lo))
        drel = max(drel, abs(lg - lo) / abs(lo))
        if all(pm[m][k] == ref[m][k] for k in ("orient", "conv", "cent_x", "cent_y")):
            same += 1
        else:
            lp = P.logpro_at(m, int(pm[m]["orient"]), int(pm[m]["conv"]), int(pm[m]["cent_x"]), int(pm[m]["cent_y"]))
            gaps.append(ref[m]["Constoadd"] - lp)
    print(f"| {name} | {cd.case.n_pixels} | {P.M} | {cd.case.likelihoods} | {dabs:.2e} | {drel:.1e} | {same}/{P.M} | "
          f"{', '.join(f'{g:.3g}' for g in gaps) or '-'} |", flush=True)
