#!/usr/bin/env python3
"""Small end-to-end runs (several image sizes, angle table on) for compute-sanitizer."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bioem_b200 import api  # noqa: E402
from bioem_b200.cases import build_case  # noqa: E402

for name, kw in (("toy32", {}), ("toy32psf", {}), ("toy36g2", {}), ("toy64", {}), ("cfg2_slice", dict(n_particles=2, n_orient=2)),
                 ("cfg4_slice", dict(n_particles=2, n_orient=1))):
    cd = build_case(name, **kw)
    hi, parts = api.inputs_for_case(cd)
    e = api.Engine(hi.cfg, 0)
    e.upload_all(hi, parts)
    e.reset()
    e.run()
    pm, _ = e.download()
    print(name, pm["orient"], pm["conv"], flush=True)
    e.close()
