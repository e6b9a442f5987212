#!/usr/bin/env python3
"""Where the end-to-end overhead of one public-API pass goes (create / uploads / run / download / destroy)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bioem_b200 import api  # noqa: E402
from bioem_b200.cases import build_case  # noqa: E402

n_or = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cd = build_case("cfg2")
hi, parts = api.inputs_for_case(cd)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
h_parts, h_ctf, h_ang = pin(parts), pin(hi.refCTF), pin(hi.angles)
out = np.zeros(parts.shape[0], dtype=api.PROB_MAP_DTYPE)
for rep in range(2):
    t = [time.perf_counter()]
    names = []

    def mark(n):
        torch.cuda.synchronize()
        t.append(time.perf_counter())
        names.append(n)

    e = api.Engine(hi.cfg, 0)
    mark("create")
    e.upload_model(hi.points, hi.NormDen)
    e.upload_orientations(h_ang)
    mark("model+orient")
    e.upload_ctf(h_ctf, hi.CtfParam)
    mark("ctf")
    e.upload_particles(h_parts)
    mark("particles")
    e.reset()
    e.run(0, n_or)
    mark(f"run({n_or})")
    e.download(out)
    mark("download")
    e.close()
    mark("destroy")
    print(" | ".join(f"{n} {1e3 * (b - a):.1f} ms" for n, a, b in zip(names, t[:-1], t[1:])))
